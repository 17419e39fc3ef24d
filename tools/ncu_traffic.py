"""profiles/traffic.json from ncu --set full captures: DRAM bytes (read + write) per launch of the dominant kernel of a
bench workload.  usage: ncu_traffic.py WORKLOAD=REPORT.ncu-rep:KERNEL_SUBSTRING ...   (merges into the existing file)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    out = json.load(open(path))
except (OSError, ValueError):
    out = {}
for arg in sys.argv[1:]:
    wl, rest = arg.split("=", 1)
    rep, kern = rest.rsplit(":", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    vals = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kern not in d.get("Kernel Name", ""):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u = units[hdr.index(k)].lower()
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            tot += float(d[k].replace(",", "")) * scale
        vals.append(tot)
    if not vals:
        print(f"{wl}: no launch of {kern} in {rep}", file=sys.stderr)
        continue
    out.setdefault(wl, {})[kern] = sum(vals) / len(vals)
    out[wl][kern + "__source"] = os.path.basename(rep) + f" ({len(vals)} launch(es), ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
