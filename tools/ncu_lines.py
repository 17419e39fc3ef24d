"""Per-source-line instruction and stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = {}; fn = 0; fname = ""
hdr = None
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r
    if r and r[0].isdigit() and hdr and len(r) > 8 and r[7].isdigit():
        k = (fname, int(r[0]))
        o = out.setdefault(k, [r[1], 0, 0, {}])
        o[1] += int(r[7]); o[2] += int(r[6])
tot = sum(o[1] for o in out.values()); ts = sum(o[2] for o in out.values())
print("instructions", tot, "samples", ts)
for (f, ln), (src, n, s, _) in sorted(out.items()):
    if n > tot * thr / 100 or s > ts * thr / 100:
        print(f"{f[:14]:14s}{ln:5d} {n / tot * 100:6.2f}% smp {s / ts * 100:5.2f}% | {src.strip()[:130]}")
