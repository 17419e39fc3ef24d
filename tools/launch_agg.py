"""Per-kernel totals of an ncu launch list (second half of the launches = warmed-up steps).  usage: launch_agg.py CSV [DIVISOR]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mn = h.index('Metric Name'); mv = h.index('Metric Value')
L = [(r[kn], float(r[mv].replace(',', '')) / 1e3) for r in rows[hi + 1:] if len(r) > mv and r[mn] == 'gpu__time_duration.sum']
half = L[len(L) // 2:]
agg = collections.defaultdict(list)
for k, t in half: agg[k.split('(')[0][-56:]].append(t)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:58s} n={len(v):4d} mean {sum(v) / len(v):8.1f} us  per unit {sum(v) / div:9.1f} us")
