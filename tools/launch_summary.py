"""Offline: per-kernel totals of the LAST UNet forward + VJP in an ncu launch list (gpu__time_duration pass).
Usage: python tools/launch_summary.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
names = [(r[ki].split('(')[0].replace('kdip::', '').replace('void ', '').split('<')[0], float(r[vi].replace(',', '')) / 1e3)
         for r in data if len(r) > vi]
idx = [i for i, (n, _) in enumerate(names) if n.startswith('im2col')]
seq = names[idx[-2]:] if len(idx) >= 2 else names
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for n, t in seq:
    a = agg[n]; a[0] += 1; a[1] += t; a[2] = max(a[2], t)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} n={a[0]:4d} total={a[1]/1e3:8.3f} ms {100*a[1]/tot:5.1f}% max={a[2]:8.1f} us")
print(f"total {tot/1e3:.3f} ms over {len(seq)} launches")
