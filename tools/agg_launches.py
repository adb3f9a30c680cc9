"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, block, grid)."""
import csv, collections, sys
f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(f)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
data = rows[hdr + 1:]
agg = collections.OrderedDict()
for r in data:
    k = (r[4][:64], r[7], r[8])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
tot = sum(a[1] for a in agg.values())
print(f, len(data), 'launches', round(tot / 1e3, 1), 'us total')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{a[1]/1e3:9.1f} us {a[0]:4d} x {a[1]/a[0]/1e3:7.1f}  {k[0]} {k[1]} {k[2]}")
