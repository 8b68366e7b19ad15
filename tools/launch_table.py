"""Per-kernel totals of an ncu launch list (csv with gpu__time_duration.sum, sm__inst_executed.sum, issue active)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ii = h.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki]), {})[r[mi]] = float(r[vi].replace(',', ''))
agg = collections.OrderedDict()
for (i, kn), v in d.items():
    short = kn.split('(')[0][-45:]
    a = agg.setdefault(short, [0, 0, 0, 0])
    a[0] += v['gpu__time_duration.sum'] / 1e3
    a[1] += v.get('sm__inst_executed.sum', 0) / 1e6
    a[2] += 1
    a[3] = max(a[3], v.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0))
tot = sum(a[0] for a in agg.values())
for k, a in agg.items():
    print(f"{k:48s} n={a[2]:2d} t={a[0]:8.1f} us ({100*a[0]/tot:4.1f}%)  inst={a[1]:7.1f}M issue%max={a[3]:.0f}")
print(f"{'total':48s}      t={tot:8.1f} us")
