"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    elif unit in ("s", "second"): v *= 1e9
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += v
s = sum(v[1] for v in tot.values())
print("%-60s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %6d %12.1f %6.1f%%" % (k[:60], c, v / 1e3, 100 * v / s))
print("%-60s %6s %12.1f" % ("TOTAL", "", s / 1e3))
