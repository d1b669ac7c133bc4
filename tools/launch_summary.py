"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for row in r:
    name = re.sub(r"\(.*", "", row[ki])
    v = float(row[vi].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000, "s": v * 1e6}.get(row[ui], v)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':62s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'max_us':>9s} share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:62]:62s} {a[0]:5d} {a[1]:11.1f} {a[1] / a[0]:9.1f} {a[2]:9.1f} {a[1] / tot:.3f}")
print(f"{'TOTAL':62s} {sum(a[0] for a in agg.values()):5d} {tot:11.1f}")
