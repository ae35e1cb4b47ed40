"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: python profiles/launch_summary.py FILE.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki].split("(")[0][:90], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi]) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:90s} {v[0]:5d} {v[1]:11.1f} us {100 * v[1] / tot:6.1f}%  {v[1] / v[0]:9.1f} us/launch")
