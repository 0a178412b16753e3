#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
total = sum(sum(v) for v in agg.values())
print(f"{'launches':>8} {'mean us':>10} {'share':>7}  kernel")
for k, v in agg.items():
    print(f"{len(v):8d} {sum(v) / len(v) / 1e3:10.2f} {100 * sum(v) / total:6.1f}%  {k[:110]}")
