#!/usr/bin/env python
"""stdin: `ncu --metrics gpu__time_duration.sum --csv` output -> per kernel name: launches, mean / min / max duration (us), share."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in sys.stdin if l.startswith('"'))]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
acc = OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    acc.setdefault(name, []).append(float(r[iv].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in acc.values())
print("%-40s %6s %10s %10s %10s %7s" % ("kernel", "n", "mean us", "min us", "max us", "share"))
for k, v in acc.items():
    print("%-40s %6d %10.2f %10.2f %10.2f %6.1f%%" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
print("%-40s %6s %10.2f" % ("sum of means", "", sum(sum(v) / len(v) for v in acc.values())))
