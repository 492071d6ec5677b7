#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [steps]"""
import collections
import csv
import re
import sys

rows = []
for line in open(sys.argv[1]):
    if line.startswith('"ID"') or (line.startswith('"') and line[1].isdigit()):
        rows.append(next(csv.reader([line])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
gi = hdr.index("Grid Size")
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    name = re.sub(r"^void (ct2b200::)?(\(anonymous namespace\)::|<unnamed>::)?", "", r[ki])
    name = re.sub(r"\(.*", "", name)
    t = float(r[vi].replace(",", "")) / 1000
    a = agg.setdefault((name[:90], r[gi]), [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
print("%d launches, %.1f us per step (sum of kernel durations, %d step(s))" % (len(rows) // steps, tot / steps, steps))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print("%9.1f us/step %5.1f%% %4d x %8.2f us  %s %s" % (t / steps, 100 * t / tot, n // steps, t / n, k[0], k[1]))
