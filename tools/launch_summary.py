"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (and list every launch with -v)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot, seq = collections.OrderedDict(), 0.0, []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("jb::", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[ui]]
    a = agg.setdefault(n, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
    seq.append((n, v))
for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{v:8.3f} ms {c:4d}  {100 * v / tot:5.1f}%  {n[:100]}")
print(f"{tot:8.3f} ms total, {len(seq)} launches")
if "-v" in sys.argv:
    for i, (n, v) in enumerate(seq):
        print(f"{i:4d} {v * 1e3:9.1f} us  {n[:90]}")
