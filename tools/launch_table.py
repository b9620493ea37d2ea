#!/usr/bin/env python
"""Developer tool: print (id, kernel, grid, metrics...) rows of an `ncu --csv --metrics ...` launch list."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, mi, vi, gi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Grid Size")
d = OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) > vi:
        d.setdefault((int(r[0]), r[ki][:44], r[gi]), {})[r[mi].split("__")[-1]] = r[vi]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for k, v in list(d.items())[-last:]:
    print(k[0], k[1].ljust(44), k[2].ljust(16), "  ".join(f"{n.split('.')[0]}={float(x.replace(',', '')):.4g}" for n, x in v.items()))
