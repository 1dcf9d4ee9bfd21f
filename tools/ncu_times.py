#!/usr/bin/env python
"""Condense an `ncu --csv` metrics log into one line per launch:  python tools/ncu_times.py gpurun_out/x.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid], {"k": r[ik][:70]})[r[im].replace("smsp__", "").replace(".avg.pct_of_peak_sustained_active", "%")] = r[iv]
for k, v in d.items():
    print(k, v)
