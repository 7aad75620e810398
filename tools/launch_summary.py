#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: count and mean duration per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h = r; start = i + 1; break
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi: continue
    n = r[ki].replace('<unnamed>::', '')[:48]
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += float(r[vi].replace(',', ''))
for n, (c, t) in agg.items():
    print(f"{c:4d} {t / c / 1000:10.1f} us  {n}")
