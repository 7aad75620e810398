#!/usr/bin/env python
"""Print selected metrics of one kernel from an `ncu --page raw --csv` dump.
usage: ncu_metrics.py raw.csv <kernel substring> [occurrence] <metric substring>..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
k = sys.argv[2]
occ = 0
args = sys.argv[3:]
if args and args[0].isdigit():
    occ = int(args[0]); args = args[1:]
n = -1
for r in rows[2:]:
    if k in r[h.index('Kernel Name')]:
        n += 1
        if n != occ:
            continue
        for i, c in enumerate(h):
            if any(p in c for p in args):
                print(f"{c:95s} {r[i]:>16s} {rows[1][i]}")
        break
