#!/usr/bin/env python
"""One line per profiled kernel from an `ncu --page raw --csv` dump: the metrics the roofline discussion uses."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
cols = [('Kernel Name', 'kernel', 26), ('gpu__time_duration.sum', 'ms', 8), ('dram__bytes_read.sum', 'rd', 9), ('dram__bytes_write.sum', 'wr', 9),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 6),
        ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex%', 6), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 6),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6), ('launch__registers_per_thread', 'regs', 5),
        ('smsp__inst_executed.sum', 'winst', 11), ('l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed', 'texwf%', 6),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'lanes', 5)]
print(' '.join(f"{n:>{w}s}" for _, n, w in cols))
units = rows[1]
for r in rows[2:]:
    out = []
    for c, n, w in cols:
        v = r[h.index(c)] if c in h else ''
        if c == 'Kernel Name':
            v = v.replace('<unnamed>::', '').split('(')[0]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        out.append(f"{v[:w]:>{w}s}")
    print(' '.join(out))
print('units:', {n: units[h.index(c)] for c, n, _ in cols if c in h and c != 'Kernel Name'})
