#!/usr/bin/env python
"""Per-launch table (ms, DRAM MB read / written) from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ki, mi, vi, ii, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID'), h.index('Metric Unit')
d = {}
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if 'byte' in u: v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[u]
    else: v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'second': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(u, 1.0)
    d.setdefault((int(r[ii]), r[ki][:48]), {})[r[mi]] = v
for (i, k), v in sorted(d.items()):
    print('%3d %-48s %8.3f ms  rd %9.1f MB  wr %9.1f MB' % (i, k, v.get('gpu__time_duration.sum', 0), v.get('dram__bytes_read.sum', 0), v.get('dram__bytes_write.sum', 0)))
