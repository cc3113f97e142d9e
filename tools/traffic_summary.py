"""Summarise `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` per kernel launch:
total / per-launch DRAM traffic of the selected kernel (used for bench.py's roofline.traffic)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ii, ki, mi, vi, ui = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = per.setdefault(r[ii], {'name': r[ki].split('(')[0]})
    d[r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
n = len(per)
rd = sum(d.get('dram__bytes_read.sum', 0) for d in per.values())
wr = sum(d.get('dram__bytes_write.sum', 0) for d in per.values())
ms = sum(d.get('gpu__time_duration.sum', 0) for d in per.values())
big = [d for d in per.values() if d.get('gpu__time_duration.sum', 0) > 1.0]
print("launches %d  total read %.3f GB  write %.3f GB  time %.2f ms" % (n, rd * 1e-9, wr * 1e-9, ms))
print("per launch (all): %.3f GB   per launch (outer updates > 1 ms, n=%d): %.3f GB, %.2f ms" % (
    (rd + wr) / max(n, 1) * 1e-9, len(big), sum(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in big) / max(len(big), 1) * 1e-9,
    sum(d['gpu__time_duration.sum'] for d in big) / max(len(big), 1)))
for k, d in list(per.items())[:400]:
    if d.get('gpu__time_duration.sum', 0) > 1.0:
        print("  id %s %s %.2f ms read %.3f GB write %.3f GB" % (k, d['name'], d['gpu__time_duration.sum'], d.get('dram__bytes_read.sum', 0) * 1e-9, d.get('dram__bytes_write.sum', 0) * 1e-9))
