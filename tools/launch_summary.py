"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step per kernel)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u, 1.0)
    name = r[ki].split('(')[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-34s %5s %12s %7s %10s" % ("kernel", "n", "total ms", "share", "avg ms"))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-34s %5d %12.3f %6.1f%% %10.4f" % (k, n, t, 100 * t / tot, t / n))
print("%-34s %5d %12.3f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
