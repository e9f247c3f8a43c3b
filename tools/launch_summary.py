"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [last_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("unnamed>::", ""), float(r[vi].replace(",", "")))
       for r in rows[h + 1:] if len(r) > vi and r[vi]]
if len(sys.argv) > 2:
    seq = seq[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
print("%d launches, %.3f ms of kernel time (serialised, cold)" % (len(seq), sum(v for _, v in seq) / 1e6))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%9.3f ms %5d  %s" % (t / 1e6, n, k[:110]))
