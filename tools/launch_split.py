"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum csv): python tools/launch_split.py file.csv [last_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[1:]:
    try:
        t = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*(<[^(]*>)?\s*\(", r[ki])
    seq.append((m.group(1) if m else r[ki], t / 1e3))
if len(sys.argv) > 2:
    seq = seq[-int(sys.argv[2]):]
agg = OrderedDict()
for k, t in seq:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
for k, (c, t) in agg.items():
    print(f"{k:32s} x{c:<4d} {t:10.1f} us  ({t / c:9.1f} us each)")
