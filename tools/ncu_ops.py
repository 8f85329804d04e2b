"""Opcode histogram (weighted by executed count) of one kernel from `ncu --page source --csv` output."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ia, isrc, ist = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data)
print("total warp-inst", tot, "sass lines", len(data))
c, s = Counter(), Counter()
for r in data:
    op = r[isrc].split()
    o = op[0] if not op[0].startswith('@') else op[1]
    k = o if len(sys.argv) > 2 else o.split('.')[0]
    c[k] += int(r[ia])
    s[k] += int(r[ist])
ts = sum(s.values())
for o, n in c.most_common(30):
    print(f"{o:28s} {n/1e6:8.1f}M {100*n/tot:5.1f}%  samples {100*s[o]/max(ts,1):5.1f}%")
