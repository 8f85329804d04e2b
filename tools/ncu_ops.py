"""Opcode histogram (executed count, stall samples, shared-memory wavefronts) of one kernel from
`ncu --page source --csv` output.  usage: ncu_ops.py file.csv [full]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('Instructions Executed')].isdigit()]
ia, isrc, ist = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
iw, iwi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[ist]) for r in data)
print("total warp-inst", tot, "sass lines", len(data))
c, s, w, wi = Counter(), Counter(), Counter(), Counter()
for r in data:
    op = r[isrc].split()
    o = op[0] if not op[0].startswith('@') else op[1]
    k = o if len(sys.argv) > 2 else o.split('.')[0]
    c[k] += int(r[ia])
    s[k] += int(r[ist])
    w[k] += int(r[iw] or 0)
    wi[k] += int(r[iwi] or 0)
for o, n in c.most_common(24):
    print(f"{o:26s} {n/1e6:8.1f}M {100*n/tot:5.1f}%  samples {100*s[o]/max(ts,1):5.1f}%  smem wavefronts {w[o]/1e6:6.1f}M (ideal {wi[o]/1e6:6.1f}M)")
