"""Hot straight-line SASS regions of one kernel (ncu --page source --csv): runs of instructions with equal executed count."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia = hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ia].isdigit()]
half = len(data) // 2 if len(sys.argv) > 2 else len(data)
data = data[:half]
isrc, it, ist = hdr.index('Source'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[ist]) for r in data)
groups, cur = [], None
for i, r in enumerate(data):
    n = int(r[ia])
    if cur and abs(n - cur['n']) <= 0.02 * max(n, cur['n'], 1):
        cur['end'] = i; cur['sum'] += n; cur['samples'] += int(r[ist])
    else:
        cur = {'start': i, 'end': i, 'n': n, 'sum': n, 'samples': int(r[ist])}
        groups.append(cur)
print("total", tot / 1e6, "M warp-inst,", len(data), "lines")
for g in sorted(groups, key=lambda g: -g['sum'])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    r = data[g['start']]
    print(f"[{g['start']:4d}-{g['end']:4d}] len {g['end']-g['start']+1:3d} exec/line {g['n']/1e6:6.2f}M sum {g['sum']/1e6:6.1f}M "
          f"({100*g['sum']/tot:4.1f}%) thr {r[it][:4]} samples {100*g['samples']/max(ts,1):4.1f}%  {r[isrc][:46]}")
