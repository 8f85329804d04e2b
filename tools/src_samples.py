"""Where a kernel's warps spend their time: stall samples and executed instructions per stretch of SASS, from an ncu
source-page CSV (ncu -i X.ncu-rep --page source --csv > src.csv).  Usage: src_samples.py src.csv [bytes per bucket=0x200]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x200
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {n: k for k, n in enumerate(rows[hdr])}
data = [r for r in rows[hdr + 1:] if len(r) > 10 and r[0].startswith("0x")]
base = int(data[0][0], 16)
tot = sum(int(r[ix["# Samples"]]) for r in data)
toti = sum(int(r[ix["Instructions Executed"]]) for r in data)
stalls = [n for n in ix if n.startswith("stall_") and "Not Issued" not in n]
b = {}
for r in data:
    k = (int(r[0], 16) - base) // step
    e = b.setdefault(k, [0, 0, 0, {}])
    e[0] += int(r[ix["# Samples"]])
    e[1] += int(r[ix["Instructions Executed"]])
    e[2] += int(r[ix["Thread Instructions Executed"]])
    for s in stalls:
        e[3][s] = e[3].get(s, 0) + int(r[ix[s]] or 0)
print(f"samples {tot}, warp instructions {toti}")
for k in sorted(b):
    s, n, t, st = b[k]
    if s > tot * 0.01 or n > toti * 0.01:
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"{k * step:#07x} samples {100 * s / tot:5.1f}%  instructions {100 * n / toti:5.1f}%  lanes {t / max(n, 1):4.1f}  ",
              ", ".join(f"{a[6:]} {100 * v / max(s, 1):.0f}%" for a, v in top))
