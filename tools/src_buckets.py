"""Where a kernel's issue slots go: executed warp instructions per 1 KiB of SASS from an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x400
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {n: k for k, n in enumerate(rows[hdr])}
data = [r for r in rows[hdr + 1:] if len(r) > 10 and r[0].startswith("0x")]
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
base = int(data[0][0], 16)
print("warp instructions", tot, "SASS instructions", len(data))
b = {}
for r in data:
    k = (int(r[0], 16) - base) // step
    e = b.setdefault(k, [0, 0, 0, 0])
    e[0] += int(r[ix["Instructions Executed"]])
    e[1] += int(r[ix["Thread Instructions Executed"]])
    e[2] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    e[3] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
for k in sorted(b):
    n, t, w, wi = b[k]
    if n > tot * 0.004:
        print(f"{k * step:#07x}  {n:>12}  {100 * n / tot:5.1f}%  lanes {t / max(n, 1):4.1f}  smem wavefronts {w:>11} (ideal {wi})")
