"""Turns ncu exports into the committed summaries under profiles/.

    python tools/profile_summary.py <round-tag> <launches.csv> <raw.csv> [workload]

  launches.csv : ncu --metrics gpu__time_duration.sum --csv --log-file ... (every launch, cold cache, serialised)
  raw.csv      : ncu -i prof.ncu-rep --page raw --csv                      (one --set full capture per kernel;
                 several files separated by commas are merged)
Writes profiles/<tag>_launches.md, profiles/<tag>_kernels.md and profiles/traffic.json (dram bytes per launch,
read by bench.py for roofline.traffic).
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, raw_csv = sys.argv[1:4]
workload = sys.argv[4] if len(sys.argv) > 4 else "text-1G"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)


def short(name):
    import re

    m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*(<[^(]*>)?\s*\(", name)
    if name.startswith("void at::") or name.startswith("at::"):
        return "at::" + (m.group(1) if m else name)
    return m.group(1) if m else name


# ---- launch list: per-kernel totals over ONE step (the last occurrence of each kernel sequence)
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(short(r[ki]), float(r[vi]) / 1e3) for r in rows[1:] if r[vi].replace(".", "").isdigit()]
ours = [(k, t) for k, t in seq if not k.startswith("at::")]
# last step = from the last histogram_kernel launch to the end
last = max(i for i, (k, _) in enumerate(ours) if k == "histogram_kernel")
step = ours[last:]
tot = sum(t for _, t in step)
agg = OrderedDict()
for k, t in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: launch list of one bench step ({workload})\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py`; times are cold-cache and\n"
            "serialised, so compare SHARES, not absolutes.  One step = encode (histogram, pack) + decode.\n\n")
    f.write("| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
    for k, (c, t) in agg.items():
        f.write(f"| {k} | {c} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    f.write(f"| **total** | {len(step)} | {tot:.1f} | 100% |\n")

# ---- full captures
# raw_csv may be several files separated by commas (captures of different kernels of the same step): each row keeps
# the units of its own file
hdr, data = None, []
for path in raw_csv.split(","):
    rows = list(csv.reader(open(path)))
    if hdr is None:
        hdr = rows[0]
    assert rows[0] == hdr, "captures were exported with different metric sets"
    data += [(r, rows[1]) for r in rows[2:]]
col = {h: i for i, h in enumerate(hdr)}
keys = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy % (of active cycles)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy % (of active cycles)"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.max", "SM elapsed cycles (max)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]
stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
traffic = {}
try:
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass
traffic.pop(workload, None)  # first capture of a kernel wins within one run (later launches of it are check rounds)
with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` per kernel ({workload})\n\n")
    for r, units in data:
        name = short(r[col["Kernel Name"]])
        f.write(f"## {name}\n\n| metric | value |\n|---|---|\n")
        for k, label in keys:
            if k in col:
                f.write(f"| {label} | {r[col[k]]} {units[col[k]]} |\n")
        top = sorted(((float(r[col[h]]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stalls),
                     reverse=True)[:5]
        f.write("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in top) + " |\n\n")

        def to_bytes(k):
            v, u = float(r[col[k]]), units[col[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        traffic.setdefault(workload, {}).setdefault(name, int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")))
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("wrote profiles/", tag)
