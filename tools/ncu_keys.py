"""Key metrics + stall reasons of every kernel in an `ncu --page raw --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(k, rows[1][i], [r[i][:24] for r in rows[2:]])
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        v = [float(r[i]) for r in rows[2:]]
        if max(v) > 0.4:
            print(h.split('issue_stalled_')[1], [round(x, 2) for x in v])
