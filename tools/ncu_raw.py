#!/usr/bin/env python
"""Key metrics per kernel from an `ncu --page raw --csv` export:  python tools/ncu_raw.py file.raw.csv [extra metric substrings]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max']
want += [h for h in hdr if 'issue_stalled' in h and 'per_issue_active.ratio' in h and 'not_issued' not in h]
for a in sys.argv[2:]:
    want += [h for h in hdr if a in h and h not in want]
names = [r[idx['Kernel Name']][:100] for r in rows[2:]]
print("kernels:")
for i, nme in enumerate(names):
    print(f"  [{i}] {nme}")
for w in want:
    if w in idx:
        vals = [r[idx[w]][:14] for r in rows[2:]]
        if 'issue_stalled' in w and all(float(v or 0) < 0.3 for v in vals):
            continue
        print(f"{w[:88]:88s} " + " ".join(f"{v:>14s}" for v in vals) + f"  {units[idx[w]]}")
