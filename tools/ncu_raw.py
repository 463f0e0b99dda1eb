#!/usr/bin/env python
"""Key metrics of an `ncu --page raw --csv` export, one block per kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"{w} = {r[i]} {units[i]}")
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try: st[h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')] = float(r[i])
            except ValueError: pass
    print('stalls/issue', {k: round(v, 2) for k, v in sorted(st.items(), key=lambda x: -x[1]) if v > 0.05})
