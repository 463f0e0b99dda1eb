#!/usr/bin/env python
"""`ncu --page raw --csv` export -> compact JSON summary (one entry per launch) for profiles/.
usage: ncu_summary.py raw.csv out.json "note" """
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
out = {"note": sys.argv[3] if len(sys.argv) > 3 else "", "kernels": []}
for r in rows[2:]:
    k = {}
    for w in keep:
        if w in hdr:
            i = hdr.index(w); k[w] = (r[i] + " " + units[i]).strip()
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try:
                v = float(r[i])
                if v > 0.05: st[h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')] = round(v, 2)
            except ValueError: pass
    k["stalls_per_issue"] = dict(sorted(st.items(), key=lambda x: -x[1]))
    out["kernels"].append(k)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"{len(out['kernels'])} launches -> {sys.argv[2]}")
