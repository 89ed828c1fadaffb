#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (reads `ncu -i rep --page raw --csv`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-' * 60)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in stalls), reverse=True)
    print('stalls:', ', '.join(f"{n}={v:.2f}" for v, n in st[:7]))
