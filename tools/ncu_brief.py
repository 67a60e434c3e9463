#!/usr/bin/env python3
"""Brief of an ncu report: per kernel the few metrics we steer by, and the stall mix from the source page."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
stalls = [n for n in h if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    if len(r) < len(h): continue
    print("##", r[h.index("Kernel Name")][:100])
    for n in want:
        if n in h: print("  %-78s %s %s" % (n, r[h.index(n)], u[h.index(n)]))
    st = sorted(((float(r[h.index(n)] or 0), n.split('stalled_')[1].split('_per_issue')[0]) for n in stalls), reverse=True)
    print("  stalls per issue:", ", ".join("%s %.2f" % (b, a) for a, b in st[:8]))
