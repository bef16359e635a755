#!/usr/bin/env python
"""One-screen summary of an .ncu-rep (first kernel): python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, io, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vals = rows[2 + idx]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum"]
for k in keys:
    if k in d: print(f"{k:75s} {d[k][0]} {d[k][1]}")
print("-- stalls per issue (warps):")
st = [(float(v[0]), h) for h, v in d.items() if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h) and v[0]]
for v, h in sorted(st, reverse=True)[:8]:
    print(f"   {h.split('stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
