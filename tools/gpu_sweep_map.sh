#!/bin/bash
# small-batch sweep: which mapping is faster at which batch size (for the MINCOB_MAP_AUTO rule)
mkdir -p gpurun_out; : > gpurun_out/sweep_map.log
for n in 5 8; do for b in 1 8 32 64 148 296 592 1184 1776 3552; do for m in latency throughput; do
echo -n "N=$n B=$b $m :: " | tee -a gpurun_out/sweep_map.log
timeout 300 python bench.py --batch $b --pieces $n --steps 10 --warmup 3 --mapping $m --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), round(d['mean_evals_per_traj'],1), d['config']['mapping'].split()[0])" | tee -a gpurun_out/sweep_map.log
done; done; done
