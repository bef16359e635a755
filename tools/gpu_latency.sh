#!/bin/bash
# Small-batch latency of one optimize call (BASELINE.json configs[0]: the online planner solves ONE 5-piece problem per plan).
mkdir -p gpurun_out; : > gpurun_out/latency.log
for b in 1 16 256 4096; do
  echo "== batch $b x 5 pieces" | tee -a gpurun_out/latency.log
  timeout 300 python bench.py --batch $b --pieces 5 --steps 20 --warmup 5 --cpu-sample $(( b < 64 ? 64 : (b < 1024 ? b : 1024) )) 2>&1 | tail -1 | tee -a gpurun_out/latency.log
done
