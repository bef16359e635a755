#!/bin/bash
# what does a LONE warp per SM do?  (the launch tail / single-problem latency regime)
mkdir -p gpurun_out
for mp in throughput latency; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -s 3 -c 1 -f -o gpurun_out/lone_$mp python bench.py --batch 148 --pieces 8 --steps 1 --warmup 3 --no-cpu --no-e2e --mapping $mp > gpurun_out/lone_$mp.log 2>&1; echo "ncu $mp rc=$?"
  python tools/ncu_summary.py gpurun_out/lone_$mp.ncu-rep > gpurun_out/lone_${mp}_summary.txt 2>&1
  cat gpurun_out/lone_${mp}_summary.txt
done
python bench.py --batch 148 --pieces 8 --steps 10 --warmup 3 --no-cpu --no-e2e --mapping throughput | tail -1 | cut -c1-400
python bench.py --batch 148 --pieces 8 --steps 10 --warmup 3 --no-cpu --no-e2e --mapping latency | tail -1 | cut -c1-400
