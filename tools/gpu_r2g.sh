#!/bin/bash
# end-of-round records: launch tail of the v16 kernel (timing build), ncu captures of the fixed-time kernel and of a truly lone problem in both mappings
mkdir -p gpurun_out; : > gpurun_out/r2g.log
for args in "65536 8 16" "65536 5 16" "65536 5 50"; do MINCOB_LIBRARY=$PWD/variants/timing.so timeout 300 python tools/tail_probe.py $args 2>&1 | tail -3 | tee -a gpurun_out/r2g.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -s 3 -c 1 -f -o gpurun_out/prof_fixed_time python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-check --no-pipeline --freeze-times > gpurun_out/ncu_fixed.log 2>&1; echo "ncu fixed rc=$?" | tee -a gpurun_out/r2g.log
python tools/ncu_summary.py gpurun_out/prof_fixed_time.ncu-rep > gpurun_out/fixed_time_summary.txt 2>&1; head -40 gpurun_out/fixed_time_summary.txt | tee -a gpurun_out/r2g.log
for mp in throughput latency; do
  timeout 600 ncu --set full --clock-control none -k regex:optimize_kernel -s 3 -c 1 -f -o gpurun_out/one_$mp python bench.py --batch 1 --pieces 8 --steps 1 --warmup 3 --no-cpu --no-e2e --no-check --no-pipeline --mapping $mp > gpurun_out/one_$mp.log 2>&1; echo "ncu one $mp rc=$?" | tee -a gpurun_out/r2g.log
  python tools/ncu_summary.py gpurun_out/one_$mp.ncu-rep > gpurun_out/one_${mp}_summary.txt 2>&1; head -36 gpurun_out/one_${mp}_summary.txt | tee -a gpurun_out/r2g.log
done
