#!/bin/bash
# A/B on one box: variants/*.so against the current build, same bench lines
mkdir -p gpurun_out; : > gpurun_out/ab.log
run() { lib=$1; shift; echo "== $lib :: $*" | tee -a gpurun_out/ab.log
  MINCOB_LIBRARY=$lib timeout 600 python bench.py "$@" --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','evals_per_s','mean_evals_per_traj')}, d['roofline']['kernel_ms'], d.get('clocks'))" | tee -a gpurun_out/ab.log; }
for lib in "" $(ls variants/*.so 2>/dev/null); do
  run "$lib" --steps 5 --warmup 3
  run "$lib" --pieces 5 --steps 5 --warmup 3
  run "$lib" --pieces 5 --K 50 --steps 3 --warmup 3
  run "$lib" --batch 1 --pieces 5 --steps 30 --warmup 5
  run "$lib" --pieces 16 --batch 32768 --steps 3 --warmup 3
done
