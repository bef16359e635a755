#!/bin/bash
# several variant libraries against variants/a_base.so, headline shape only, interleaved
mkdir -p gpurun_out; : > gpurun_out/ab8.log
run() { lib=$1; shift; echo -n "$lib :: $* :: " | tee -a gpurun_out/ab8.log
  MINCOB_LIBRARY=$PWD/$lib timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab8.log; }
for i in 1 2 3; do for lib in variants/a_base.so $LIBS; do run $lib --steps 4 --warmup 3; done; done
