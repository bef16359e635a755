#!/bin/bash
# knob sweep on one box: current build first and last (drift check), each variant once per shape
mkdir -p gpurun_out; : > gpurun_out/ab5.log
run() { lib=$1; shift; echo -n "$lib :: $* :: " | tee -a gpurun_out/ab5.log
  MINCOB_LIBRARY=$lib timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],2), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab5.log; }
for lib in "" $(ls variants/*.so | grep -v timing) ""; do
  run "$lib" --steps 4 --warmup 3
  run "$lib" --pieces 5 --steps 4 --warmup 3
done
