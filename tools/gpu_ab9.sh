#!/bin/bash
# block-size variants for 8 pieces against the default build: value and the two-batches-in-flight number
mkdir -p gpurun_out; : > gpurun_out/ab9.log
run() { lib=$1; shift; echo -n "$lib :: $* :: " | tee -a gpurun_out/ab9.log
  MINCOB_LIBRARY=$PWD/$lib timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), 'pipelined', round((d.get('pipelined') or {}).get('value',0)), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab9.log; }
for i in 1 2 3; do for lib in variants/a_base.so variants/t_threads32.so variants/t_threads64.so; do run $lib --steps 4 --warmup 3; done; done
