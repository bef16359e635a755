#!/bin/bash
# A/B on one box: the headline config, current build vs each variants/*.so (alternating, 2 rounds)
mkdir -p gpurun_out; : > gpurun_out/ab2.log
run() { lib=$1; shift; echo -n "$lib :: $* :: " | tee -a gpurun_out/ab2.log
  MINCOB_LIBRARY=$lib timeout 600 python bench.py "$@" --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],2), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab2.log; }
for rep in 1 2; do for lib in "" $(ls variants/*.so 2>/dev/null | grep -v timing); do
  run "$lib" --steps 4 --warmup 3
  run "$lib" --pieces 5 --steps 4 --warmup 3
done; done
