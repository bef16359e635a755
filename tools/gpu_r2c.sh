#!/bin/bash
# round 2, session 3: hybrid staging + fixed-time kernel: parity suite, then A/B against the previous build (variants/a_base.so)
mkdir -p gpurun_out; : > gpurun_out/r2c.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c.log; tail -5 gpurun_out/pytest_gpu.log | tee -a gpurun_out/r2c.log
run() { envs=$1; shift; echo -n "$envs :: $* :: " | tee -a gpurun_out/r2c.log
  env $envs timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],2), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'])" | tee -a gpurun_out/r2c.log; }
B=MINCOB_LIBRARY=$PWD/variants/a_base.so
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
run "MINCOB_LIBRARY=$PWD/variants/b_fastmem0.so" --steps 4 --warmup 3
run "$B" --pieces 5 --K 50 --steps 3
for ks in 0 16 24 32 40; do run "MINCOB_KS=$ks" --pieces 5 --K 50 --steps 3; done
run "$B" --pieces 8 --K 40 --steps 3
for ks in 0 12 20 28; do run "MINCOB_KS=$ks" --pieces 8 --K 40 --steps 3; done
run "$B" --pieces 5 --steps 5 --freeze-times
run "A=1" --pieces 5 --steps 5 --freeze-times
run "MINCOB_NO_FRZ=1" --pieces 5 --steps 5 --freeze-times
run "$B" --steps 4 --freeze-times
run "A=1" --steps 4 --freeze-times
run "$B" --pieces 5 --K 50 --steps 3 --freeze-times
run "A=1" --pieces 5 --K 50 --steps 3 --freeze-times
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
