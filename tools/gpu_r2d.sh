#!/bin/bash
# fixed-time kernel (FRZ): GPU parity suite, then A/B against the previous build and against the generic kernel
mkdir -p gpurun_out; : > gpurun_out/r2d.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d.log; tail -5 gpurun_out/pytest_gpu.log | tee -a gpurun_out/r2d.log
run() { envs=$1; shift; echo -n "$envs :: $* :: " | tee -a gpurun_out/r2d.log
  env $envs timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],2), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'])" | tee -a gpurun_out/r2d.log; }
B=MINCOB_LIBRARY=$PWD/variants/a_base.so
for rep in 1 2; do
run "$B" --pieces 5 --steps 5 --freeze-times
run "A=1" --pieces 5 --steps 5 --freeze-times
run "MINCOB_NO_FRZ=1" --pieces 5 --steps 5 --freeze-times
run "$B" --steps 4 --freeze-times
run "A=1" --steps 4 --freeze-times
done
run "A=1" --pieces 16 --steps 3 --freeze-times
run "MINCOB_NO_FRZ=1" --pieces 16 --steps 3 --freeze-times
run "A=1" --batch 1 --pieces 5 --steps 30 --warmup 5 --freeze-times
run "MINCOB_NO_FRZ=1" --batch 1 --pieces 5 --steps 30 --warmup 5 --freeze-times
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
