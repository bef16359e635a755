#!/bin/bash
# ordered replica accumulation (latency mapping bit-identical to the throughput mapping): parity suite + A/B
mkdir -p gpurun_out; : > gpurun_out/r2e.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e.log; tail -8 gpurun_out/pytest_gpu.log | tee -a gpurun_out/r2e.log
run() { envs=$1; shift; echo -n "$envs :: $* :: " | tee -a gpurun_out/r2e.log
  env $envs timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'], d['config']['mapping'].split()[0])" | tee -a gpurun_out/r2e.log; }
B=MINCOB_LIBRARY=$PWD/variants/a_base.so
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
run "$B" --pieces 5 --steps 5
run "A=1" --pieces 5 --steps 5
run "$B" --batch 1 --pieces 5 --steps 30 --warmup 5
run "A=1" --batch 1 --pieces 5 --steps 30 --warmup 5
run "$B" --batch 148 --steps 10 --warmup 3
run "A=1" --batch 148 --steps 10 --warmup 3
run "$B" --batch 148 --steps 10 --warmup 3 --mapping throughput
run "A=1" --batch 148 --steps 10 --warmup 3 --mapping throughput
run "$B" --batch 1776 --steps 5 --warmup 3 --mapping latency
run "A=1" --batch 1776 --steps 5 --warmup 3 --mapping latency
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
