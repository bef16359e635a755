#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/r2f.log
timeout 900 python -m pytest tests -m gpu -q -s -k "near_optimum or latency_mapping or fixed_time" 2>&1 | grep -v "^$" | tail -8 | tee -a gpurun_out/r2f.log
run() { envs=$1; shift; echo -n "$envs :: $* :: " | tee -a gpurun_out/r2f.log
  env $envs timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'], d['config']['mapping'].split()[0])" | tee -a gpurun_out/r2f.log; }
B=MINCOB_LIBRARY=$PWD/variants/a_base.so
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
run "$B" --pieces 5 --steps 5
run "A=1" --pieces 5 --steps 5
for n in 5 8; do for m in latency throughput; do
run "$B" --batch 1 --pieces $n --steps 30 --warmup 5 --mapping $m
run "A=1" --batch 1 --pieces $n --steps 30 --warmup 5 --mapping $m
done; done
run "$B" --batch 1776 --steps 5 --warmup 3 --mapping latency
run "A=1" --batch 1776 --steps 5 --warmup 3 --mapping latency
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
