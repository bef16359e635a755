#!/bin/bash
# A/B of two libraries: $LIBA against $LIBB (bit-equality, then alternating speed runs)
mkdir -p gpurun_out; : > gpurun_out/ab7.log
LIBA=${LIBA:-allocnet_b200/libmincob.so}; LIBB=${LIBB:-variants/a_base.so}
python tools/diff_variants.py $LIBB $LIBA 2048 8 2>&1 | tail -3 | tee -a gpurun_out/ab7.log
python tools/diff_variants.py $LIBB $LIBA 1200 5 2>&1 | tail -1 | tee -a gpurun_out/ab7.log
run() { lib=$1; shift; echo -n "$lib :: $* :: " | tee -a gpurun_out/ab7.log
  MINCOB_LIBRARY=$PWD/$lib timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'], d['config']['mapping'].split()[0])" | tee -a gpurun_out/ab7.log; }
for i in 1 2 3; do run $LIBB --steps 4 --warmup 3; run $LIBA --steps 4 --warmup 3; done
for i in 1 2; do run $LIBB --pieces 5 --steps 5; run $LIBA --pieces 5 --steps 5; done
run $LIBB --pieces 16 --steps 3; run $LIBA --pieces 16 --steps 3
run $LIBB --batch 1 --pieces 5 --steps 30 --warmup 5; run $LIBA --batch 1 --pieces 5 --steps 30 --warmup 5
run $LIBB --batch 1776 --steps 5 --mapping latency; run $LIBA --batch 1776 --steps 5 --mapping latency
