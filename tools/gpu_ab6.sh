#!/bin/bash
# A/B: variants/a_base.so (previous commit) against the working tree; bit-equality first, then speed (alternating)
mkdir -p gpurun_out; : > gpurun_out/ab6.log
python tools/diff_variants.py variants/a_base.so allocnet_b200/libmincob.so 2048 8 2>&1 | tail -4 | tee -a gpurun_out/ab6.log
python tools/diff_variants.py variants/a_base.so allocnet_b200/libmincob.so 1200 5 2>&1 | tail -2 | tee -a gpurun_out/ab6.log
run() { envs=$1; shift; echo -n "$envs :: $* :: " | tee -a gpurun_out/ab6.log
  env $envs timeout 600 python bench.py "$@" --no-cpu --no-e2e --no-check --no-pipeline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['kernel_ms'],3), round(d['mean_evals_per_traj'],1), d['clocks']['sm_mhz'], d['config']['mapping'].split()[0])" | tee -a gpurun_out/ab6.log; }
B=MINCOB_LIBRARY=$PWD/variants/a_base.so
for i in 1 2 3; do
run "$B" --steps 4 --warmup 3
run "A=1" --steps 4 --warmup 3
done
for i in 1 2; do
run "$B" --pieces 5 --steps 5
run "A=1" --pieces 5 --steps 5
done
run "$B" --pieces 16 --steps 3
run "A=1" --pieces 16 --steps 3
${EXTRA_CMD}
