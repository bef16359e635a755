#!/bin/bash
# Quick iteration call: parity tests, a short bench, optional ncu capture (NCU=1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -s 3 -c 1 -f -o gpurun_out/prof_optimize python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
