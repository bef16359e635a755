#!/bin/bash
# round 2 check A: full GPU parity suite, then throughput / latency benches of the shapes the verdict names
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
: > gpurun_out/r2a_bench.log
run() { echo "== $*" | tee -a gpurun_out/r2a_bench.log; timeout 600 python bench.py "$@" --no-cpu 2>&1 | tail -1 | tee -a gpurun_out/r2a_bench.log; }
run --steps 5 --warmup 3
run --pieces 5 --steps 5 --warmup 3
run --pieces 5 --K 50 --steps 5 --warmup 3
run --batch 1 --pieces 5 --steps 30 --warmup 5 --no-e2e
run --batch 64 --pieces 5 --steps 30 --warmup 5 --no-e2e
run --batch 1024 --pieces 8 --steps 20 --warmup 5 --no-e2e
