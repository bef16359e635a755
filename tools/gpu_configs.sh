#!/bin/bash
# BASELINE.json configs on one GPU (device-resident value + e2e), for the result table.
mkdir -p gpurun_out; : > gpurun_out/configs.log
run() { echo "== $*" | tee -a gpurun_out/configs.log; timeout 900 python bench.py --steps 3 --warmup 3 "$@" 2>&1 | tail -1 | tee -a gpurun_out/configs.log; }
run --batch 4096 --K 0 --cpu-sample 512                       # config 2: energy-only
run --batch 65536 --pieces 5 --cpu-sample 2048                 # reference ModelMaxSeg
run --batch 65536 --pieces 16 --cpu-sample 512                 # config 4 shape (trapezoid start, no net)
run --batch 65536 --S 4 --cpu-sample 512                       # MINCO_S4NU
run --batch 65536 --mem-size 16 --no-cpu
run --batch 65536 --mem-size 32 --no-cpu
