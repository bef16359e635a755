#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --config 4 --steps 3 --no-cpu > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; echo "config4 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config4.json')); print(d['value'], d['ms_per_step'], d['warm_start'], d.get('parity_check'), d['e2e']['value'])"; tail -3 gpurun_out/bench_config4.err
timeout 900 python bench.py --config 2 --steps 5 --no-cpu > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "config2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2.json')); print(d['value'], d['ms_per_step'], d['status_hist'], d.get('parity_check'), d['config']['mapping'])"
