#!/bin/bash
# A/B of kernel builds: variants/*.so, same short bench each.
mkdir -p gpurun_out
for so in variants/*.so; do
  echo "== $so"
  MINCOB_LIBRARY=$PWD/$so timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 ${BENCH_ARGS} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(json.dumps({k:d[k] for k in ('value','evals_per_s','ms_per_step','mean_evals_per_traj','ok_fraction')}), d['clocks'])
" | tee -a gpurun_out/variants.log
done
