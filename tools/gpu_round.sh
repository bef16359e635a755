#!/bin/bash
# Round record: GPU parity tests, both bench arms, ncu launch list + full capture of optimize_kernel, sanitizer.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; cat gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-check > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:optimize_kernel -s 3 -c 1 -f -o gpurun_out/prof_optimize python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
if [ -n "$SANITIZE" ]; then bash tools/gpu_sanity.sh; fi
# the other configs of BASELINE.json and the reference-sized shapes, one line each
: > gpurun_out/configs.log
cfg() { echo "== $*" >> gpurun_out/configs.log; timeout 600 python bench.py "$@" --no-cpu 2>/dev/null | tail -1 >> gpurun_out/configs.log; }
cfg --config 2 --steps 10
cfg --config 2 --steps 10 --mapping latency
cfg --config 4 --steps 3
cfg --pieces 5 --steps 5
cfg --pieces 5 --K 50 --steps 3
cfg --pieces 5 --steps 5 --freeze-times
cfg --steps 4 --freeze-times
cfg --pieces 16 --steps 3 --freeze-times
cfg --pieces 5 --K 50 --steps 3 --freeze-times
cfg --pieces 16 --steps 3
cfg --batch 1 --pieces 5 --steps 30 --warmup 5 --no-e2e
cfg --batch 1 --pieces 5 --steps 30 --warmup 5 --no-e2e --mapping throughput
cfg --batch 64 --pieces 5 --steps 30 --warmup 5 --no-e2e
cfg --batch 1 --steps 30 --warmup 5 --no-e2e
cfg --batch 1 --steps 30 --warmup 5 --no-e2e --mapping throughput
cfg --S 4 --steps 3
python - <<'PY'
import json
for line in open("gpurun_out/configs.log"):
    if line.startswith("=="): print(line.strip()); continue
    try:
        d = json.loads(line)
        print("   value %.0f traj/s  ms/step %.3f  evals/traj %.0f  mapping %s  parity %s  e2e %s" % (d["value"], d["ms_per_step"], d["mean_evals_per_traj"], d["config"]["mapping"].split()[0], (d.get("parity_check") or {}).get("max_rel"), (d.get("e2e") or {}).get("value")))
    except Exception as e: print("   ?", e)
PY
