#!/bin/bash
# smoke(), compute-sanitizer (memcheck + racecheck) on a small optimize/evaluate, full GPU tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from allocnet_b200 import api, synth
from allocnet_b200.params import default_params
from allocnet_b200 import params as P
for S, N, K, B, mp, fl in ((3, 8, 16, 96, P.MAP_THROUGHPUT, 0), (3, 8, 16, 9, P.MAP_LATENCY, 0), (3, 5, 50, 40, P.MAP_THROUGHPUT, 0), (3, 5, 16, 13, P.MAP_LATENCY, 0),
                       (3, 3, 7, 31, P.MAP_THROUGHPUT, 0), (4, 8, 16, 40, P.MAP_THROUGHPUT, 0), (4, 5, 16, 7, P.MAP_LATENCY, 0), (3, 16, 16, 24, P.MAP_LATENCY, 0), (3, 1, 4, 8, P.MAP_AUTO, 0),
                       # the fixed-time kernel (FRZ instantiation), both mappings
                       (3, 8, 16, 70, P.MAP_THROUGHPUT, P.FLAG_FREEZE_TIMES), (3, 5, 16, 11, P.MAP_LATENCY, P.FLAG_FREEZE_TIMES), (4, 8, 16, 20, P.MAP_THROUGHPUT, P.FLAG_FREEZE_TIMES)):
    prm = default_params(S, max_iterations=12, mapping=mp, flags=fl)
    pb = synth.make_problems(B, N=N, K=K, S=S, ragged_rows=(K == 50))
    mb = api.MincoBatch(prm, device=0); mb.set_problems(pb)
    f, g = mb.evaluate(pb.x0()); r = mb.optimize(pb.x0())
    q = pb.q0; out = mb.minco_forward(pb.head, pb.tail, q, pb.T0)
    rep = mb.check_feasibility(r["coeffs"], r["T"], samples=16); rates = mb.max_rates(r["coeffs"], r["T"])
    print(S, N, K, B, fl, float(f[0]), int(r["evals"].sum()))
    mb.close()
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
