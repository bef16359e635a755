"""Experiment (needs a -DMINCOB_TIMING build, MINCOB_LIBRARY=...): when does the work queue of the persistent
optimize kernel run dry, and how fast do groups go idle after that?"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, '.')
from allocnet_b200 import api, synth
from allocnet_b200.params import default_params
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 8
KK = int(sys.argv[3]) if len(sys.argv) > 3 else 16
pb = synth.make_problems(B, N=NP, K=KK, S=3)
mb = api.MincoBatch(default_params(3), device=0); mb.set_problems(pb)
for rep in range(2):
    r = mb.optimize(pb.x0(), want_coeffs=False)
out = (C.c_ulonglong * 128)()
mb.L.mincob_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
assert mb.L.mincob_debug_counters(mb.h, out, 128) == 0
v = np.array(out[:], dtype=np.uint64)
t0, tq, te = int(v[3]), int(v[1]), int(v[2])
print(f"B={B}: kernel {1e-6*(te-t0):.1f} ms; queue ran dry at {1e-6*(tq-t0):.1f} ms; tail {1e-6*(te-tq):.1f} ms ({100*(te-tq)/(te-t0):.0f} %)")
bins = v[4:68].astype(np.int64)
print("groups going idle per ms after that:", bins[:int(1e-6*(te-tq))+2].tolist())
ev = r["evals"]; print("evals: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %d" % (ev.mean(), *np.percentile(ev, [50, 90, 99]), ev.max()))
