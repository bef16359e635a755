"""Where does the end-to-end step spend its time?  (host-pointer C-ABI path of bench.py, one GPU)"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from allocnet_b200 import api, synth
from allocnet_b200.params import default_params
B, N, K, S = 65536, 8, 16, 3
pb = synth.make_problems(B, N=N, K=K, S=S)
mb = api.MincoBatch(default_params(S), device=0)
pin = api.pinned_empty
n = 4 * N - 3
h = synth.ProblemBatch(S, N, K, pin(pb.head.shape), pin(pb.tail.shape), pin(pb.hpolys.shape), pin(pb.hrows.shape, np.int32), pb.q0, pb.T0)
h.head[...] = pb.head; h.tail[...] = pb.tail; h.hpolys[...] = pb.hpolys; h.hrows[...] = pb.hrows
h_x0, h_x = pin((B, n)), pin((B, n)); h_x0[...] = pb.x0()
h_f, h_T = pin((B,)), pin((B, N))
h_st, h_it, h_ev = pin((B,), np.int32), pin((B,), np.int32), pin((B,), np.int32)
h_c = pin((B * N * 3 * 2 * S,))
for mode in ("sync", "async", "sync", "async"):
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); h_x[...] = h_x0
        t1 = time.perf_counter()
        (mb.set_problems if mode == "sync" else mb.set_problems_async)(h)
        t2 = time.perf_counter()
        mb.optimize_host_buffers(h_x, h_f, h_st, h_it, h_ev, h_c, h_T)
        t3 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2, mb.last_kernel_ms()[0]))
    a = np.array(ts[1:]).mean(axis=0) * [1e3, 1e3, 1e3, 1]
    print(f"{mode}: x copy {a[0]:.2f} ms, set_problems {a[1]:.2f} ms, optimize call {a[2]:.2f} ms (kernel {a[3]:.2f} ms), total {a[:3].sum():.2f} ms")
