"""Diagnostic (GPU box): actual error of the device MINCO path for S = 3 / 4 against the banded oracle and against
the numpy statement of the device algebra (oracle/reduced_proto.py), per piece count.  Test infrastructure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from allocnet_b200 import api
from allocnet_b200.params import default_params
from oracle.pyoracle import Oracle
from oracle.reduced_proto import ReducedMinco

orc = Oracle()
for S in (3, 4):
    mb = api.MincoBatch(default_params(S), device=0)
    for N in (1, 2, 5, 8, 9, 16, 17, 32):
        rng = np.random.default_rng(10 * N + S)
        B = 37
        head = rng.normal(size=(B, S, 3)); tail = rng.normal(size=(B, S, 3))
        q = np.cumsum(rng.normal(size=(B, max(N - 1, 1), 3)), axis=1)[:, : max(N - 1, 0)]
        if N == 1:
            q = np.zeros((B, 0, 3))
        q = np.ascontiguousarray(q)
        T = rng.uniform(0.5, 2.5, size=(B, N))
        out = mb.minco_forward(head, tail, q, T)
        gdC = rng.normal(size=(B, 2 * S * N, 3)); gdT = rng.normal(size=(B, N))
        gq, gT = mb.minco_propagate(head, tail, q, T, gdC, gdT)
        e = dict(c=0.0, crow=0.0, E=0.0, gdC=0.0, gdT=0.0, gq=0.0, gT=0.0, c_np=0.0, gq_np=0.0, gT_np=0.0)
        for b in range(B):
            ref = orc.minco_forward(S, head[b], tail[b], q[b], T[b])
            sc = np.abs(ref["coeffs"]).max()
            e["c"] = max(e["c"], np.abs(out["coeffs"][b] - ref["coeffs"]).max() / sc)
            e["E"] = max(e["E"], abs(out["energy"][b] - ref["energy"]) / abs(ref["energy"]))
            e["gdC"] = max(e["gdC"], np.abs(out["gdC"][b] - ref["gdC"]).max() / np.abs(ref["gdC"]).max())
            e["gdT"] = max(e["gdT"], np.abs(out["gdT"][b] - ref["gdT"]).max() / np.abs(ref["gdT"]).max())
            gq_ref, gT_ref = orc.minco_propagate(S, head[b], tail[b], q[b], T[b], gdC[b], gdT[b])
            if N > 1:
                e["gq"] = max(e["gq"], np.abs(gq[b] - gq_ref).max() / np.abs(gq_ref).max())
            e["gT"] = max(e["gT"], np.abs(gT[b] - gT_ref).max() / np.abs(gT_ref).max())
            rm = ReducedMinco(S); rm.set_conditions(head[b], tail[b], N)
            c = rm.set_parameters(q[b], T[b])
            e["c_np"] = max(e["c_np"], np.abs(c - ref["coeffs"]).max() / sc)
            g2, t2 = rm.propagate_grad(gdC[b], gdT[b])
            if N > 1:
                e["gq_np"] = max(e["gq_np"], np.abs(g2 - gq_ref).max() / np.abs(gq_ref).max())
            e["gT_np"] = max(e["gT_np"], np.abs(t2 - gT_ref).max() / np.abs(gT_ref).max())
        print(f"S={S} N={N:2d} " + " ".join(f"{k}={v:.1e}" for k, v in e.items()), flush=True)
    mb.close()
