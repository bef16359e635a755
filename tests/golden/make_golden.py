#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs + CPU-oracle outputs for the MINCO hot path.

The reference ships no golden vectors for this path, so these are produced by the
pinned oracle (oracle/, see its headers for how it is pinned) and, for the L-BFGS traces, by the
REFERENCE's own gcopter/lbfgs.hpp compiled verbatim (oracle/_ref) -- that part is real reference output.
Run from the repo root in the build container:   python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from allocnet_b200 import synth                      # noqa: E402
from allocnet_b200.params import default_params, energy_only   # noqa: E402
from oracle.pyoracle import Oracle                   # noqa: E402

CASES = [  # name, S, N, K, B, ragged, energy-only
    ("s3_n8_k16", 3, 8, 16, 48, False, False),
    ("s3_n5_k16", 3, 5, 16, 32, False, False),
    ("s3_n16_k16", 3, 16, 16, 16, False, False),
    ("s3_n8_k0", 3, 8, 0, 32, False, True),
    ("s3_n12_k9_ragged", 3, 12, 9, 24, True, False),
    ("s4_n8_k16", 4, 8, 16, 24, False, False),
    ("s3_n1_k4", 3, 1, 4, 8, False, False),
]


def main():
    orc = Oracle()
    assert orc.ref is not None, "oracle/_ref missing: build it first (make -C oracle)"
    for name, S, N, K, B, ragged, eonly in CASES:
        prm = default_params(S)
        if eonly:
            prm = energy_only(prm)
        pb = synth.make_problems(B, N=N, K=K, S=S, ragged_rows=ragged)
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        x0 = pb.x0()
        x1 = x0 + 0.05 * rng.standard_normal(x0.shape)
        f0, g0 = orc.cost_batch(prm, pb, x0)
        f1, g1 = orc.cost_batch(prm, pb, x1)
        T = synth.forward_t(x1[:, :N])
        q = x1[:, N:].reshape(B, max(N - 1, 0), 3)
        fw = [orc.minco_forward(S, pb.head[b], pb.tail[b], q[b], T[b]) for b in range(B)]
        gdC = rng.standard_normal((B, 2 * S * N, 3)); gdT = rng.standard_normal((B, N))
        pr = [orc.minco_propagate(S, pb.head[b], pb.tail[b], q[b], T[b], gdC[b], gdT[b]) for b in range(B)]
        out = dict(S=S, N=N, K=K, B=B, ragged=ragged, energy_only=eonly, first=0,
                   x0=x0, x1=x1, f0=f0, g0=g0, f1=f1, g1=g1,
                   coeffs=np.stack([r["coeffs"] for r in fw]), energy=np.array([r["energy"] for r in fw]),
                   gdC_E=np.stack([r["gdC"] for r in fw]), gdT_E=np.stack([r["gdT"] for r in fw]),
                   flat=np.stack([r["flat"] for r in fw]), gdC_in=gdC, gdT_in=gdT,
                   gradByPoints=np.stack([r[0] for r in pr]) if N > 1 else np.zeros((B, 0, 3)),
                   gradByTimes=np.stack([r[1] for r in pr]))
        # optimisation: the reference's lbfgs.hpp (oracle/_ref) around the oracle cost; capped and full
        for cap in (3, 5000):
            p2 = default_params(S, max_iterations=cap)
            if eonly:
                p2 = energy_only(p2)
            r = orc.optimize_batch_ref(p2, pb)
            tag = f"opt{cap}_"
            for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
                out[tag + k] = r[k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "f0[0]=%.6e" % f0[0], "evals(full) mean=%.1f" % out["opt5000_evals"].mean())


if __name__ == "__main__":
    main()
