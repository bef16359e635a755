"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU half: the oracle as built HERE reproduces the committed vectors (guards against a silently different
oracle build on the GPU box) and its restated L-BFGS reproduces the traces that the REFERENCE's own
lbfgs.hpp produced when the fixtures were made.  GPU half: the CUDA path against the same vectors,
1e-9 relative on fp64 cost and gradients (BASELINE.json north_star), integers exact."""
import glob
import os

import numpy as np
import pytest

from allocnet_b200 import synth
from allocnet_b200 import params as P
from allocnet_b200.params import default_params, energy_only

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "s[34]_*.npz")))   # ref_qp_*.npz: tests/ref_qp_util.py
IDS = [os.path.basename(f)[:-4] for f in FILES]


def _load(path):
    z = np.load(path)
    S, N, K, B = int(z["S"]), int(z["N"]), int(z["K"]), int(z["B"])
    pb = synth.make_problems(B, N=N, K=K, S=S, ragged_rows=bool(z["ragged"]))
    prm = default_params(S)
    if bool(z["energy_only"]):
        prm = energy_only(prm)
    return z, pb, prm, S, N, K, B


def _rel(a, b):
    a = np.asarray(a, float).reshape(len(a), -1); b = np.asarray(b, float).reshape(len(b), -1)
    return float(np.max(np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-300)))


def test_fixture_set_is_present():
    assert len(FILES) >= 7


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_reproduces_golden(oracle, path):
    z, pb, prm, S, N, K, B = _load(path)
    np.testing.assert_array_equal(pb.x0(), z["x0"])          # the seeded generator is part of the fixture
    for x, f, g in ((z["x0"], z["f0"], z["g0"]), (z["x1"], z["f1"], z["g1"])):
        fo, go = oracle.cost_batch(prm, pb, x)
        assert _rel(fo[:, None], f[:, None]) <= 1e-12 and _rel(go, g) <= 1e-11
    T = synth.forward_t(z["x1"][:, :N]); q = z["x1"][:, N:].reshape(B, max(N - 1, 0), 3)
    for b in range(0, B, 5):
        r = oracle.minco_forward(S, pb.head[b], pb.tail[b], q[b], T[b])
        np.testing.assert_allclose(r["coeffs"], z["coeffs"][b], rtol=0, atol=1e-11 * np.abs(z["coeffs"][b]).max())
        np.testing.assert_allclose(r["energy"], z["energy"][b], rtol=1e-12)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_restated_lbfgs_reproduces_reference_traces(oracle_strict, path):
    """opt3_*: three iterations of the reference's lbfgs.hpp (recorded).  The restated driver, same cost,
    must give the same counts and codes; iterates to 1e-9 (the recording used the FMA-contracted cost build)."""
    z, pb, prm, S, N, K, B = _load(path)
    p3 = default_params(S, max_iterations=3)
    if bool(z["energy_only"]):
        p3 = energy_only(p3)
    r = oracle_strict.optimize_batch(p3, pb)
    same = (r["evals"] == z["opt3_evals"]) & (r["iters"] == z["opt3_iters"]) & (r["status"] == z["opt3_status"])
    assert same.mean() >= 0.95
    assert _rel(r["x"][same], z["opt3_x"][same]) <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_gpu_matches_golden(path):
    from allocnet_b200 import api
    z, pb, prm, S, N, K, B = _load(path)
    tol = 1e-9
    mb = api.MincoBatch(prm, device=0)
    try:
        mb.set_problems(pb)
        for x, f, g in ((z["x0"], z["f0"], z["g0"]), (z["x1"], z["f1"], z["g1"])):
            fd, gd = mb.evaluate(x)
            assert _rel(fd[:, None], f[:, None]) <= tol and _rel(gd, g) <= tol
        # MINCO building blocks at x1
        T = synth.forward_t(z["x1"][:, :N]); q = z["x1"][:, N:].reshape(B, max(N - 1, 0), 3)
        out = mb.minco_forward(pb.head, pb.tail, q, T)
        ctol = 1e-9
        assert _rel(out["coeffs"], z["coeffs"]) <= ctol and _rel(out["flat"], z["flat"]) <= ctol
        assert _rel(out["energy"][:, None], z["energy"][:, None]) <= tol
        assert _rel(out["gdC"], z["gdC_E"]) <= tol and _rel(out["gdT"], z["gdT_E"]) <= tol
        gq, gT = mb.minco_propagate(pb.head, pb.tail, q, T, z["gdC_in"], z["gdT_in"])
        ptol = 1e-9
        if N > 1:
            assert _rel(gq, z["gradByPoints"]) <= ptol
        assert _rel(gT, z["gradByTimes"]) <= ptol
        # three L-BFGS iterations: same control flow as the reference's lbfgs.hpp on (almost) the same numbers
        mb.set_params(default_params(S, max_iterations=3) if not bool(z["energy_only"]) else energy_only(default_params(S, max_iterations=3)))
        r = mb.optimize(z["x0"])
        same = (r["evals"] == z["opt3_evals"]) & (r["iters"] == z["opt3_iters"]) & (r["status"] == z["opt3_status"])
        assert same.mean() >= 0.9, same.mean()
        assert (r["status"][same] == P.LBFGSERR_MAXIMUMITERATION).all() or N == 1
        assert _rel(r["x"][same], z["opt3_x"][same]) <= 1e-6
        # full runs end where the reference-driven CPU run ends (statistically: iterates fork at near-ties)
        mb.set_params(default_params(S, max_iterations=5000) if not bool(z["energy_only"]) else energy_only(default_params(S, max_iterations=5000)))
        r = mb.optimize(z["x0"])
        ok = (r["status"] >= 0) & (z["opt5000_status"] >= 0)
        assert ((r["status"] >= 0) == (z["opt5000_status"] >= 0)).mean() >= 0.85 and ok.mean() >= 0.8
        rel = np.abs(r["f"][ok] - z["opt5000_f"][ok]) / np.abs(z["opt5000_f"][ok])
        assert np.median(rel) <= 1e-2, np.median(rel)      # 16-48 problems; CPU-vs-CPU forks are of this size
    finally:
        mb.close()
