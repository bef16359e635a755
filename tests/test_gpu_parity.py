"""GPU parity: the sm_100a path (through the C-ABI, allocnet_b200/libmincob.so) against the CPU
oracle on the same seeded inputs.  Tolerance: 1e-9 relative on fp64 cost and gradients
(BASELINE.json north_star), stated per test.  Integer outputs (status, counts, layout) exact."""
import numpy as np
import pytest

from allocnet_b200 import api, synth
from allocnet_b200 import params as P
from allocnet_b200.params import default_params, energy_only

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rel_rows(a, b):
    """max over problems of ||a-b||_inf / ||b||_inf (row-wise)."""
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    return float(np.max(np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-300)))


@pytest.fixture(scope="module")
def handles():
    hs = {S: api.MincoBatch(default_params(S), device=0) for S in (3, 4)}
    yield hs
    for h in hs.values():
        h.close()


def _rand_minco(rng, B, S, N):
    head = rng.normal(size=(B, S, 3)); tail = rng.normal(size=(B, S, 3))
    q = np.cumsum(rng.normal(size=(B, max(N - 1, 1), 3)), axis=1)[:, : max(N - 1, 0)]
    if N == 1:
        q = np.zeros((B, 0, 3))
    T = rng.uniform(0.5, 2.5, size=(B, N))
    return head, tail, np.ascontiguousarray(q), T


@pytest.mark.parametrize("S", [3, 4])
@pytest.mark.parametrize("N", [1, 2, 5, 8, 9, 16, 17, 32])
def test_minco_forward_and_propagate(handles, oracle, S, N):
    """setParameters/getCoeffs/getEnergy/getEnergyPartialGradBy{Coeffs,Times}/getTrajectory and
    propogateGrad vs the banded oracle: 1e-9 everywhere, MINCO_S3NU and MINCO_S4NU alike (measured on the
    B200, tools/s4_precision.py / profiles/r02_s4_precision.txt: <= 7e-14 for S=3, <= 1.6e-11 for S=4 up to 32 pieces)."""
    rng = np.random.default_rng(10 * N + S)
    B = 37
    head, tail, q, T = _rand_minco(rng, B, S, N)
    out = handles[S].minco_forward(head, tail, q, T)
    ctol = TOL
    gdC = rng.normal(size=(B, 2 * S * N, 3)); gdT = rng.normal(size=(B, N))
    gq, gT = handles[S].minco_propagate(head, tail, q, T, gdC, gdT)
    for b in range(B):
        ref = oracle.minco_forward(S, head[b], tail[b], q[b], T[b])
        sc = np.abs(ref["coeffs"]).max()
        assert np.abs(out["coeffs"][b] - ref["coeffs"]).max() <= ctol * sc
        assert np.abs(out["flat"][b] - ref["flat"]).max() <= ctol * sc
        assert abs(out["energy"][b] - ref["energy"]) <= TOL * abs(ref["energy"])
        assert np.abs(out["gdC"][b] - ref["gdC"]).max() <= TOL * np.abs(ref["gdC"]).max()
        assert np.abs(out["gdT"][b] - ref["gdT"]).max() <= TOL * np.abs(ref["gdT"]).max()
        gq_ref, gT_ref = oracle.minco_propagate(S, head[b], tail[b], q[b], T[b], gdC[b], gdT[b])
        ptol = TOL
        if N > 1:
            assert np.abs(gq[b] - gq_ref).max() <= ptol * np.abs(gq_ref).max()
        assert np.abs(gT[b] - gT_ref).max() <= ptol * np.abs(gT_ref).max()


@pytest.mark.parametrize("name", ["n2", "n3", "n5", "n8"])
def test_device_minco_against_reference_qp_matrices(handles, name):
    """The DEVICE's MINCO_S3NU (setParameters, getEnergy, partial gradients, propogateGrad) against matrices built
    by the reference's own code (tests/golden/ref_qp_*.npz: Q, A, b and autograd Jacobians dQ/dT, dA/dT of
    network/utils/min_traj_opt.py::fill_eq_obj, the twin of planner/qp_solver.hpp:148-242) -- no oracle involved:
    coefficients == KKT minimiser (1e-9), A z == b, E == z^T Q z, dE/dc == 2 Q z, dE/dT == z^T dQ z, and the total
    gradients == KKT sensitivities (2e-9)."""
    from ref_qp_util import load_ref_qp, ref_qp_problem, ref_qp_kkt
    c = load_ref_qp(name)
    head, tail, q, T = ref_qp_problem(c, 8)
    N = len(T)
    z, lam_a, lam_w = ref_qp_kkt(c, q)
    out = handles[3].minco_forward(head[None], tail[None], q[None], T[None])
    zd = out["flat"][0].reshape(-1)
    Q, A, b, dQ, dA = c["Q"], c["A"], c["b"], c["dQ"], c["dA"]
    np.testing.assert_allclose(zd, z, rtol=0, atol=TOL * np.abs(z).max())
    assert np.max(np.abs(A @ zd - b)) <= TOL * max(1.0, np.abs(b).max(), np.abs(zd).max())
    assert abs(out["energy"][0] - zd @ Q @ zd) <= TOL * abs(out["energy"][0])
    gdC_flat = out["gdC"][0].reshape(N, 6, 3)[:, ::-1, :].transpose(0, 2, 1).reshape(-1)
    np.testing.assert_allclose(gdC_flat, 2.0 * Q @ zd, rtol=0, atol=TOL * np.abs(Q @ zd).max())
    gdT_ref = np.array([zd @ dQ[i] @ zd for i in range(N)])
    np.testing.assert_allclose(out["gdT"][0], gdT_ref, rtol=0, atol=TOL * np.abs(gdT_ref).max())
    gq, gT = handles[3].minco_propagate(head[None], tail[None], q[None], T[None], out["gdC"], out["gdT"])
    gT_ref = np.array([z @ dQ[i] @ z + 2.0 * lam_a @ (dA[i] @ z) for i in range(N)])
    gq_ref = -2.0 * lam_w.reshape(N - 1, 3)
    np.testing.assert_allclose(gT[0], gT_ref, rtol=0, atol=2 * TOL * np.abs(gT_ref).max())
    np.testing.assert_allclose(gq[0], gq_ref, rtol=0, atol=2 * TOL * np.abs(gq_ref).max())


CASES = [  # (S, N, K, B, ragged, energy_only)
    (3, 8, 16, 4096, False, False),   # BASELINE config 3 shape
    (3, 8, 0, 4096, False, True),     # BASELINE config 2: energy only
    (3, 5, 16, 257, False, False),    # BASELINE config 1 shape (reference ModelMaxSeg = 5)
    (3, 16, 16, 300, False, False),   # BASELINE config 4 shape
    (3, 8, 16, 301, True, False),     # ragged polytope row counts (zero-padded rows)
    (3, 1, 4, 19, False, False),      # single piece
    (3, 2, 7, 33, False, False),
    (3, 12, 9, 65, True, False),
    (3, 32, 16, 40, False, False),    # MINCOB_MAX_PIECES
    (3, 8, 40, 96, True, False),      # > 32 rows per polytope: no row bitmask, half-planes read from global memory
    (3, 5, 50, 64, True, False),      # the reference pads polytopes to 50 rows, ModelMaxSeg = 5 (learning_planner.hpp:157-168)
    (4, 8, 16, 512, False, False),    # MINCO_S4NU
    (4, 5, 16, 64, True, False),
]


@pytest.mark.parametrize("S,N,K,B,ragged,eonly", CASES)
def test_cost_functional_parity(handles, oracle, S, N, K, B, ragged, eonly):
    """f and g of costFunctional at the generator's x0 (many active penalty terms) and at a
    perturbed point: <= 1e-9 relative (S = 3 and 4)."""
    prm = default_params(S)
    if eonly:
        prm = energy_only(prm)
    mb = handles[S]
    mb.set_params(prm)
    pb = synth.make_problems(B, N=N, K=K, S=S, ragged_rows=ragged)
    mb.set_problems(pb)
    rng = np.random.default_rng(S * 1000 + N)
    tol = TOL
    for x in (pb.x0(), pb.x0() + 0.05 * rng.normal(size=(B, pb.nvars))):
        f, g = mb.evaluate(x)
        fo, go = oracle.cost_batch(prm, pb, x, nthreads=8)
        assert np.all(np.isfinite(f)) and np.all(np.isfinite(g))
        assert float(np.max(np.abs(f - fo) / np.abs(fo))) <= tol
        assert rel_rows(g, go) <= tol
    mb.set_params(default_params(S))


def test_cost_functional_near_optimum(handles, oracle):
    """At converged points the gradient is a small difference of large hinge terms (weights 1e4,
    mu 1e-2): the ORACLE's own g moves by up to ~1e-5 absolute when x is perturbed by 1e-15
    relative (a few ulp).  So parity there is stated backward-stably: cost to 1e-9 relative, and per
    problem |g_gpu - g_oracle| <= max(1e-9 * ||g||_inf, 4 x the oracle's own sensitivity to that
    few-ulp perturbation of x)."""
    prm = default_params(3)
    pb = synth.make_problems(512, N=8, K=16, S=3)
    mb = handles[3]
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    x = res["x"]
    f, g = mb.evaluate(x)
    fo, go = oracle.cost_batch(prm, pb, x, nthreads=8)
    assert float(np.max(np.abs(f - fo) / np.abs(fo))) <= TOL
    rng = np.random.default_rng(7)
    sens = np.zeros(x.shape[0])
    for _ in range(12):   # (the sensitivity is a maximum over perturbations: a handful of draws underestimates it)
        xp = x * (1.0 + 1e-15 * np.sign(rng.normal(size=x.shape)))
        _, g2 = oracle.cost_batch(prm, pb, xp, nthreads=8)
        sens = np.maximum(sens, np.abs(g2 - go).max(axis=1))
    err = np.abs(g - go).max(axis=1)
    bound = np.maximum(TOL * np.abs(go).max(axis=1), 4.0 * sens)
    ratio = err / bound
    print("near-optimum gradient: worst err/bound", float(ratio.max()), "over the bound:", int((ratio > 1).sum()), "of", len(ratio))
    assert (err <= bound).all(), (float(ratio.max()), int((err > bound).sum()))
    assert np.median(err / np.maximum(np.abs(go).max(axis=1), 1e-300)) <= TOL   # typical problem: plain 1e-9


def test_optimize_trace_matches_oracle(handles, oracle):
    """Same L-BFGS control flow as gcopter/lbfgs.hpp: capped at a few iterations, the device driver
    must report the same status / iteration / evaluation counts and the same iterate as the CPU
    restatement (which is pinned bit-exact to the verbatim reference header)."""
    for iters in (1, 2, 4):
        prm = default_params(3, max_iterations=iters)
        pb = synth.make_problems(256, N=8, K=16, S=3)
        mb = handles[3]
        mb.set_params(prm)
        mb.set_problems(pb)
        res = mb.optimize(pb.x0())
        ref = oracle.optimize_batch(prm, pb, nthreads=8)
        same = (res["evals"] == ref["evals"]) & (res["iters"] == ref["iters"]) & (res["status"] == ref["status"])
        assert same.mean() >= 0.99, same.mean()
        assert (res["status"][same] == P.LBFGSERR_MAXIMUMITERATION).all()
        assert rel_rows(res["x"][same], ref["x"][same]) <= 1e-7
        assert float(np.max(np.abs(res["f"][same] - ref["f"][same]) / np.abs(ref["f"][same]))) <= 1e-7
    handles[3].set_params(default_params(3))


@pytest.mark.parametrize("S,N,K,B", [(3, 8, 16, 1024), (3, 5, 16, 200), (3, 16, 16, 128), (4, 8, 16, 128), (3, 8, 0, 256),
                                     (3, 5, 50, 96), (3, 32, 8, 48)])
def test_optimize_converges_like_oracle(handles, oracle, S, N, K, B):
    """Full runs: every problem ends with a success code on both sides; the device's reported cost
    at its final x equals the oracle's cost at that x (1e-9); its coefficients equal the oracle's
    getTrajectory at that x; and converged costs agree statistically with the CPU run (iterates
    fork at Armijo near-ties, and the `past` stop test is 1e-5 relative, so not bitwise)."""
    # max_iterations: the default 1000 is a latency cap that a few 16-piece problems reach; lift it
    # here so that both sides run to their own stopping test
    prm = default_params(S, max_iterations=5000)
    if K == 0:
        prm = energy_only(prm)
    mb = handles[S]
    mb.set_params(prm)
    pb = synth.make_problems(B, N=N, K=K, S=S)
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    ref = oracle.optimize_batch(prm, pb, nthreads=8)
    # success codes, except that a few slow problems may stop at the max_iterations safety cap
    # (LBFGSERR_MAXIMUMITERATION, the code the reference returns) on either side
    for st in (res["status"], ref["status"]):
        assert ((st >= 0) | (st == P.LBFGSERR_MAXIMUMITERATION)).all(), np.unique(st, return_counts=True)
        assert (st >= 0).mean() >= 0.98
    tol = TOL
    fo, _ = oracle.cost_batch(prm, pb, res["x"], nthreads=8)
    assert float(np.max(np.abs(res["f"] - fo) / np.abs(fo))) <= tol
    # coefficients / durations of the final x
    for b in range(0, B, max(1, B // 16)):
        inst = oracle.cost_instance(prm, pb, b)
        flat = np.zeros((N, 3, 2 * S)); T = np.zeros(N)
        oracle.lib.orc_cost_flat(inst.inst, res["x"][b].ctypes.data_as(api._dp), flat.ctypes.data_as(api._dp),
                                 T.ctypes.data_as(api._dp))
        inst.close()
        # coefficient k of a piece scales like T^-(D-1-k): compare what it contributes over the piece,
        # c_k T^k (descending storage: power D-1-k), against the largest such term of the trajectory
        pw = T[:, None, None] ** np.arange(2 * S - 1, -1, -1)[None, None, :]
        assert np.abs((res["coeffs"][b] - flat) * pw).max() <= 1e-9 * np.abs(flat * pw).max()
        assert np.abs(res["coeffs"][b] - flat).max() <= 1e-8 * np.abs(flat).max()
        np.testing.assert_allclose(res["T"][b], T, rtol=1e-14)
    # How close can two runs of the SAME algorithm end?  The CPU oracle built with and without FMA
    # contraction (nothing else differs) ends within 2e-2 of itself on 97 % of the corridor problems and
    # on 88 % of the energy-only ones (median 1.5e-4 / 7e-4; the `past` stop test is loose and iterates
    # fork at Armijo near-ties).  The device run is held to that same band, and to the same typical optimum.
    rel = np.abs(res["f"] - ref["f"]) / np.abs(ref["f"])
    # (32 pieces: 97 unknowns and ~2 000 evaluations per problem, so forks are wider; two CPU builds of the oracle agree
    # within 2e-2 on 75-85 % of such problems, tools/diff_variants.py shows the same between two GPU builds)
    assert np.median(rel) <= (2e-3 if N <= 16 else 1e-2) and np.mean(rel < 2e-2) >= (0.90 if K > 0 and N <= 16 else 0.80 if N <= 16 else 0.65), (np.median(rel), rel.max())
    assert abs(np.median(res["f"]) / np.median(ref["f"]) - 1.0) <= (1e-2 if N <= 16 else 3e-2)   # 48 32-piece problems: medians of forked runs
    # effort is comparable (same algorithm): mean evaluation count within 15 % (35 % for the 24-problem case:
    # single 32-piece problems fork by hundreds of evaluations on a 1-ulp difference, see tools/diff_variants.py)
    assert abs(res["evals"].mean() / ref["evals"].mean() - 1.0) <= (0.15 if B >= 96 else 0.30)
    mb.set_params(default_params(S))


def test_repeated_optimize_is_stateless(handles):
    """The same batch optimized twice through one handle gives identical results: no optimizer state (the
    `past` cost ring, history slabs, shared-memory stages) may leak from one problem or call to the next."""
    pb = synth.make_problems(64, N=16, K=16, S=3)
    mb = handles[3]
    mb.set_params(default_params(3, max_iterations=3))
    mb.set_problems(pb)
    mb.optimize(pb.x0())
    mb.set_params(default_params(3))
    a = mb.optimize(pb.x0())
    b = mb.optimize(pb.x0())
    for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
        np.testing.assert_array_equal(a[k], b[k])
    assert np.median(a["iters"]) > 50


def test_lbfgs_parameter_validation_codes(handles):
    """lbfgs.hpp:450-495: invalid parameters return the reference's code before x is touched."""
    pb = synth.make_problems(8, N=8, K=16, S=3)
    mb = handles[3]
    mb.set_problems(pb)
    for kw, code in ((dict(mem_size=0), P.LBFGSERR_INVALID_MEMSIZE), (dict(delta=-1.0), P.LBFGSERR_INVALID_DELTA),
                     (dict(f_dec_coeff=1.5), P.LBFGSERR_INVALID_FDECCOEFF), (dict(max_linesearch=0), P.LBFGSERR_INVALID_MAXLINESEARCH),
                     (dict(s_curv_coeff=1e-5), P.LBFGSERR_INVALID_SCURVCOEFF)):
        mb.set_params(default_params(3, **kw))
        x0 = pb.x0()
        res = mb.optimize(x0)
        assert (res["status"] == code).all()
        np.testing.assert_array_equal(res["x"], x0)
    mb.set_params(default_params(3))


def test_invalid_function_value_status(handles):
    """NaN input -> LBFGSERR_INVALID_FUNCVAL on that problem only (lbfgs.hpp:322), others unaffected."""
    pb = synth.make_problems(16, N=8, K=16, S=3)
    mb = handles[3]
    mb.set_problems(pb)
    x0 = pb.x0(); x0[3, 2] = np.nan
    res = mb.optimize(x0)
    assert res["status"][3] < 0
    ok = np.arange(16) != 3
    assert (res["status"][ok] >= 0).all()


def test_round_trip_properties_full_size(handles):
    """BASELINE size (65 536 x 8 pieces): size-independent properties instead of an oracle run.
    (1) coefficient layout: Trajectory-order coefficients evaluate to the waypoints/boundary states
    of the final x (C^0..C^2 continuity, head/tail PVA);  (2) idempotence: optimizing the optimum
    again stops within a few iterations without increasing the cost."""
    S, N, K, B = 3, 8, 16, 65536
    pb = synth.make_problems(B, N=N, K=K, S=S)
    mb = handles[S]
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    # every problem ends with a reference code; all but a handful with a success code (the rest: the
    # max_iterations safety cap or a line-search give-up at the optimum, as lbfgs.hpp reports them)
    ok = res["status"] >= 0
    assert ok.mean() >= 0.999, np.unique(res["status"], return_counts=True)
    assert np.isin(res["status"][~ok], [P.LBFGSERR_MAXIMUMITERATION, P.LBFGSERR_MAXIMUMLINESEARCH,
                                         P.LBFGSERR_WIDTHTOOSMALL, P.LBFGSERR_MINIMUMSTEP]).all()
    c = res["coeffs"]; T = res["T"]            # [B][N][3][6] descending powers
    D = 2 * S
    pw = np.arange(D - 1, -1, -1)
    def ev(ci, t, d):
        k = pw
        fac = np.array([np.prod([kk - u for u in range(d)]) if kk >= d else 0.0 for kk in k])
        return np.einsum("bak,bk->ba", ci, fac[None, :] * np.where(k[None, :] >= d, t[:, None] ** np.maximum(k[None, :] - d, 0), 0.0))
    q = res["x"][:, N:].reshape(B, N - 1, 3)
    zero = np.zeros(B)
    np.testing.assert_allclose(ev(c[:, 0], zero, 0), pb.head[:, 0], atol=1e-9)
    np.testing.assert_allclose(ev(c[:, 0], zero, 1), pb.head[:, 1], atol=1e-9)
    np.testing.assert_allclose(ev(c[:, N - 1], T[:, N - 1], 0), pb.tail[:, 0], atol=1e-7)
    for i in range(N - 1):
        np.testing.assert_allclose(ev(c[:, i], T[:, i], 0), q[:, i], atol=1e-7)
        np.testing.assert_allclose(ev(c[:, i + 1], zero, 0), q[:, i], atol=1e-9)
        for d in (1, 2, 3, 4):
            a = ev(c[:, i], T[:, i], d); b2 = ev(c[:, i + 1], zero, d)
            assert np.abs(a - b2).max() <= 1e-6 * max(1.0, np.abs(a).max())
    res2 = mb.optimize(res["x"])
    ok2 = ok & (res2["status"] >= 0)
    assert ok2.mean() >= 0.995
    assert (res2["f"][ok2] <= res["f"][ok2] * (1 + 1e-12)).all()
    assert np.median(res2["iters"]) <= 10


def test_gpu_trajectory_in_reference_class(handles, oracle):
    """What the device writes as `coeffs`/`T` goes into the REFERENCE's Trajectory<5> (gcopter/trajectory.hpp
    compiled verbatim, oracle/_ref) through emplace_back, as learning_planner.hpp:205-216 does: the
    reference's getPositions must return head, the optimized waypoints and tail, and twice its
    getTrajCost(3) must be the jerk energy the device reports for the same (q, T)."""
    import ctypes as C
    ref = oracle.ref
    if ref is None or not hasattr(ref, "ref_traj5_create"):
        pytest.skip("oracle/_ref not present")
    dp = C.POINTER(C.c_double)
    ref.ref_traj5_create.restype = C.c_void_p; ref.ref_traj5_create.argtypes = [C.c_int, dp, dp]
    ref.ref_traj5_destroy.argtypes = [C.c_void_p]
    ref.ref_traj5_cost.restype = C.c_double; ref.ref_traj5_cost.argtypes = [C.c_void_p, C.c_int]
    ref.ref_traj5_positions.argtypes = [C.c_void_p, dp]
    B, N = 64, 8
    pb = synth.make_problems(B, N=N, K=16, S=3)
    mb = handles[3]
    mb.set_params(default_params(3))
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    q = res["x"][:, N:].reshape(B, N - 1, 3)
    fw = mb.minco_forward(pb.head, pb.tail, q, res["T"])
    for b in range(B):
        c = np.ascontiguousarray(res["coeffs"][b]); T = np.ascontiguousarray(res["T"][b])
        h = ref.ref_traj5_create(N, T.ctypes.data_as(dp), c.ctypes.data_as(dp))
        try:
            P = np.zeros((N + 1, 3)); ref.ref_traj5_positions(h, P.ctypes.data_as(dp))
            np.testing.assert_allclose(P[0], pb.head[b, 0], atol=1e-9)
            np.testing.assert_allclose(P[1:N], q[b], atol=1e-8)
            np.testing.assert_allclose(P[N], pb.tail[b, 0], atol=1e-7)
            assert abs(2.0 * ref.ref_traj5_cost(h, 3) - fw["energy"][b]) <= 1e-9 * abs(fw["energy"][b])
        finally:
            ref.ref_traj5_destroy(h)


def test_feasibility_report(handles, oracle):
    """mincob_check_feasibility (sampled max |v|, |a|, |j| and corridor residual) against numpy on the same
    grid, and -- for the derivative values themselves -- against the reference's Trajectory<5>::getVel/getAcc/
    getJer (gcopter/trajectory.hpp compiled verbatim, oracle/_ref) at the same times."""
    import ctypes as C
    B, N, K, R = 48, 8, 16, 24
    pb = synth.make_problems(B, N=N, K=K, S=3)
    mb = handles[3]
    mb.set_params(default_params(3))
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    rep = mb.check_feasibility(res["coeffs"], res["T"], samples=R)
    c = res["coeffs"]; T = res["T"]                       # [B][N][3][6] descending
    pw = np.arange(5, -1, -1)
    u = np.linspace(0.0, 1.0, R + 1)
    want = np.zeros((B, 4)); want[:, 3] = -np.inf
    for i in range(N):
        t = T[:, i, None] * np.arange(R + 1)[None, :] / R  # same expression as the kernel: T*j/R
        def deriv(d):
            fac = np.array([np.prod([k - q for q in range(d)]) if k >= d else 0.0 for k in pw])
            tp = np.where(pw[None, None, :] >= d, t[:, :, None] ** np.maximum(pw - d, 0)[None, None, :], 0.0)
            return np.einsum("bak,btk->bta", c[:, i] * fac[None, None, :], tp)
        for col, d in ((0, 1), (1, 2), (2, 3)):
            want[:, col] = np.maximum(want[:, col], np.linalg.norm(deriv(d), axis=2).max(axis=1))
        pos = deriv(0)
        hp = pb.hpolys[:, i]                              # [B][K][4]
        viol = np.einsum("bkx,btx->btk", hp[:, :, :3], pos) + hp[:, None, :, 3]
        want[:, 3] = np.maximum(want[:, 3], viol.reshape(B, -1).max(axis=1))
    np.testing.assert_allclose(rep[:, :3], want[:, :3], rtol=1e-9)
    np.testing.assert_allclose(rep[:, 3], want[:, 3], rtol=1e-9, atol=1e-9)
    # the optimizer did its job: limits (4, 6, 12) and corridor hold up to the softness of the penalty
    assert np.median(rep[:, 0]) <= 4.0 * 1.02 and np.median(rep[:, 1]) <= 6.0 * 1.02 and np.median(rep[:, 3]) <= 0.02
    ref = oracle.ref
    if ref is not None and hasattr(ref, "ref_traj5_create"):
        dp = C.POINTER(C.c_double)
        ref.ref_traj5_create.restype = C.c_void_p; ref.ref_traj5_create.argtypes = [C.c_int, dp, dp]
        ref.ref_traj5_destroy.argtypes = [C.c_void_p]
        ref.ref_traj5_eval.argtypes = [C.c_void_p, C.c_double, dp, dp, dp, dp]
        for b in range(0, B, 7):
            cc = np.ascontiguousarray(c[b]); TT = np.ascontiguousarray(T[b])
            h = ref.ref_traj5_create(N, TT.ctypes.data_as(dp), cc.ctypes.data_as(dp))
            vm = 0.0
            t0 = np.concatenate([[0.0], np.cumsum(TT)])
            for i in range(N):
                for j in range(1, R):                     # interior points: no piece-boundary ambiguity
                    o = [np.zeros(3) for _ in range(4)]
                    ref.ref_traj5_eval(h, t0[i] + TT[i] * j / R, *[q.ctypes.data_as(dp) for q in o])
                    vm = max(vm, np.linalg.norm(o[1]))
            ref.ref_traj5_destroy(h)
            assert vm <= rep[b, 0] * (1 + 1e-9) and vm >= rep[b, 0] * 0.97


def test_fp64_peak_measurement_is_plausible(handles):
    """mincob_measure_fp64_peak: the denominator of bench.py's roofline_fp64.  B200 vector fp64 is 64 DFMA/clk/SM
    (ncu: sm__sass_thread_inst_executed_op_dfma_pred_on peak_sustained), ~37 TFLOP/s at 1.96 GHz."""
    tf = handles[3].measure_fp64_peak()
    assert 20.0 < tf < 45.0, tf


def _dense_rates(coeffs, T, pts=40001):
    want = np.zeros(3)
    for i in range(coeffs.shape[0]):
        t = np.linspace(0.0, T[i], pts)
        for d in (1, 2, 3):
            v = np.stack([np.polyval(np.polyder(coeffs[i, a], d), t) for a in range(3)])
            want[d - 1] = max(want[d - 1], np.sqrt((v * v).sum(axis=0).max()))
    return want


def test_exact_max_rates(handles, oracle):
    """mincob_max_rates against the REFERENCE's Trajectory<5>::getMaxVelRate / getMaxAccRate (trajectory.hpp and
    root_finder.hpp compiled verbatim, oracle/_ref; its root tolerance is FLT_EPSILON / T) and, for all three rates and
    also without oracle/_ref, against a dense sampling; checkMaxVelRate / checkMaxAccRate are `rate < limit`."""
    import ctypes as C
    B, N = 96, 8
    pb = synth.make_problems(B, N=N, K=16, S=3)
    mb = handles[3]
    mb.set_params(default_params(3))
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    rates = mb.max_rates(res["coeffs"], res["T"])
    assert rates.shape == (B, 3) and np.isfinite(rates).all()
    # never below the sampled report of the same trajectories
    rep = mb.check_feasibility(res["coeffs"], res["T"], samples=64)
    assert (rates >= rep[:, :3] * (1.0 - 1e-12)).all()
    for b in range(0, B, 5):
        np.testing.assert_allclose(rates[b], _dense_rates(res["coeffs"][b], res["T"][b]), rtol=1e-7)
    ref = oracle.ref
    if ref is not None and hasattr(ref, "ref_traj5_max_vel_rate"):
        dp = C.POINTER(C.c_double)
        ref.ref_traj5_create.restype = C.c_void_p; ref.ref_traj5_create.argtypes = [C.c_int, dp, dp]
        ref.ref_traj5_destroy.argtypes = [C.c_void_p]
        for fn in (ref.ref_traj5_max_vel_rate, ref.ref_traj5_max_acc_rate):
            fn.restype = C.c_double; fn.argtypes = [C.c_void_p]
        for fn in (ref.ref_traj5_check_max_vel_rate, ref.ref_traj5_check_max_acc_rate):
            fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_double]
        for b in range(B):
            c = np.ascontiguousarray(res["coeffs"][b]); T = np.ascontiguousarray(res["T"][b])
            h = ref.ref_traj5_create(N, T.ctypes.data_as(dp), c.ctypes.data_as(dp))
            try:
                assert abs(ref.ref_traj5_max_vel_rate(h) - rates[b, 0]) <= 1e-8 * rates[b, 0]
                assert abs(ref.ref_traj5_max_acc_rate(h) - rates[b, 1]) <= 1e-8 * rates[b, 1]
                for lim in (4.0, 4.1, 3.9):
                    if abs(rates[b, 0] - lim) > 1e-6:
                        assert bool(ref.ref_traj5_check_max_vel_rate(h, lim)) == bool(rates[b, 0] < lim)
                for lim in (6.0, 6.2, 5.5):
                    if abs(rates[b, 1] - lim) > 1e-6:
                        assert bool(ref.ref_traj5_check_max_acc_rate(h, lim)) == bool(rates[b, 1] < lim)
            finally:
                ref.ref_traj5_destroy(h)
    # S = 4 (septic pieces): dense sampling, and the reference's Trajectory<7> -- which also takes the device's
    # coefficients through emplace_back and returns the waypoints and the snap energy (up to the reference's own
    # m_34 typo, trajectory.hpp:385: 2 getTrajCost(4) = E - sum 80 T^2 c5.c4, tests/test_oracle_minco.py)
    B4, N4 = 32, 6
    pb4 = synth.make_problems(B4, N=N4, K=16, S=4)
    mb4 = handles[4]
    mb4.set_params(default_params(4))
    mb4.set_problems(pb4)
    r4 = mb4.optimize(pb4.x0())
    rt4 = mb4.max_rates(r4["coeffs"], r4["T"])
    for b in range(0, B4, 7):
        np.testing.assert_allclose(rt4[b], _dense_rates(r4["coeffs"][b], r4["T"][b]), rtol=1e-7)
    if ref is not None and hasattr(ref, "ref_traj7_create"):
        dp = C.POINTER(C.c_double)
        ref.ref_traj7_create.restype = C.c_void_p; ref.ref_traj7_create.argtypes = [C.c_int, dp, dp]
        ref.ref_traj7_destroy.argtypes = [C.c_void_p]
        ref.ref_traj7_cost.restype = C.c_double; ref.ref_traj7_cost.argtypes = [C.c_void_p, C.c_int]
        ref.ref_traj7_positions.argtypes = [C.c_void_p, dp]
        for fn in (ref.ref_traj7_max_vel_rate, ref.ref_traj7_max_acc_rate):
            fn.restype = C.c_double; fn.argtypes = [C.c_void_p]
        q4 = r4["x"][:, N4:].reshape(B4, N4 - 1, 3)
        fw4 = mb4.minco_forward(pb4.head, pb4.tail, q4, r4["T"])
        for b in range(B4):
            c = np.ascontiguousarray(r4["coeffs"][b]); T = np.ascontiguousarray(r4["T"][b])   # [N][3][8] descending
            h = ref.ref_traj7_create(N4, T.ctypes.data_as(dp), c.ctypes.data_as(dp))
            try:
                assert abs(ref.ref_traj7_max_vel_rate(h) - rt4[b, 0]) <= 1e-7 * rt4[b, 0]
                assert abs(ref.ref_traj7_max_acc_rate(h) - rt4[b, 1]) <= 1e-7 * rt4[b, 1]
                P = np.zeros((N4 + 1, 3)); ref.ref_traj7_positions(h, P.ctypes.data_as(dp))
                np.testing.assert_allclose(P[0], pb4.head[b, 0], atol=1e-9)
                np.testing.assert_allclose(P[1:N4], q4[b], atol=1e-6)
                np.testing.assert_allclose(P[N4], pb4.tail[b, 0], atol=1e-5)
                quirk = sum(80.0 * T[i] ** 2 * float((c[i, :, 2] * c[i, :, 3]).sum()) for i in range(N4))   # c5 = k 2, c4 = k 3
                E = fw4["energy"][b]
                assert abs(2.0 * ref.ref_traj7_cost(h, 4) + quirk - E) <= 1e-7 * abs(E)
            finally:
                ref.ref_traj7_destroy(h)


def test_shape_and_state_validation():
    """Error behaviour of the boundary (include/mincob.h): bad shapes are MINCOB_E_INVALID before anything is
    launched, compute calls before set_problems are MINCOB_E_STATE, and the handle stays usable afterwards."""
    import ctypes as C
    mb = api.MincoBatch(default_params(3), device=0)
    L = mb.L
    pb = synth.make_problems(4, N=8, K=16, S=3)
    x = pb.x0(); f = np.zeros(4); g = np.zeros_like(x)
    E_INVALID = L.mincob_set_problems(mb.h, 0, 8, 16, api._np_ptr(pb.head), api._np_ptr(pb.tail), api._np_ptr(pb.hpolys), api._np_ptr(pb.hrows))
    assert E_INVALID != 0 and b"B must be" in L.mincob_last_error(mb.h)            # empty batch
    E_STATE = L.mincob_evaluate(mb.h, api._np_ptr(x), api._np_ptr(f), api._np_ptr(g))
    assert E_STATE != 0 and E_STATE != E_INVALID                                     # nothing set yet
    for N in (0, 33):                                                                # MINCOB_MAX_PIECES = 32
        assert L.mincob_set_problems(mb.h, 4, N, 16, api._np_ptr(pb.head), api._np_ptr(pb.tail), api._np_ptr(pb.hpolys), api._np_ptr(pb.hrows)) == E_INVALID
    assert L.mincob_set_problems(mb.h, 4, 8, 16, api._np_ptr(pb.head), api._np_ptr(pb.tail), None, None) == E_INVALID   # K > 0 without rows
    assert L.mincob_set_problems(mb.h, 4, 8, 16, None, api._np_ptr(pb.tail), api._np_ptr(pb.hpolys), api._np_ptr(pb.hrows)) == E_INVALID
    mb.set_problems(pb)                                                              # still works
    assert L.mincob_evaluate(mb.h, None, api._np_ptr(f), api._np_ptr(g)) == E_INVALID
    f1, g1 = mb.evaluate(x)
    assert np.isfinite(f1).all() and np.isfinite(g1).all()
    mb.close()


@pytest.mark.parametrize("S,N", [(3, 5), (3, 8), (4, 4), (3, 1)])
def test_autograd_layer_against_torch_reference(handles, S, N):
    """allocnet_b200.autograd.minco_layer (forward = setParameters/getEnergy kernel, backward = propogateGrad kernel)
    against torch.autograd through a plain PyTorch fp64 dense MINCO: values and the gradients of a loss that uses both
    outputs (energy + time + a functional of the coefficients), with respect to waypoints and durations."""
    import torch
    from allocnet_b200.autograd import minco_layer
    from torch_minco_ref import torch_dense_minco
    dev = torch.device("cuda:0")
    B = 6
    pb = synth.make_problems(B, N=N, K=0, S=S)
    mb = handles[S]
    head = torch.tensor(pb.head, device=dev); tail = torch.tensor(pb.tail, device=dev)
    g = torch.Generator().manual_seed(5)
    q = torch.tensor(pb.q0 if N > 1 else np.zeros((B, 0, 3)), device=dev).requires_grad_(N > 1)
    T = (torch.tensor(pb.T0, device=dev) * (0.7 + 0.6 * torch.rand(B, N, generator=g, dtype=torch.float64).to(dev))).requires_grad_(True)
    wc = torch.randn(B, 2 * S * N, 3, generator=g, dtype=torch.float64).to(dev)

    def loss_of(E, c, T):
        return (E * torch.linspace(0.5, 1.5, B, dtype=torch.float64, device=dev)).sum() + 20.0 * T.sum() + (wc * c).sum() + 0.5 * (c * c).sum()
    E, c = minco_layer(mb, head, tail, q, T)
    L = loss_of(E, c, T)
    grads = torch.autograd.grad(L, [T] + ([q] if N > 1 else []))
    q2 = q.detach().clone().requires_grad_(N > 1); T2 = T.detach().clone().requires_grad_(True)
    Es, cs = [], []
    for b in range(B):
        e, cc = torch_dense_minco(S, head[b], tail[b], q2[b], T2[b])
        Es.append(e); cs.append(cc)
    E2 = torch.stack(Es); c2 = torch.stack(cs)
    L2 = loss_of(E2, c2, T2)
    grads2 = torch.autograd.grad(L2, [T2] + ([q2] if N > 1 else []))
    tol = 1e-9 if S == 3 else 1e-6                      # septic monomial systems are worse conditioned
    assert float(((E - E2).abs() / E2.abs()).max().detach()) <= tol
    assert float(((c - c2).abs().max() / c2.abs().max()).detach()) <= tol
    for ga, gb in zip(grads, grads2):
        assert float((ga - gb).abs().max() / gb.abs().max()) <= 1e-8, (S, N)


def test_config4_pipeline_net_warm_start_on_device(handles, oracle):
    """BASELINE.json configs[3] end to end on the device: 16-piece problems, initial durations from the batched
    time-allocation forward (allocnet_b200/timealloc.py, run on cuda:0 over sliding 5-piece windows; weights here
    are synthetic because the reference's model files do not travel -- tests/test_timealloc.py checks the same code
    against the reference TorchScript model where it exists), then the optimizer from that start: cost/gradient parity
    at the warm start (1e-9) and the same converged optimum statistics as the CPU oracle from the same start."""
    import torch
    from allocnet_b200 import timealloc
    B, N, K = 192, 16, 16
    pb = synth.make_problems(B, N=N, K=K, S=3)
    w = timealloc.random_weights(seed=9, device="cuda:0")
    w["tfs_output_layer.bias"] = torch.tensor([1.2], device="cuda:0")          # durations around 1.2 s, all positive
    w["stop_token_output_layer.0.bias"] = torch.tensor([-30.0], device="cuda:0")  # never stop early
    T0 = timealloc.warm_start_durations(w, pb, device="cuda:0")
    assert T0.shape == (B, N) and (T0 > 0).all() and np.abs(T0 - pb.T0).max() > 1e-3   # the net's answer, not the fallback
    import dataclasses
    pbw = dataclasses.replace(pb, T0=T0)
    x0 = pbw.x0()
    prm = default_params(3, max_iterations=5000)
    mb = handles[3]
    mb.set_params(prm)
    mb.set_problems(pbw)
    f, g = mb.evaluate(x0)
    fo, go = oracle.cost_batch(prm, pbw, x0, nthreads=8)
    assert float(np.max(np.abs(f - fo) / np.abs(fo))) <= TOL
    assert float(np.max(np.abs(g - go).max(axis=1) / np.abs(go).max(axis=1))) <= TOL
    res = mb.optimize(x0)
    ref = oracle.optimize_batch(prm, pbw, x0=x0, nthreads=8)
    assert (res["status"] >= 0).mean() >= 0.98 and (ref["status"] >= 0).mean() >= 0.98
    fo2, _ = oracle.cost_batch(prm, pbw, res["x"], nthreads=8)
    assert float(np.max(np.abs(res["f"] - fo2) / np.abs(fo2))) <= TOL
    rel = np.abs(res["f"] - ref["f"]) / np.abs(ref["f"])
    assert np.median(rel) <= 2e-3 and np.mean(rel < 2e-2) >= 0.85, (np.median(rel), rel.max())
    assert abs(res["evals"].mean() / ref["evals"].mean() - 1.0) <= 0.15
    mb.set_params(default_params(3))


def test_gradient_norm_stop(handles, oracle):
    """g_epsilon > 0 with the `past` test switched off: every problem must end with LBFGS_CONVERGENCE (0), the code of
    lbfgs.hpp:531 / :600, and at the returned x the ORACLE's gradient satisfies |g|_inf / max(1, |x|_inf) < g_epsilon
    -- the device evaluates that test as a product and skips its two reductions altogether when g_epsilon = 0."""
    B, N = 64, 5
    prm = energy_only(default_params(3, g_epsilon=1e-2, past=0, max_iterations=2000))
    pb = synth.make_problems(B, N=N, K=0, S=3)
    mb = handles[3]
    mb.set_params(prm)
    mb.set_problems(pb)
    res = mb.optimize(pb.x0())
    ref = oracle.optimize_batch(prm, pb, nthreads=8)
    assert (res["status"] == P.LBFGS_CONVERGENCE).all(), np.unique(res["status"], return_counts=True)
    assert (ref["status"] == P.LBFGS_CONVERGENCE).all()
    _, go = oracle.cost_batch(prm, pb, res["x"], nthreads=8)
    crit = np.abs(go).max(axis=1) / np.maximum(1.0, np.abs(res["x"]).max(axis=1))
    assert (crit < 1e-2 * (1.0 + 1e-6)).all(), float(crit.max())
    assert abs(res["iters"].mean() / ref["iters"].mean() - 1.0) <= 0.25
    mb.set_params(default_params(3))


# ---- round 2: latency mapping ("one warp per trajectory"), fixed-time mode, planner-form rows --------------------

@pytest.mark.parametrize("S,N,K,B", [(3, 8, 16, 96), (3, 5, 16, 50), (3, 16, 16, 40), (4, 8, 16, 33), (3, 5, 50, 31),
                                     (3, 3, 7, 20), (3, 8, 0, 40)])
def test_latency_mapping_follows_the_throughput_mapping(handles, oracle, S, N, K, B):
    """MINCOB_MAP_LATENCY (the lane groups of a warp share one trajectory and split the TESTS of the penalty samples; every
    replica then accumulates the active samples in ascending order, like the throughput mapping) gives the bits of
    MINCOB_MAP_THROUGHPUT: every output of full runs identical, and the reported cost is the oracle's cost at the final x
    to 1e-9.  (Until round 2 the replicas added partial sums and the mappings agreed to rounding only.)"""
    pb = synth.make_problems(B, N=N, K=K, S=S, ragged_rows=(K == 50))
    mb = handles[S]
    mb.set_problems(pb)
    base = energy_only(default_params(S)) if K == 0 else default_params(S)
    out = {}
    for it in (3, 0):
        for mp in (P.MAP_THROUGHPUT, P.MAP_LATENCY):
            prm = P.MincobParams.from_buffer_copy(base)
            prm.mapping = mp
            prm.max_iterations = it if it else 1000
            mb.set_params(prm)
            out[it, mp] = mb.optimize(pb.x0())
            assert mb.last_mapping() == (mp if N <= 16 else P.MAP_THROUGHPUT)
    for it in (3, 0):
        a, b = out[it, P.MAP_THROUGHPUT], out[it, P.MAP_LATENCY]
        for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"max_iterations {it}: {k}")
    full = out[0, P.MAP_LATENCY]
    assert ((full["status"] >= 0) | (full["status"] == P.LBFGSERR_MAXIMUMITERATION)).all()
    fo, _ = oracle.cost_batch(base, pb, full["x"], nthreads=8)
    assert float(np.max(np.abs(full["f"] - fo) / np.abs(fo))) <= TOL
    # coefficients written by replica 0 are the trajectory at the final x
    T = synth.forward_t(full["x"][:, :N])
    np.testing.assert_allclose(full["T"], T, rtol=1e-14)
    mb.set_params(default_params(S))


def test_latency_mapping_is_reproducible_and_batch_independent(handles):
    """For a pinned mapping the result of a problem does not depend on what else is in the batch or on timing:
    problem p optimized alone, in a batch of 7 and in a batch of 300 gives the same bits."""
    prm = default_params(3, mapping=P.MAP_LATENCY)
    mb = handles[3]
    mb.set_params(prm)
    pb = synth.make_problems(300, N=8, K=16, S=3)
    mb.set_problems(pb)
    big = mb.optimize(pb.x0())
    for lo, hi in ((0, 1), (17, 24), (299, 300)):
        sub = pb.slice(lo, hi)
        mb.set_problems(sub)
        r = mb.optimize(sub.x0())
        for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
            np.testing.assert_array_equal(r[k], big[k][lo:hi])
    prm.mapping = P.MAP_THROUGHPUT
    mb.set_params(prm)
    mb.set_problems(pb)
    big = mb.optimize(pb.x0())
    sub = pb.slice(40, 45)
    mb.set_problems(sub)
    r = mb.optimize(sub.x0())
    for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
        np.testing.assert_array_equal(r[k], big[k][40:45])
    mb.set_params(default_params(3))


def test_auto_mapping(handles):
    """MINCOB_MAP_AUTO: a single problem (the planner's call, learning_planning.cpp:143-188) and small batches take the
    latency mapping, a batch that fills the device the throughput mapping."""
    mb = handles[3]
    mb.set_params(default_params(3))
    for B, want in ((1, P.MAP_LATENCY), (64, P.MAP_LATENCY), (16384, P.MAP_THROUGHPUT)):
        pb = synth.make_problems(B, N=5, K=16, S=3)
        mb.set_problems(pb)
        r = mb.optimize(pb.x0())
        assert mb.last_mapping() == want
        assert (r["status"] >= 0).mean() >= 0.97


@pytest.mark.parametrize("S,N,K", [(3, 5, 16), (3, 8, 16), (4, 8, 16), (3, 12, 9)])
def test_fixed_time_mode(handles, oracle, S, N, K):
    """MINCOB_FLAG_FREEZE_TIMES = the call the reference makes today (qp_solver.solve with the network's times fixed,
    learning_planner.hpp:196): the tau block of the gradient is zero, the durations come back bit-identical, only the
    waypoints move; cost/gradient parity with the oracle in the same mode 1e-9, converged costs as in the free mode."""
    prm = default_params(S, flags=P.FLAG_FREEZE_TIMES)
    pb = synth.make_problems(128, N=N, K=K, S=S, ragged_rows=(K == 9))
    mb = handles[S]
    mb.set_params(prm)
    mb.set_problems(pb)
    x0 = pb.x0()
    f, g = mb.evaluate(x0)
    fo, go = oracle.cost_batch(prm, pb, x0, nthreads=8)
    assert (g[:, :N] == 0.0).all() and (go[:, :N] == 0.0).all()
    assert float(np.max(np.abs(f - fo) / np.abs(fo))) <= TOL and rel_rows(g, go) <= TOL
    for mp in (P.MAP_THROUGHPUT, P.MAP_LATENCY):
        prm.mapping = mp
        mb.set_params(prm)
        res = mb.optimize(x0)
        np.testing.assert_array_equal(res["x"][:, :N], x0[:, :N])           # durations untouched
        np.testing.assert_allclose(res["T"], pb.T0, rtol=1e-14)
        assert (res["status"] >= 0).mean() >= 0.97
        assert (res["f"] < f).all()
        fo2, _ = oracle.cost_batch(prm, pb, res["x"], nthreads=8)
        assert float(np.max(np.abs(res["f"] - fo2) / np.abs(fo2))) <= TOL
    ref = oracle.optimize_batch(prm, pb, nthreads=8)
    np.testing.assert_array_equal(ref["x"][:, :N], x0[:, :N])
    # with the generator's durations frozen most problems stay penalty-dominated (the limits cannot be met without
    # changing T), a rough landscape on which two runs fork more widely than in the free mode: populations are compared
    rel = np.abs(res["f"] - ref["f"]) / np.abs(ref["f"])
    assert np.median(rel) <= 5e-2, np.median(rel)
    assert abs(np.median(res["f"]) / np.median(ref["f"]) - 1.0) <= 0.1
    mb.set_params(default_params(S))


def test_planner_rows_flag(handles, oracle):
    """MINCOB_FLAG_PLANNER_ROWS: rows [n, b] with n.p <= b (what LearningPlanner::plan produces, learning_planner.hpp:
    293-299) give bit-identical results to the same rows in GCOPTER sign [n, -b]; host and device entry points."""
    import torch
    pb = synth.make_problems(200, N=5, K=50, S=3, ragged_rows=True)
    pl = pb.slice(0, pb.B)
    pl.hpolys = pb.hpolys.copy()      # slice() shares memory with pb when the range is the whole batch
    pl.hpolys[..., 3] *= -1.0
    mb = handles[3]
    mb.set_params(default_params(3))
    mb.set_problems(pb)
    f0, g0 = mb.evaluate(pb.x0())
    prm = default_params(3, flags=P.FLAG_PLANNER_ROWS)
    mb.set_params(prm)
    mb.set_problems(pl)
    f1, g1 = mb.evaluate(pb.x0())
    np.testing.assert_array_equal(f0, f1); np.testing.assert_array_equal(g0, g1)
    fo, go = oracle.cost_batch(prm, pl, pb.x0(), nthreads=8)
    assert float(np.max(np.abs(f1 - fo) / np.abs(fo))) <= TOL and rel_rows(g1, go) <= TOL
    dev = torch.device("cuda:0")
    th = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    head, tail, hp, hr, x = th(pl.head), th(pl.tail), th(pl.hpolys), th(pl.hrows), th(pb.x0())
    hp_before = hp.clone()
    mb.set_problems_device(pl.B, pl.N, pl.K, head, tail, hp, hr)
    f2 = torch.empty(pl.B, dtype=torch.float64, device=dev); g2 = torch.empty_like(x)
    mb.evaluate_device(x, f2, g2)
    mb.synchronize()
    np.testing.assert_array_equal(f2.cpu().numpy(), f0); np.testing.assert_array_equal(g2.cpu().numpy(), g0)
    assert torch.equal(hp, hp_before)                    # the caller's rows are not modified
    mb.set_params(default_params(3))


@pytest.mark.parametrize("mem,past", [(1, 0), (3, 3), (16, 8), (8, 0)])
def test_rolled_two_loop_recursion_on_device(handles, oracle, mem, past):
    """mem_size != 8 takes the rolled two-loop recursion (ring indexing through shared alpha / y.s arrays), `past`
    0 / 8 the extremes of the past-cost ring: same control flow as the oracle for a few iterations, and full runs
    converge like it."""
    pb = synth.make_problems(192, N=8, K=16, S=3)
    mb = handles[3]
    mb.set_problems(pb)
    for iters in (2, 5):
        prm = default_params(3, mem_size=mem, past=past, max_iterations=iters)
        mb.set_params(prm)
        res = mb.optimize(pb.x0())
        ref = oracle.optimize_batch(prm, pb, nthreads=8)
        same = (res["evals"] == ref["evals"]) & (res["iters"] == ref["iters"]) & (res["status"] == ref["status"])
        assert same.mean() >= 0.98, (mem, past, iters, same.mean())
        assert rel_rows(res["x"][same], ref["x"][same]) <= 1e-7
    prm = default_params(3, mem_size=mem, past=past, max_iterations=400 if past == 0 else 1000)
    mb.set_params(prm)
    res = mb.optimize(pb.x0())
    ref = oracle.optimize_batch(prm, pb, nthreads=8)
    fo, _ = oracle.cost_batch(prm, pb, res["x"], nthreads=8)
    assert float(np.max(np.abs(res["f"] - fo) / np.abs(fo))) <= TOL
    rel = np.abs(res["f"] - ref["f"]) / np.abs(ref["f"])
    assert np.median(rel) <= 5e-3, np.median(rel)
    mb.set_params(default_params(3))


def test_autograd_layer_on_the_default_stream_large_batch(handles):
    """on torch's default stream (handle 0) the layer must be ordered with the torch kernels that
    produce its inputs and consume its outputs.  A large batch whose inputs are produced by a chain of torch
    kernels right before the call, compared with the same call after a full synchronisation."""
    import torch
    from allocnet_b200.autograd import minco_layer
    dev = torch.device("cuda:0")
    S, N, B = 3, 8, 60000
    g = torch.Generator(device=dev); g.manual_seed(5)
    mb = handles[S]
    head = torch.randn(B, S, 3, dtype=torch.float64, device=dev, generator=g)
    tail = torch.randn(B, S, 3, dtype=torch.float64, device=dev, generator=g)
    base_q = torch.randn(B, N - 1, 3, dtype=torch.float64, device=dev, generator=g)
    base_T = torch.rand(B, N, dtype=torch.float64, device=dev, generator=g) + 0.5
    torch.cuda.synchronize()
    assert torch.cuda.current_stream(dev).cuda_stream == 0

    def run(sync_first):
        q = base_q.clone().requires_grad_(True); T = base_T.clone().requires_grad_(True)
        qq, TT = q, T
        for _ in range(30):                      # a queue of default-stream kernels the layer has to wait for
            qq = qq * 1.0000001 + 1e-9; TT = TT * 1.0000001 + 1e-9
        if sync_first:
            torch.cuda.synchronize()
        e, c = minco_layer(mb, head, tail, qq, TT)
        loss = (e * 1e-3).sum() + (c * c).sum() * 1e-3
        loss.backward()
        if sync_first:
            torch.cuda.synchronize()
        return loss.detach().clone(), q.grad.clone(), T.grad.clone()
    la, gqa, gTa = run(True)
    for _ in range(3):
        lb, gqb, gTb = run(False)
        assert torch.equal(la, lb) and torch.equal(gqa, gqb) and torch.equal(gTa, gTb)
    mb.set_stream(None)


def test_division_free_decisions_match_reference_forms():
    """the shipped build writes the L-BFGS decisions of gcopter/lbfgs.hpp without fp64 divisions
    (rsqrt for 1/|d| :543, products for the quotient tests :531 / :610-614, squared cautious test :655, 1/(y.s) kept
    instead of dividing :676-701).  allocnet_b200/libmincob_strict.so is the same code with those lines written exactly
    as the reference writes them.  On the full 65 536-problem batch, capped at 20 iterations, both builds must take the
    same decisions: identical status / iteration / evaluation counts on >= 99.9 % of the problems (measured: 65 531 of
    65 536), with iterates that differ only by the amplified last-bit differences of the rewritten expressions
    (median and 99th percentile of the row-wise relative difference stated below; 20 L-BFGS iterations on stiff penalty
    terms amplify an ulp by orders of magnitude on a few problems, which is why this is a distribution, not a maximum)."""
    import os
    strict = os.path.join(os.path.dirname(api.LIB_PATH), "libmincob_strict.so")
    assert os.path.exists(strict), "build() makes it (allocnet_b200/build.py::build_strict_library)"
    pb = synth.make_problems(65536, N=8, K=16, S=3)
    prm = default_params(3, max_iterations=20, mapping=P.MAP_THROUGHPUT)
    res = []
    for path in (None, strict):
        mb = api.MincoBatch(prm, device=0, lib_path=path)
        mb.set_problems(pb)
        res.append(mb.optimize(pb.x0(), want_coeffs=False))
        mb.close()
    a, b = res
    same = (a["evals"] == b["evals"]) & (a["iters"] == b["iters"]) & (a["status"] == b["status"])
    assert same.mean() >= 0.999, same.mean()
    rr = np.abs(a["x"][same] - b["x"][same]).max(axis=1) / np.abs(b["x"][same]).max(axis=1)
    rf = np.abs(a["f"][same] - b["f"][same]) / np.abs(b["f"][same])
    stats = (np.median(rr), np.percentile(rr, 99), rr.max(), np.median(rf), np.percentile(rf, 99), rf.max())
    print("strict vs shipped after 20 iterations: same decisions %.5f; x rel median/p99/max %.1e/%.1e/%.1e; f rel %.1e/%.1e/%.1e"
          % ((same.mean(),) + stats))
    assert stats[0] <= 1e-11 and stats[1] <= 1e-7 and stats[3] <= 1e-11 and stats[4] <= 1e-7, stats
    # and full runs of the strict build converge like the shipped one
    prm = default_params(3, mapping=P.MAP_THROUGHPUT)
    sub = pb.slice(0, 4096)
    out = []
    for path in (None, strict):
        mb = api.MincoBatch(prm, device=0, lib_path=path)
        mb.set_problems(sub)
        out.append(mb.optimize(sub.x0(), want_coeffs=False))
        mb.close()
    rel = np.abs(out[0]["f"] - out[1]["f"]) / np.abs(out[1]["f"])
    assert np.median(rel) <= 2e-3 and abs(out[0]["evals"].mean() / out[1]["evals"].mean() - 1.0) <= 0.05


def test_overlapped_upload_gives_the_same_results(handles):
    """mincob_set_problems_async: the batch is uploaded in chunks while the optimize kernel already runs (its work
    queue waits per problem for the arrival counter).  Results are bit-identical to the synchronous upload, also
    when the batch is replaced by another one right away (the second upload must wait for the first kernel), and any
    other entry point called after an asynchronous upload sees the complete batch."""
    mb = handles[3]
    mb.set_params(default_params(3, max_iterations=40))
    B = 20000                                                     # several chunks of 4096
    pbs = [synth.make_problems(B, N=8, K=16, S=3, first=f) for f in (0, 50000)]
    want = []
    for pb in pbs:
        mb.set_problems(pb)
        want.append(mb.optimize(pb.x0()))
    pin = api.pinned_empty

    def pinned(pb):
        q = synth.ProblemBatch(pb.S, pb.N, pb.K, pin(pb.head.shape), pin(pb.tail.shape), pin(pb.hpolys.shape),
                               pin(pb.hrows.shape, np.int32), pb.q0, pb.T0)
        q.head[...] = pb.head; q.tail[...] = pb.tail; q.hpolys[...] = pb.hpolys; q.hrows[...] = pb.hrows
        return q
    pp = [pinned(pb) for pb in pbs]
    for rep in range(2):
        for pb, ref in zip(pp, want):
            mb.set_problems_async(pb)
            got = mb.optimize(pb.x0())
            for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
                np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    mb.set_problems_async(pp[0])
    f, g = mb.evaluate(pp[0].x0())                                 # waits for the whole upload
    mb.set_problems(pbs[0])
    f2, g2 = mb.evaluate(pbs[0].x0())
    np.testing.assert_array_equal(f, f2); np.testing.assert_array_equal(g, g2)
    mb.set_params(default_params(3))


@pytest.mark.parametrize("S,N,K,B", [(3, 5, 16, 12000), (3, 8, 16, 9000), (4, 8, 16, 130), (3, 16, 16, 64), (3, 5, 50, 90), (3, 2, 8, 33), (3, 1, 8, 20)])
def test_fixed_time_kernel_equals_generic_kernel(handles, S, N, K, B, monkeypatch):
    """The fixed-time specialisation (optimize_kernel<.., FRZ = true>: block factorisation computed once per problem,
    no time gradient) against the generic kernel run with DevParams::freeze (MINCOB_NO_FRZ in the environment): every
    output bit-identical, both mappings.  Batches larger than the resident groups, so lane groups fetch new problems
    while their warp-mates are mid-run (the warp then refactorises as a whole)."""
    pb = synth.make_problems(B, N=N, K=K, S=S, ragged_rows=(K == 50))
    mb = handles[S]
    for mp in (P.MAP_THROUGHPUT, P.MAP_LATENCY):
        prm = default_params(S, flags=P.FLAG_FREEZE_TIMES, mapping=mp, max_iterations=60)
        mb.set_params(prm)
        mb.set_problems(pb)
        monkeypatch.delenv("MINCOB_NO_FRZ", raising=False)
        a = mb.optimize(pb.x0())
        monkeypatch.setenv("MINCOB_NO_FRZ", "1")
        b = mb.optimize(pb.x0())
        monkeypatch.delenv("MINCOB_NO_FRZ", raising=False)
        for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"mapping {mp} {k}")
        if N > 1:   # (one piece: no free waypoint, nothing to iterate on in the fixed-time mode)
            assert np.median(a["iters"]) >= 10
    mb.set_params(default_params(S))


@pytest.mark.parametrize("kappa", [1, 5, 6, 7, 13, 31, 32, 40])
def test_integral_intervals(handles, oracle, kappa):
    """IntegralIntervs (kappa sub-intervals, kappa + 1 trapezoid samples per piece) other than the default 16: block
    boundaries of the 6-sample register blocks, a single sample pair, and more samples than the latency mapping has flag
    bits for (kappa > 31: MINCOB_MAP_LATENCY falls back to the throughput mapping).  Cost and gradient against the oracle
    (1e-9), and the two mappings bit-identical on a short optimisation."""
    pb = synth.make_problems(70, N=5, K=16, S=3)
    mb = handles[3]
    prm = default_params(3, kappa=kappa, max_iterations=25)
    mb.set_params(prm)
    mb.set_problems(pb)
    rng = np.random.default_rng(kappa)
    x = pb.x0() + 0.05 * rng.normal(size=pb.x0().shape)
    f, g = mb.evaluate(x)
    fo, go = oracle.cost_batch(prm, pb, x, nthreads=8)
    assert float(np.max(np.abs(f - fo) / np.abs(fo))) <= TOL and rel_rows(g, go) <= TOL
    out = {}
    for mp in (P.MAP_THROUGHPUT, P.MAP_LATENCY):
        prm.mapping = mp
        mb.set_params(prm)
        out[mp] = mb.optimize(pb.x0())
        assert mb.last_mapping() == (mp if kappa <= 31 else P.MAP_THROUGHPUT)
    for k in ("x", "f", "status", "iters", "evals", "coeffs", "T"):
        np.testing.assert_array_equal(out[P.MAP_THROUGHPUT][k], out[P.MAP_LATENCY][k], err_msg=k)
    fo2, _ = oracle.cost_batch(prm, pb, out[P.MAP_LATENCY]["x"], nthreads=8)
    assert float(np.max(np.abs(out[P.MAP_LATENCY]["f"] - fo2) / np.abs(fo2))) <= TOL
    mb.set_params(default_params(3))
