"""Pins for the MINCO / cost-functional part of the CPU oracle (oracle/minco_oracle.hpp).

The reference has no MINCO source and no tests for this path, so the
oracle is pinned by independent restatements written here:
  (1) banded LU / solve / solveAdj vs numpy dense solves,
  (2) explicit S=3 rows typed from SURVEY.md Appendix A.2 vs the oracle's general-S rule,
  (3) KKT of the REFERENCE QP formulation reproduces the MINCO coefficients -- with Q, A, b, G, h taken from
      tests/golden/ref_qp_*.npz, i.e. built by the reference's own network/utils/min_traj_opt.py
      (fill_eq_obj :377-533, fill_ineq :535-607; twin of planner/qp_solver.hpp:148-296), not typed here,
  (4) E == 2 * Trajectory<5>::getTrajCost(3) (gcopter/trajectory.hpp:396-420 constants),
  (5) finite differences on propogateGrad and on the whole cost functional,
  (6) smoothedL1 (gcopter/firi.hpp:60-84) and the tau<->T map properties,
  (7) an mpmath 40-digit solve on a few seeds.
"""
import math

import numpy as np
import pytest

from allocnet_b200 import synth
from allocnet_b200.params import default_params, energy_only


def fact(d):
    return float(math.factorial(d))


def beta(t, d, D):
    out = np.zeros(D)
    for k in range(d, D):
        out[k] = fact(k) / fact(k - d) * t ** (k - d)
    return out


def dense_minco(S, head, tail, q, T):
    """Independent dense build of A, b by the general rule of SURVEY.md Appendix A.2."""
    D, N = 2 * S, len(T)
    A = np.zeros((D * N, D * N)); b = np.zeros((D * N, 3))
    for d in range(S):
        A[d, d] = fact(d); b[d] = head[d]
    for i in range(N - 1):
        r, c0 = D * i + S, D * i
        for j in range(S - 1):
            d = S + j
            A[r + j, c0:c0 + D] = beta(T[i], d, D); A[r + j, c0 + D + d] = -fact(d)
        A[r + S - 1, c0:c0 + D] = beta(T[i], 0, D); b[r + S - 1] = q[i]
        for d in range(S):
            A[r + S + d, c0:c0 + D] = beta(T[i], d, D); A[r + S + d, c0 + D + d] = -fact(d)
    for d in range(S):
        A[D * N - S + d, D * (N - 1):] = beta(T[N - 1], d, D); b[D * N - S + d] = tail[d]
    return A, b


def rand_problem(rng, S, N):
    head = rng.normal(size=(S, 3)); tail = rng.normal(size=(S, 3))
    q = np.cumsum(rng.normal(size=(max(N - 1, 0), 3)), axis=0)
    T = rng.uniform(0.6, 2.0, size=N)
    return head, tail, q, T


def test_banded_matches_dense(oracle):
    rng = np.random.default_rng(1)
    for n, p, q in [(12, 3, 2), (48, 6, 6), (64, 8, 8), (5, 1, 1)]:
        A = np.zeros((n, n))
        for i in range(n):
            for j in range(max(0, i - p), min(n, i + q + 1)):
                A[i, j] = rng.normal()
            A[i, i] += 6.0
        b = rng.normal(size=(n, 3))
        np.testing.assert_allclose(oracle.banded_solve(A, p, q, b), np.linalg.solve(A, b), rtol=0, atol=1e-12)
        np.testing.assert_allclose(oracle.banded_solve(A, p, q, b, adj=True), np.linalg.solve(A.T, b), rtol=0, atol=1e-12)


def test_s3_rows_typed_from_appendix_a2():
    """The general-S rule must produce exactly the S=3 entries listed in SURVEY.md Appendix A.2."""
    rng = np.random.default_rng(2)
    head, tail, q, T = rand_problem(rng, 3, 4)
    A, b = dense_minco(3, head, tail, q, T)
    t = T[1]; r = 6
    T1, T2, T3, T4, T5 = t, t**2, t**3, t**4, t**5
    exp = np.zeros((6, 24))
    exp[0, r + 3], exp[0, r + 4], exp[0, r + 5], exp[0, r + 9] = 6, 24 * T1, 60 * T2, -6
    exp[1, r + 4], exp[1, r + 5], exp[1, r + 10] = 24, 120 * T1, -24
    exp[2, r:r + 6] = [1, T1, T2, T3, T4, T5]
    exp[3, r:r + 6] = [1, T1, T2, T3, T4, T5]; exp[3, r + 6] = -1
    exp[4, r + 1:r + 6] = [1, 2 * T1, 3 * T2, 4 * T3, 5 * T4]; exp[4, r + 7] = -1
    exp[5, r + 2:r + 6] = [2, 6 * T1, 12 * T2, 20 * T3]; exp[5, r + 8] = -2
    np.testing.assert_allclose(A[r + 3:r + 9], exp, rtol=1e-15)
    assert A[0, 0] == 1 and A[1, 1] == 1 and A[2, 2] == 2
    np.testing.assert_array_equal(b[r + 5], q[1])


@pytest.mark.parametrize("S,N", [(3, 1), (3, 2), (3, 5), (3, 8), (3, 16), (4, 2), (4, 8), (4, 16)])
def test_coefficients_match_dense_solve(oracle, S, N):
    rng = np.random.default_rng(10 * S + N)
    head, tail, q, T = rand_problem(rng, S, N)
    out = oracle.minco_forward(S, head, tail, q, T)
    A, b = dense_minco(S, head, tail, q, T)
    c = np.linalg.solve(A, b)
    np.testing.assert_allclose(out["coeffs"], c, rtol=0, atol=2e-9 * max(1.0, np.abs(c).max()))
    # band half-width is 2S
    ii, jj = np.nonzero(A)
    assert np.max(np.abs(ii - jj)) <= 2 * S
    # Trajectory packing: [piece][axis][k], k=0 highest power (trajectory.hpp:79-83)
    D = 2 * S
    for i in range(N):
        for a in range(3):
            np.testing.assert_array_equal(out["flat"][i, a], out["coeffs"][D * i:D * i + D, a][::-1])


def test_energy_s3_explicit_and_trajcost_identity(oracle):
    rng = np.random.default_rng(3)
    head, tail, q, T = rand_problem(rng, 3, 5)
    out = oracle.minco_forward(3, head, tail, q, T)
    c = out["coeffs"].reshape(5, 6, 3)
    E = 0.0; half = 0.0
    for i in range(5):
        t = T[i]; c3, c4, c5 = c[i, 3], c[i, 4], c[i, 5]
        E += (36 * c3 @ c3 * t + 144 * c4 @ c3 * t**2 + 192 * c4 @ c4 * t**3 + 240 * c5 @ c3 * t**3
              + 720 * c5 @ c4 * t**4 + 720 * c5 @ c5 * t**5)
        # Trajectory<5>::getTrajCost(3): 0.5 z^T Q z, z = leading 3 (descending) coeffs; trajectory.hpp:396-420
        Q = np.array([[720 * t**5, 360 * t**4, 120 * t**3], [360 * t**4, 192 * t**3, 72 * t**2],
                      [120 * t**3, 72 * t**2, 36 * t]])
        for a in range(3):
            z = out["flat"][i, a, :3]
            half += 0.5 * z @ Q @ z
    assert out["energy"] == pytest.approx(E, rel=1e-13)
    assert out["energy"] == pytest.approx(2.0 * half, rel=1e-13)


# ---- reference-executed pins: matrices built by the reference's own fill_eq_obj / fill_ineq -----------------
# tests/golden/ref_qp_*.npz are written by tests/golden/make_ref_qp_fixtures.py, which imports
# /root/reference/network/utils/min_traj_opt.py unmodified (the Python twin of planner/qp_solver.hpp:119-360)
# and stores the Q, A, b, G1, h1, G2, h2 it builds.  No reference constant is typed into this file.
from ref_qp_util import REF_QP_CASES, load_ref_qp, ref_qp_problem, solve_ld, ref_qp_kkt  # noqa: E402


@pytest.mark.parametrize("name", REF_QP_CASES)
def test_reference_qp_matrices_reproduce_minco(oracle, name):
    """min 1/2 z^T Q z  s.t.  A z = b (the reference's boundary + continuity rows) and waypoint rows
    == MINCO_S3NU coefficients of the oracle, in the reference's flatten order idx = i*3*d + j*d + k."""
    c = load_ref_qp(name)
    head, tail, q, T = ref_qp_problem(c, 5)
    N, d = len(T), 6
    Q, A, b = c["Q"], c["A"], c["b"]
    nv = N * 3 * d
    assert Q.shape == (nv, nv) and A.shape[1] == nv
    W = np.zeros((3 * (N - 1), nv)); wq = np.zeros(3 * (N - 1))
    for i in range(N - 1):
        for a in range(3):      # start position of piece i+1, selected with the reference's own zero_A row 0
            W[3 * i + a, (i + 1) * 3 * d + a * d:(i + 1) * 3 * d + a * d + d] = c["zero_A"][0]
            wq[3 * i + a] = q[i, a]
    Cm = np.vstack([A, W]); rhs = np.concatenate([b, wq]); m = Cm.shape[0]
    KKT = np.block([[Q, Cm.T], [Cm, np.zeros((m, m))]])
    z = solve_ld(KKT, np.concatenate([np.zeros(nv), rhs]))[:nv]
    out = oracle.minco_forward(3, head, tail, q, T)
    zo = out["flat"].reshape(-1)
    np.testing.assert_allclose(zo, z, rtol=0, atol=1e-9 * np.abs(z).max())
    assert 0.5 * zo @ Q @ zo == pytest.approx(out["energy"] / 2.0, rel=1e-10)
    # the oracle's coefficients satisfy the reference's equality rows
    assert np.max(np.abs(A @ zo - b)) <= 1e-9 * max(1.0, np.abs(b).max(), np.abs(zo).max())


@pytest.mark.parametrize("name", REF_QP_CASES)
def test_reference_qp_sensitivities_pin_energy_partials_and_propagate_grad(oracle, name):
    """getEnergyPartialGradByCoeffs == 2 Q z, getEnergyPartialGradByTimes == z^T dQ/dT z, and propogateGrad ==
    the KKT sensitivities of the reference's equality-constrained QP: with L = 1/2 z^T Q z + lam^T (C z - r),
    d(1/2 z^T Q z)*/dT_i = 1/2 z^T dQ_i z + lam^T dC_i z and d/dq = -lam_waypoint.  dQ_i, dC_i are autograd
    Jacobians through the reference's fill_eq_obj (fixture); E = 2 * (1/2 z^T Q z)."""
    c = load_ref_qp(name)
    head, tail, q, T = ref_qp_problem(c, 7)
    N, d = len(T), 6
    Q, A, b, dQ, dA = c["Q"], c["A"], c["b"], c["dQ"], c["dA"]
    nv = N * 3 * d
    W = np.zeros((3 * (N - 1), nv)); wq = np.zeros(3 * (N - 1))
    for i in range(N - 1):
        for a in range(3):
            W[3 * i + a, (i + 1) * 3 * d + a * d:(i + 1) * 3 * d + a * d + d] = c["zero_A"][0]
            wq[3 * i + a] = q[i, a]
    Cm = np.vstack([A, W]); rhs = np.concatenate([b, wq]); m = Cm.shape[0]
    KKT = np.block([[Q, Cm.T], [Cm, np.zeros((m, m))]])
    sol = solve_ld(KKT, np.concatenate([np.zeros(nv), rhs]))
    z, lam = sol[:nv], sol[nv:]
    out = oracle.minco_forward(3, head, tail, q, T)
    zo = out["flat"].reshape(-1)
    # partials at the oracle's own coefficients; flat[i][a][k] = coeffs[6i + 5 - k][a]
    gdC_flat = out["gdC"].reshape(N, 6, 3)[:, ::-1, :].transpose(0, 2, 1).reshape(-1)
    np.testing.assert_allclose(gdC_flat, 2.0 * Q @ zo, rtol=0, atol=1e-10 * np.abs(Q @ zo).max())
    gdT_ref = np.array([zo @ dQ[i] @ zo for i in range(N)])
    np.testing.assert_allclose(out["gdT"], gdT_ref, rtol=1e-10, atol=1e-10 * np.abs(gdT_ref).max())
    # total derivatives of the optimal energy
    gq, gT = oracle.minco_propagate(3, head, tail, q, T, out["gdC"], out["gdT"])
    gT_ref = np.array([z @ dQ[i] @ z + 2.0 * lam[:A.shape[0]] @ (dA[i] @ z) for i in range(N)])
    gq_ref = -2.0 * lam[A.shape[0]:].reshape(N - 1, 3)
    np.testing.assert_allclose(gT, gT_ref, rtol=0, atol=2e-9 * np.abs(gT_ref).max())
    np.testing.assert_allclose(gq, gq_ref, rtol=0, atol=2e-9 * np.abs(gq_ref).max())


def _cpp_qp(oracle, c):
    st, hp = c["state"], c["hpolys"]
    ini = st[:, 0].reshape(3, 3); fin = st[:, 1].reshape(3, 3)            # rows axis, columns P,V,A
    return oracle.ref_qp_build(ini, fin, hp.transpose(2, 0, 1), c["times"].astype(np.float32), order=3, res=int(c["res"]))


@pytest.mark.parametrize("name", REF_QP_CASES)
def test_reference_cpp_qp_solver_builds_the_matrices_of_its_python_twin(oracle, name):
    """planner/qp_solver.hpp, compiled VERBATIM (oracle/_ref/libref_qp.so: ROS / OsqpEigen / Eigen stand-ins that record what
    QPSolver::solve hands to OSQP), builds the same Q, A, b, G, h as the reference's Python twin did for the fixtures --
    up to the float32 arithmetic the C++ uses for the powers of the durations (qp_solver.hpp:183-184,252; SURVEY Appendix C
    quirk Q2): the fixtures that pin MINCO are the C++ back-end's own problem."""
    c = load_ref_qp(name)
    q = _cpp_qp(oracle, c)
    if q is None:
        pytest.skip("oracle/_ref/libref_qp.so not built (no reference tree on this box)")
    meq, K, N, res = q["n_eq"], c["hpolys"].shape[0], len(c["times"]), int(c["res"])
    f32 = 3e-7
    np.testing.assert_allclose(q["hessian"], c["Q"], rtol=0, atol=f32 * np.abs(c["Q"]).max())
    np.testing.assert_allclose(q["constraints"][:meq], c["A"], rtol=0, atol=f32 * np.abs(c["A"]).max())
    np.testing.assert_array_equal(q["upper"][:meq], c["b"]); np.testing.assert_array_equal(q["lower"][:meq], c["b"])
    rows, hs = [], []
    for i in range(N):                        # the C++ interleaves corridor and box rows per sample, the twin keeps two arrays
        for j in range(res):
            s0 = i * res + j
            rows += [c["G1"][s0 * K:(s0 + 1) * K], c["G2"][s0 * 12:(s0 + 1) * 12]]
            hs += [c["h1"][s0 * K:(s0 + 1) * K], c["h2"][s0 * 12:(s0 + 1) * 12]]
    G, h = np.vstack(rows), np.concatenate(hs)
    np.testing.assert_allclose(q["constraints"][meq:], G, rtol=0, atol=f32 * np.abs(G).max())
    np.testing.assert_array_equal(q["upper"][meq:], h)
    assert np.isneginf(q["lower"][meq:]).all()


@pytest.mark.parametrize("name", REF_QP_CASES)
def test_reference_cpp_qp_matrices_reproduce_minco(oracle, name):
    """Same KKT statement as test_reference_qp_matrices_reproduce_minco, with the matrices of the C++ back-end itself and
    its float32 durations: the minimiser is the oracle's MINCO_S3NU at T = float32(times), to what float32 time powers in
    Q and A allow."""
    c = load_ref_qp(name)
    q = _cpp_qp(oracle, c)
    if q is None:
        pytest.skip("oracle/_ref/libref_qp.so not built (no reference tree on this box)")
    head, tail, wp, _ = ref_qp_problem(c, 5)
    T = c["times"].astype(np.float32).astype(np.float64)
    c2 = dict(c); c2["Q"] = q["hessian"]; c2["A"] = q["constraints"][:q["n_eq"]]; c2["b"] = q["upper"][:q["n_eq"]]
    z, _, _ = ref_qp_kkt(c2, wp)
    out = oracle.minco_forward(3, head, tail, wp, T)
    zo = out["flat"].reshape(-1)
    np.testing.assert_allclose(zo, z, rtol=0, atol=2e-5 * np.abs(z).max())
    assert 0.5 * zo @ q["hessian"] @ zo == pytest.approx(out["energy"] / 2.0, rel=2e-5)     # float32 entries of Q


@pytest.mark.parametrize("name", REF_QP_CASES)
def test_reference_qp_inequality_rows_pin_sampling_layout(oracle, name):
    """G1 z - h1 (corridor) and G2 z - h2 (+-v, +-a boxes) of the reference, evaluated on the oracle's
    coefficients, equal the residuals computed from the polynomial pieces at t = j*T_i/res, j = 0..res-1
    (left end points, qp_solver.hpp:252-296 / min_traj_opt.py:553-607)."""
    c = load_ref_qp(name)
    head, tail, q, T = ref_qp_problem(c, 6)
    N, res = len(T), int(c["res"])
    out = oracle.minco_forward(3, head, tail, q, T)
    zo = out["flat"].reshape(-1)
    co = out["coeffs"].reshape(N, 6, 3)               # ascending powers
    r1 = c["G1"] @ zo - c["h1"]; r2 = c["G2"] @ zo - c["h2"]
    e1, e2 = [], []
    for i in range(N):
        hp = c["hpolys"][:, :, i]
        for j in range(res):
            t = j * T[i] / res
            p = beta(t, 0, 6) @ co[i]; v = beta(t, 1, 6) @ co[i]; a = beta(t, 2, 6) @ co[i]
            e1.extend(hp[:, :3] @ p - hp[:, 3])
            for ax in range(3):
                e2.extend([v[ax] - 4.0, a[ax] - 6.0, -v[ax] - 4.0, -a[ax] - 6.0])
    scale = max(1.0, np.abs(zo).max())
    np.testing.assert_allclose(r1, np.array(e1), rtol=0, atol=1e-10 * scale)
    np.testing.assert_allclose(r2, np.array(e2), rtol=0, atol=1e-10 * scale)


@pytest.mark.parametrize("S,N", [(3, 2), (3, 5), (3, 8), (4, 8), (3, 16)])
def test_propagate_grad_finite_differences(oracle, S, N):
    rng = np.random.default_rng(100 + S * 17 + N)
    head, tail, q, T = rand_problem(rng, S, N)
    W = rng.normal(size=(2 * S * N, 3))  # J = E + <W, c>

    def J(qq, TT):
        o = oracle.minco_forward(S, head, tail, qq, TT)
        return o["energy"] + np.sum(W * o["coeffs"])
    o = oracle.minco_forward(S, head, tail, q, T)
    gq, gT = oracle.minco_propagate(S, head, tail, q, T, o["gdC"] + W, o["gdT"])
    h = 1e-4  # 5-point stencil: truncation O(h^4), round-off ~ eps*|J|*cond/h

    def fd5(fun):
        return (8.0 * (fun(h) - fun(-h)) - (fun(2 * h) - fun(-2 * h))) / (12.0 * h)
    for i in range(N):
        def along(e, i=i):
            TT = T.copy(); TT[i] += e
            return J(q, TT)
        fd = fd5(along)
        assert gT[i] == pytest.approx(fd, rel=5e-6, abs=5e-6 * max(1.0, np.abs(gT).max()))
    for i in range(N - 1):
        for a in range(3):
            def along(e, i=i, a=a):
                qq = q.copy(); qq[i, a] += e
                return J(qq, T)
            fd = fd5(along)
            assert gq[i, a] == pytest.approx(fd, rel=5e-6, abs=5e-6 * max(1.0, np.abs(gq).max()))


@pytest.mark.parametrize("S,N,K", [(3, 5, 16), (3, 8, 16), (4, 8, 16), (3, 8, 0)])
def test_cost_functional_finite_differences(oracle, S, N, K):
    p = default_params(S)
    if K == 0:
        p = energy_only(p)
    pb = synth.make_problems(3, N, K, S, time_scale=0.8)  # short times: many active hinges
    x = pb.x0()
    for b in range(pb.B):
        ci = oracle.cost_instance(p, pb, b)
        f0, g0 = ci(x[b])
        h = 1e-7
        scale = max(1.0, np.abs(g0).max())
        for i in range(x.shape[1]):
            xp = x[b].copy(); xp[i] += h; xm = x[b].copy(); xm[i] -= h
            fd = (ci(xp)[0] - ci(xm)[0]) / (2 * h)
            assert abs(fd - g0[i]) <= 5e-6 * scale, (b, i, fd, g0[i])
        ci.close()


def test_smoothed_l1_and_time_map(oracle):
    mu = 1e-2
    assert oracle.smoothed_l1(mu, -1e-3) == (False, 0.0, 0.0)
    hit, f, df = oracle.smoothed_l1(mu, 0.5)
    assert hit and f == pytest.approx(0.5 - 0.5 * mu) and df == 1.0
    hit, f, df = oracle.smoothed_l1(mu, 0.5 * mu)   # (mu - x/2)(x/mu)^3
    assert hit and f == pytest.approx((mu - 0.25 * mu) * 0.125) and df == pytest.approx(0.25 * (-0.25 + 3 * 0.75))
    # C1 at x = mu
    _, f1, d1 = oracle.smoothed_l1(mu, mu * (1 - 1e-9)); _, f2, d2 = oracle.smoothed_l1(mu, mu * (1 + 1e-9))
    assert f1 == pytest.approx(f2, abs=1e-10) and d1 == pytest.approx(d2, abs=1e-7)
    L = oracle.lib
    for tau in [-3.0, -0.5, 0.0, 0.3, 2.5]:
        T = L.orc_forward_t(tau)
        assert T > 0 and L.orc_backward_t(T) == pytest.approx(tau, abs=1e-12)
        h = 1e-6
        fd = (L.orc_forward_t(tau + h) - L.orc_forward_t(tau - h)) / (2 * h)
        assert L.orc_backward_grad_t(tau, 1.0) == pytest.approx(fd, rel=1e-8)
    np.testing.assert_allclose(synth.forward_t(synth.backward_t(np.array([0.2, 1.0, 3.0]))), [0.2, 1.0, 3.0], rtol=1e-14)


def test_mpmath_certifies_fp64_oracle(oracle):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(7)
    for S, N in [(3, 8), (4, 5)]:
        head, tail, q, T = rand_problem(rng, S, N)
        A, b = dense_minco(S, head, tail, q, T)  # entries are exact products/powers of doubles up to 1 ulp
        D = 2 * S
        Am = mp.matrix(D * N, D * N)
        for d in range(S):
            Am[d, d] = mp.factorial(d)
        def mbeta(t, d):
            return [mp.factorial(k) / mp.factorial(k - d) * mp.mpf(t) ** (k - d) if k >= d else mp.mpf(0) for k in range(D)]
        for i in range(N - 1):
            r, c0 = D * i + S, D * i
            for j in range(S - 1):
                d = S + j
                for k, v in enumerate(mbeta(T[i], d)): Am[r + j, c0 + k] = v
                Am[r + j, c0 + D + d] = -mp.factorial(d)
            for k, v in enumerate(mbeta(T[i], 0)): Am[r + S - 1, c0 + k] = v
            for d in range(S):
                for k, v in enumerate(mbeta(T[i], d)): Am[r + S + d, c0 + k] = v
                Am[r + S + d, c0 + D + d] = -mp.factorial(d)
        for d in range(S):
            for k, v in enumerate(mbeta(T[N - 1], d)): Am[D * N - S + d, D * (N - 1) + k] = v
        c = np.zeros((D * N, 3))
        for a in range(3):
            col = mp.lu_solve(Am, mp.matrix(b[:, a].tolist()))
            c[:, a] = [float(col[i]) for i in range(D * N)]
        out = oracle.minco_forward(S, head, tail, q, T)
        assert np.max(np.abs(out["coeffs"] - c)) <= 1e-10 * max(1.0, np.abs(c).max())


def test_smoothed_l1_equals_reference_function(oracle_strict):
    """firi.hpp:60-84 itself (cut out of the reference file by oracle/Makefile and compiled unmodified into
    oracle/_ref) against the oracle's restatement: bit-equal on a grid that covers x<0, the quartic blend and
    the linear branch, for several mu."""
    import ctypes as C
    ref = oracle_strict.ref
    if ref is None or not hasattr(ref, "ref_smoothed_l1"):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref.ref_smoothed_l1.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(5)
    for mu in (1e-2, 1e-3, 0.5):
        xs = np.concatenate([np.linspace(-2 * mu, 3 * mu, 401), rng.uniform(-mu, 2 * mu, 500), [0.0, mu, np.nextafter(mu, 1)]])
        for x in xs:
            f, df = C.c_double(-7.0), C.c_double(-7.0)
            hit = ref.ref_smoothed_l1(mu, float(x), C.byref(f), C.byref(df))
            h2, f2, df2 = oracle_strict.smoothed_l1(mu, float(x))
            assert bool(hit) == h2
            if hit:
                assert f.value == f2 and df.value == df2, (mu, x)


def test_output_contract_against_reference_trajectory(oracle):
    """The reference's own Piece<5>/Trajectory<5> (gcopter/trajectory.hpp, compiled verbatim into oracle/_ref)
    fed with getTrajectory-order coefficients exactly as learning_planner.hpp:205-216 feeds it:
    getPos/getVel/getAcc/getJer reproduce the spline (descending powers, trajectory.hpp:79-133),
    getPositions returns head, waypoints, tail, locatePieceIdx walks the durations, and the reference's
    energy formula gives E_MINCO == 2 * getTrajCost(3) (trajectory.hpp:396-420)."""
    import ctypes as C
    ref = oracle.ref
    if ref is None or not hasattr(ref, "ref_traj5_create"):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    dp = C.POINTER(C.c_double)
    ref.ref_traj5_create.restype = C.c_void_p
    ref.ref_traj5_create.argtypes = [C.c_int, dp, dp]
    ref.ref_traj5_destroy.argtypes = [C.c_void_p]
    ref.ref_traj5_cost.restype = C.c_double; ref.ref_traj5_cost.argtypes = [C.c_void_p, C.c_int]
    ref.ref_traj5_total_duration.restype = C.c_double; ref.ref_traj5_total_duration.argtypes = [C.c_void_p]
    ref.ref_traj5_eval.argtypes = [C.c_void_p, C.c_double, dp, dp, dp, dp]
    ref.ref_traj5_positions.argtypes = [C.c_void_p, dp]
    ref.ref_traj5_locate.argtypes = [C.c_void_p, dp]
    rng = np.random.default_rng(11)
    for N in (1, 2, 5, 8):
        head = rng.normal(size=(3, 3)); tail = rng.normal(size=(3, 3))
        q = np.cumsum(rng.normal(size=(max(N - 1, 1), 3)), axis=0)[: N - 1]
        T = rng.uniform(0.5, 2.5, size=N)
        out = oracle.minco_forward(3, head, tail, q, T)
        flat = np.ascontiguousarray(out["flat"]); Tc = np.ascontiguousarray(T)
        h = ref.ref_traj5_create(N, Tc.ctypes.data_as(dp), flat.ctypes.data_as(dp))
        try:
            assert abs(2.0 * ref.ref_traj5_cost(h, 3) - out["energy"]) <= 1e-12 * abs(out["energy"])
            assert abs(ref.ref_traj5_total_duration(h) - T.sum()) <= 1e-14 * T.sum()
            P = np.zeros((N + 1, 3)); ref.ref_traj5_positions(h, P.ctypes.data_as(dp))
            np.testing.assert_allclose(P[0], head[0], atol=1e-12)
            np.testing.assert_allclose(P[1:N], q, atol=1e-9)
            np.testing.assert_allclose(P[N], tail[0], atol=1e-9)
            c = out["coeffs"].reshape(N, 6, 3)                     # ascending powers, row k = c_k
            t0 = np.concatenate([[0.0], np.cumsum(T)])
            for t in rng.uniform(0.0, T.sum(), size=12):
                tl = C.c_double(t); i = ref.ref_traj5_locate(h, C.byref(tl))
                assert i == min(np.searchsorted(t0, t, side="left") - 1, N - 1) or t == 0.0
                s_ = t - t0[i]
                assert abs(tl.value - s_) <= 1e-12
                got = [np.zeros(3) for _ in range(4)]
                ref.ref_traj5_eval(h, t, *[g.ctypes.data_as(dp) for g in got])
                for d, g in enumerate(got):
                    fac = np.array([np.prod([k - u for u in range(d)]) if k >= d else 0.0 for k in range(6)])
                    want = (c[i] * (fac * s_ ** np.maximum(np.arange(6) - d, 0))[:, None]).sum(axis=0)
                    np.testing.assert_allclose(g, want, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(want).max()))
        finally:
            ref.ref_traj5_destroy(h)


def _dense_rate_maxima(coeffs, T, pts=20001):
    """max |p'|, |p''|, |p'''| per trajectory by dense sampling (coeffs [N][3][6] descending, T [N])."""
    out = np.zeros(3)
    for i in range(coeffs.shape[0]):
        t = np.linspace(0.0, T[i], pts)
        for d in (1, 2, 3):
            v = np.stack([np.polyval(np.polyder(coeffs[i, a], d), t) for a in range(3)])
            out[d - 1] = max(out[d - 1], np.sqrt((v * v).sum(axis=0).max()))
    return out


def test_reference_max_rates_against_dense_sampling(oracle):
    """The reference's own Trajectory<5>::getMaxVelRate / getMaxAccRate / checkMaxVelRate / checkMaxAccRate
    (gcopter/trajectory.hpp:177-313, 598-646 and gcopter/root_finder.hpp, both compiled verbatim into oracle/_ref
    against the Eigen stand-in) on optimized trajectories: equal to a dense sampling of the same polynomials, and the
    check members flip exactly around that maximum.  This is the oracle of mincob_max_rates (GPU test)."""
    import ctypes as C
    ref = oracle.ref
    if ref is None or not hasattr(ref, "ref_traj5_max_vel_rate"):
        pytest.skip("oracle/_ref (with root_finder.hpp) not present")
    dp = C.POINTER(C.c_double)
    ref.ref_traj5_create.restype = C.c_void_p; ref.ref_traj5_create.argtypes = [C.c_int, dp, dp]
    ref.ref_traj5_destroy.argtypes = [C.c_void_p]
    for fn in (ref.ref_traj5_max_vel_rate, ref.ref_traj5_max_acc_rate):
        fn.restype = C.c_double; fn.argtypes = [C.c_void_p]
    for fn in (ref.ref_traj5_check_max_vel_rate, ref.ref_traj5_check_max_acc_rate):
        fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_double]
    B, N = 24, 5
    pb = synth.make_problems(B, N=N, K=16, S=3)
    res = oracle.optimize_batch(default_params(3), pb, nthreads=4)
    for b in range(B):
        c = np.ascontiguousarray(res["coeffs"][b]); T = np.ascontiguousarray(res["T"][b])
        want = _dense_rate_maxima(c, T)
        h = ref.ref_traj5_create(N, T.ctypes.data_as(dp), c.ctypes.data_as(dp))
        try:
            v, a = ref.ref_traj5_max_vel_rate(h), ref.ref_traj5_max_acc_rate(h)
            assert abs(v - want[0]) <= 1e-7 * want[0] and abs(a - want[1]) <= 1e-7 * want[1], (v, a, want)
            assert ref.ref_traj5_check_max_vel_rate(h, v * 1.001) == 1 and ref.ref_traj5_check_max_vel_rate(h, v * 0.999) == 0
            assert ref.ref_traj5_check_max_acc_rate(h, a * 1.001) == 1 and ref.ref_traj5_check_max_acc_rate(h, a * 0.999) == 0
        finally:
            ref.ref_traj5_destroy(h)


@pytest.mark.parametrize("S,N", [(3, 5), (4, 4), (3, 1)])
def test_torch_reference_equals_oracle(oracle, S, N):
    """tests/torch_minco_ref.py (the torch.autograd reference of the differentiable layer) reproduces the oracle's
    coefficients, energy and -- through autograd -- getEnergyPartialGradByTimes + propogateGrad of the energy."""
    import torch
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from torch_minco_ref import torch_dense_minco
    rng = np.random.default_rng(11)
    head, tail, q, T = rand_problem(rng, S, N)
    ref = oracle.minco_forward(S, head, tail, q, T)
    tq = torch.tensor(q, dtype=torch.float64).requires_grad_(N > 1); tT = torch.tensor(T, dtype=torch.float64).requires_grad_(True)
    E, c = torch_dense_minco(S, torch.tensor(head), torch.tensor(tail), tq, tT)
    tol = 1e-9 if S == 3 else 1e-7
    assert abs(float(E.detach()) - ref["energy"]) <= tol * abs(ref["energy"])
    assert np.abs(c.detach().numpy() - ref["coeffs"]).max() <= tol * np.abs(ref["coeffs"]).max()
    grads = torch.autograd.grad(E, [tT] + ([tq] if N > 1 else []))
    gq, gT = oracle.minco_propagate(S, head, tail, q, T, ref["gdC"], ref["gdT"])
    assert np.abs(grads[0].numpy() - gT).max() <= 10 * tol * np.abs(gT).max()
    if N > 1:
        assert np.abs(grads[1].numpy() - gq).max() <= 10 * tol * np.abs(gq).max()


def test_output_contract_s4_against_reference_trajectory7(oracle):
    """MINCO_S4NU (septic pieces) into the reference's Trajectory<7> (trajectory.hpp compiled verbatim): positions and
    derivative evaluation as for S = 3, and the snap energy.  trajectory.hpp:385 carries the reference's known typo
    (m_34 = 1400 where the integral gives 1440), so 2 * getTrajCost(4) misses the true
    int |p''''|^2 by exactly sum over pieces and axes of 80 T^2 c5 c4 -- which pins layout AND documents the quirk."""
    import ctypes as C
    ref = oracle.ref
    if ref is None or not hasattr(ref, "ref_traj7_create"):
        pytest.skip("oracle/_ref (Trajectory<7>) not built")
    dp = C.POINTER(C.c_double)
    ref.ref_traj7_create.restype = C.c_void_p; ref.ref_traj7_create.argtypes = [C.c_int, dp, dp]
    ref.ref_traj7_destroy.argtypes = [C.c_void_p]
    ref.ref_traj7_cost.restype = C.c_double; ref.ref_traj7_cost.argtypes = [C.c_void_p, C.c_int]
    ref.ref_traj7_eval.argtypes = [C.c_void_p, C.c_double, dp, dp, dp, dp]
    ref.ref_traj7_positions.argtypes = [C.c_void_p, dp]
    rng = np.random.default_rng(12)
    for N in (1, 3, 6):
        head = rng.normal(size=(4, 3)); tail = rng.normal(size=(4, 3))
        q = np.cumsum(rng.normal(size=(max(N - 1, 1), 3)), axis=0)[: N - 1]
        T = rng.uniform(0.6, 2.0, size=N)
        out = oracle.minco_forward(4, head, tail, q, T)
        flat = np.ascontiguousarray(out["flat"]); Tc = np.ascontiguousarray(T)     # [N][3][8], k = 0 highest power
        h = ref.ref_traj7_create(N, Tc.ctypes.data_as(dp), flat.ctypes.data_as(dp))
        try:
            c = out["coeffs"].reshape(N, 8, 3)                                    # ascending powers
            quirk = sum(80.0 * T[i] ** 2 * float((c[i, 5] * c[i, 4]).sum()) for i in range(N))
            assert abs(2.0 * ref.ref_traj7_cost(h, 4) + quirk - out["energy"]) <= 1e-10 * abs(out["energy"])
            assert abs(quirk) > 1e-9 * abs(out["energy"])                          # the typo is really there
            P = np.zeros((N + 1, 3)); ref.ref_traj7_positions(h, P.ctypes.data_as(dp))
            np.testing.assert_allclose(P[0], head[0], atol=1e-12)
            np.testing.assert_allclose(P[1:N], q, atol=1e-8)
            np.testing.assert_allclose(P[N], tail[0], atol=1e-7)
            t0 = np.concatenate([[0.0], np.cumsum(T)])
            for t in rng.uniform(0.0, T.sum(), size=8):
                i = min(int(np.searchsorted(t0, t, side="left")) - 1, N - 1); i = max(i, 0)
                s_ = t - t0[i]
                got = [np.zeros(3) for _ in range(4)]
                ref.ref_traj7_eval(h, t, *[g.ctypes.data_as(dp) for g in got])
                for d, g in enumerate(got):
                    fac = np.array([np.prod([k - u for u in range(d)]) if k >= d else 0.0 for k in range(8)])
                    want = (c[i] * (fac * s_ ** np.maximum(np.arange(8) - d, 0))[:, None]).sum(axis=0)
                    np.testing.assert_allclose(g, want, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(want).max()))
        finally:
            ref.ref_traj7_destroy(h)


def test_reference_max_rates_septic_against_dense_sampling(oracle):
    """Same pin for the reference's Trajectory<7> (MINCO_S4NU): getMaxVelRate / getMaxAccRate through root_finder.hpp
    (degree-12 / degree-10 rate polynomials, Sturm isolation) equal a dense sampling."""
    import ctypes as C
    ref = oracle.ref
    if ref is None or not hasattr(ref, "ref_traj7_max_vel_rate"):
        pytest.skip("oracle/_ref (Trajectory<7>) not present")
    dp = C.POINTER(C.c_double)
    ref.ref_traj7_create.restype = C.c_void_p; ref.ref_traj7_create.argtypes = [C.c_int, dp, dp]
    ref.ref_traj7_destroy.argtypes = [C.c_void_p]
    for fn in (ref.ref_traj7_max_vel_rate, ref.ref_traj7_max_acc_rate):
        fn.restype = C.c_double; fn.argtypes = [C.c_void_p]
    rng = np.random.default_rng(21)
    for N in (1, 4):
        head, tail, q, T = rand_problem(rng, 4, N)
        out = oracle.minco_forward(4, head, tail, q, T)
        flat = np.ascontiguousarray(out["flat"]); Tc = np.ascontiguousarray(T)
        want = np.zeros(2)
        for i in range(N):
            t = np.linspace(0.0, T[i], 40001)
            for d in (1, 2):
                v = np.stack([np.polyval(np.polyder(flat[i, a], d), t) for a in range(3)])
                want[d - 1] = max(want[d - 1], np.sqrt((v * v).sum(axis=0).max()))
        h = ref.ref_traj7_create(N, Tc.ctypes.data_as(dp), flat.ctypes.data_as(dp))
        try:
            assert abs(ref.ref_traj7_max_vel_rate(h) - want[0]) <= 1e-7 * want[0]
            assert abs(ref.ref_traj7_max_acc_rate(h) - want[1]) <= 1e-7 * want[1]
        finally:
            ref.ref_traj7_destroy(h)


def test_oracle_fixed_time_and_planner_row_flags(oracle):
    """mincob_params.flags in the oracle: FREEZE_TIMES zeroes the tau block of the gradient and leaves f and the
    waypoint block unchanged; PLANNER_ROWS on rows [n, b] (n.p <= b, learning_planner.hpp:293-299) equals the
    GCOPTER-sign rows [n, -b] bit for bit; lbfgs_optimize in fixed-time mode never moves tau."""
    from allocnet_b200 import params as P
    pb = synth.make_problems(24, N=5, K=16, S=3)
    x = pb.x0()
    f0, g0 = oracle.cost_batch(default_params(3), pb, x)
    f1, g1 = oracle.cost_batch(default_params(3, flags=P.FLAG_FREEZE_TIMES), pb, x)
    np.testing.assert_array_equal(f0, f1)
    assert (g1[:, :5] == 0).all() and (np.abs(g0[:, :5]).max(axis=1) > 0).all()
    np.testing.assert_array_equal(g0[:, 5:], g1[:, 5:])
    pl = pb.slice(0, pb.B); pl.hpolys = pb.hpolys.copy(); pl.hpolys[..., 3] *= -1.0
    f2, g2 = oracle.cost_batch(default_params(3, flags=P.FLAG_PLANNER_ROWS), pl, x)
    np.testing.assert_array_equal(f0, f2); np.testing.assert_array_equal(g0, g2)
    r = oracle.optimize_batch(default_params(3, flags=P.FLAG_FREEZE_TIMES), pb)
    np.testing.assert_array_equal(r["x"][:, :5], x[:, :5])
    assert (r["status"] >= 0).all() and (r["f"] < f0).all()
    rf = oracle.optimize_batch(default_params(3), pb)
    assert (rf["f"] <= r["f"] * (1 + 1e-6)).mean() >= 0.9      # freeing the durations can only help
