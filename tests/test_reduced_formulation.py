"""The junction-state (block-tridiagonal + lane-to-lane elimination sweep) formulation used by the CUDA kernels,
stated in numpy (oracle/reduced_proto.py), must reproduce the banded MINCO oracle: coefficients
(setParameters) and propogateGrad, for S = 3 and 4 and every piece count the kernels accept."""
import numpy as np
import pytest

from oracle.reduced_proto import ReducedMinco, hermite_constants, pcr_solve, pcr_resolve, sweep_solve, sweep_resolve


def _rand(rng, S, N):
    head = rng.normal(size=(S, 3)); tail = rng.normal(size=(S, 3))
    q = np.cumsum(rng.normal(size=(max(N - 1, 0), 3)), axis=0)
    T = rng.uniform(0.4, 2.5, size=N)
    return head, tail, q, T


@pytest.mark.parametrize("S,tol", [(3, 2e-11), (4, 1e-10)])
@pytest.mark.parametrize("N", [1, 2, 3, 5, 8, 16, 32])
def test_reduced_equals_banded(oracle, S, N, tol):
    rng = np.random.default_rng(100 * S + N)
    head, tail, q, T = _rand(rng, S, N)
    ref = oracle.minco_forward(S, head, tail, q, T)
    rm = ReducedMinco(S); rm.set_conditions(head, tail, N)
    c = rm.set_parameters(q, T)
    assert np.abs(c - ref["coeffs"]).max() <= tol * np.abs(ref["coeffs"]).max()
    gdC = rng.normal(size=(2 * S * N, 3)); gdT = rng.normal(size=N)
    gq_ref, gT_ref = oracle.minco_propagate(S, head, tail, q, T, gdC, gdT)
    gq, gT = rm.propagate_grad(gdC, gdT)
    if N > 1:
        assert np.abs(gq - gq_ref).max() <= tol * np.abs(gq_ref).max()
    assert np.abs(gT - gT_ref).max() <= tol * np.abs(gT_ref).max()


def test_hermite_constants_exact():
    """What must be symmetric, W[:,0] == -W[:,S], and H must interpolate the boundary derivatives."""
    for S in (3, 4):
        H, Q, W = hermite_constants(S, exact=True)
        D = 2 * S
        assert W == W.T
        assert all(W[i, 0] == -W[i, S] for i in range(D))
        assert all(H[k, 0] == -H[k, S] for k in range(S, D))
        # polynomial with coefficients H[:, j] has unit j-th boundary derivative and zero others
        import math
        import sympy as sp
        u = sp.symbols("u")
        for j in range(D):
            p = sum(H[k, j] * u ** k for k in range(D))
            for d in range(S):
                assert sp.diff(p, u, d).subs(u, 0) == (1 if j == d else 0)
                assert sp.diff(p, u, d).subs(u, 1) == (1 if j == S + d else 0)


def test_pcr_matches_dense():
    rng = np.random.default_rng(7)
    for n, b in [(1, 2), (2, 2), (7, 2), (8, 3), (15, 3), (31, 2)]:
        # SPD block tridiagonal
        M = np.zeros((n * b, n * b))
        blocks = [rng.normal(size=(b, b)) for _ in range(n - 1)]
        for j in range(n):
            A = rng.normal(size=(b, b)); M[j * b:(j + 1) * b, j * b:(j + 1) * b] = A @ A.T + 4 * b * np.eye(b)
        for j, Bk in enumerate(blocks):
            M[j * b:(j + 1) * b, (j + 1) * b:(j + 2) * b] = Bk
            M[(j + 1) * b:(j + 2) * b, j * b:(j + 1) * b] = Bk.T
        Db = np.stack([M[j * b:(j + 1) * b, j * b:(j + 1) * b] for j in range(n)])
        Ub = np.stack([M[j * b:(j + 1) * b, (j + 1) * b:(j + 2) * b] if j < n - 1 else np.zeros((b, b)) for j in range(n)])
        Lb = np.stack([M[j * b:(j + 1) * b, (j - 1) * b:j * b] if j > 0 else np.zeros((b, b)) for j in range(n)])
        R = rng.normal(size=(n, b, 3))
        X, fact = pcr_solve(Lb, Db, Ub, R)
        np.testing.assert_allclose(X.reshape(n * b, 3), np.linalg.solve(M, R.reshape(n * b, 3)), rtol=0, atol=1e-12)
        R2 = rng.normal(size=(n, b, 3))
        np.testing.assert_allclose(pcr_resolve(fact, R2).reshape(n * b, 3), np.linalg.solve(M, R2.reshape(n * b, 3)),
                                   rtol=0, atol=1e-12)
        # the sweep the kernels use now; extra (warp-uniform) rounds must not change the answer
        for extra in (0, 3):
            X, fact = sweep_solve(Lb, Db, Ub, R, rounds=max(n - 1, 0) + extra)
            np.testing.assert_allclose(X.reshape(n * b, 3), np.linalg.solve(M, R.reshape(n * b, 3)), rtol=0, atol=1e-12)
            np.testing.assert_allclose(sweep_resolve(fact, R2).reshape(n * b, 3), np.linalg.solve(M, R2.reshape(n * b, 3)),
                                       rtol=0, atol=1e-12)
