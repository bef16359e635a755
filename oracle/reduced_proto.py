"""numpy prototype of the REDUCED (junction-state) MINCO formulation used by the CUDA kernels.

TEST INFRASTRUCTURE ONLY (same rule as oracle/minco_oracle.hpp): it is the readable statement of
the math in allocnet_b200/csrc/minco_device.cuh, checked against the banded oracle by
tests/test_reduced_formulation.py.  Nothing under allocnet_b200/ imports it.

Why a second formulation.  The banded system of upstream minco.hpp (SURVEY.md Appendix A.2,
2S*N unknowns, half-bandwidth 2S) is a strictly sequential elimination: 48 pivot steps for
N=8.  The same spline is the minimiser of the energy over the junction derivatives
y_j = (p', .., p^(S-1)) at the N-1 inner waypoints with positions q_j and times T fixed, because
the rows of A are exactly "continuity up to derivative 2S-2" = stationarity of
E = sum_i int |p_i^(S)|^2 with respect to y_j.  Writing piece i through its boundary states
s_i = [start derivs 0..S-1 ; end derivs 0..S-1] (Hermite form, c_i = H(T_i) s_i) gives

    E = sum_i s_i^T W(T_i) s_i ,   W(T) = T^(1-2S) L(T) What L(T),  L = diag(1,T,..,T^(S-1)) twice,

and dE/dy_j = 0 is a symmetric positive definite BLOCK-TRIDIAGONAL system with (S-1)x(S-1) blocks
and N-1 block rows: 7 block rows of 2x2 for N=8, S=3.  One lane per piece builds its own blocks
from T_i and the system is solved by the lane-to-lane block elimination sweep the kernels use
(`sweep_solve`: N-2 rounds in which EVERY row recomputes itself from its own data and its neighbour's
current value, so the loop body has no row-dependent predicate); the first kernel versions used
parallel cyclic reduction (`pcr_solve`, kept as a cross-check).

Adjoint.  With G_i = dF/dc_i (energy + penalties) and the partial dF/dT_i, let
z_i = Hhat^T Gamma G_i (Gamma = diag(T^-k)), g_s = L z_i, gather g_y at junctions, solve
M mu = g_y with the SAME factorisation (M is symmetric), m_i = [0,mu_i ; 0,mu_{i+1}], then
    dJ/dq_j = g_p[j] - [W_{j-1} m_{j-1}]_(end p) - [W_j m_j]_(start p)
    dJ/dT_i = dF/dT_i - (1/T) sum_k k G_ik.c_ik + z_i.(L' s_i) - m_i^T W_i' s_i
which equals upstream propogateGrad to rounding.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np


def _fallfac(k, d):
    f = 1
    for u in range(d):
        f *= (k - u)
    return f


def hermite_constants(S, exact=False):
    """Hhat (2S x 2S): unit-interval Hermite -> monomial; Qhat; What = Hhat^T Qhat Hhat."""
    D = 2 * S
    if exact:
        import sympy as sp
        V = sp.zeros(D, D)
        for d in range(S):
            V[d, d] = math.factorial(d)
            for k in range(d, D):
                V[S + d, k] = _fallfac(k, d)
        H = V.inv()
        Q = sp.zeros(D, D)
        for a in range(S, D):
            for c in range(S, D):
                Q[a, c] = sp.Rational(_fallfac(a, S) * _fallfac(c, S), a + c - 2 * S + 1)
        W = H.T * Q * H
        return H, Q, W
    V = np.zeros((D, D))
    for d in range(S):
        V[d, d] = math.factorial(d)
        for k in range(d, D):
            V[S + d, k] = _fallfac(k, d)
    H = np.linalg.inv(V)
    Q = np.zeros((D, D))
    for a in range(S, D):
        for c in range(S, D):
            Q[a, c] = _fallfac(a, S) * _fallfac(c, S) / (a + c - 2 * S + 1)
    return H, Q, H.T @ Q @ H


def pcr_solve(Lb, Db, Ub, R):
    """Parallel cyclic reduction on a block-tridiagonal system, n block rows (any n >= 1).

    Lb,Db,Ub: (n,b,b) (Lb[0] and Ub[n-1] ignored/zero), R: (n,b,m).  Returns X (n,b,m) and the
    multiplier record needed to re-solve with another right-hand side (see pcr_resolve)."""
    n = Db.shape[0]
    Lb, Db, Ub, R = Lb.copy(), Db.copy(), Ub.copy(), R.copy()
    Lb[0] = 0.0
    Ub[n - 1] = 0.0
    rec = []
    s = 1
    while s < n:
        Ln, Dn, Un, Rn = Lb.copy(), Db.copy(), Ub.copy(), R.copy()
        al = np.zeros_like(Db)
        ga = np.zeros_like(Db)
        for j in range(n):
            if j - s >= 0:
                al[j] = Lb[j] @ np.linalg.inv(Db[j - s])
                Dn[j] -= al[j] @ Ub[j - s]
                Rn[j] -= al[j] @ R[j - s]
                Ln[j] = -al[j] @ Lb[j - s]
            else:
                Ln[j] = 0.0
            if j + s < n:
                ga[j] = Ub[j] @ np.linalg.inv(Db[j + s])
                Dn[j] -= ga[j] @ Lb[j + s]
                Rn[j] -= ga[j] @ R[j + s]
                Un[j] = -ga[j] @ Ub[j + s]
            else:
                Un[j] = 0.0
        rec.append((s, al, ga))
        Lb, Db, Ub, R = Ln, Dn, Un, Rn
        s *= 2
    Dinv = np.linalg.inv(Db)
    return np.einsum("nab,nbm->nam", Dinv, R), (rec, Dinv)


def pcr_resolve(fact, R):
    rec, Dinv = fact
    R = R.copy()
    n = R.shape[0]
    for s, al, ga in rec:
        Rn = R.copy()
        for j in range(n):
            if j - s >= 0:
                Rn[j] -= al[j] @ R[j - s]
            if j + s < n:
                Rn[j] -= ga[j] @ R[j + s]
        R = Rn
    return np.einsum("nab,nbm->nam", Dinv, R)


def sweep_solve(Lb, Db, Ub, R, rounds=None):
    """Block elimination as the device does it (allocnet_b200/csrc/minco_device.cuh: spline_solve).

    Forward:  M_j = L_j Dinv'_{j-1},  D'_j = D_j - M_j U_{j-1},  r'_j = r_j - M_j r'_{j-1};
    backward: y_j = Dinv'_j (r'_j - U_j y_{j+1}).
    Every round ALL rows are recomputed from their ORIGINAL D, r and the neighbour's current values (a
    Jacobi sweep): row j is final after round j and is reproduced unchanged afterwards, so n-1 rounds
    are exact for n rows and any larger warp-uniform count is harmless."""
    n, b = Db.shape[0], Db.shape[1]
    rounds = max(n - 1, 0) if rounds is None else rounds
    Lb = Lb.copy(); Ub = Ub.copy()
    Lb[0] = 0.0; Ub[n - 1] = 0.0
    Dinv = np.linalg.inv(Db)
    rr = R.copy()
    M = np.zeros_like(Db)
    for _ in range(rounds):
        Dp = np.concatenate([np.eye(b)[None], Dinv[:-1]])          # shuffle-up by one row
        Up = np.concatenate([np.zeros((1, b, b)), Ub[:-1]])
        rp = np.concatenate([np.zeros_like(R[:1]), rr[:-1]])
        M = np.einsum("nab,nbc->nac", Lb, Dp)
        Dn = Db - np.einsum("nab,nbc->nac", M, Up)
        rr = R - np.einsum("nab,nbm->nam", M, rp)
        Dinv = np.linalg.inv(Dn)
    y = np.einsum("nab,nbm->nam", Dinv, rr)
    for _ in range(rounds):
        yn = np.concatenate([y[1:], np.zeros_like(y[:1])])          # shuffle-down by one row
        y = np.einsum("nab,nbm->nam", Dinv, rr - np.einsum("nab,nbm->nam", Ub, yn))
    return y, (M, Dinv, Ub, rounds)


def sweep_resolve(fact, R):
    """Same factorisation, new right-hand side (allocnet_b200/csrc/minco_device.cuh: sweep_apply)."""
    M, Dinv, Ub, rounds = fact
    rr = R.copy()
    for _ in range(rounds):
        rp = np.concatenate([np.zeros_like(R[:1]), rr[:-1]])
        rr = R - np.einsum("nab,nbm->nam", M, rp)
    y = np.einsum("nab,nbm->nam", Dinv, rr)
    for _ in range(rounds):
        yn = np.concatenate([y[1:], np.zeros_like(y[:1])])
        y = np.einsum("nab,nbm->nam", Dinv, rr - np.einsum("nab,nbm->nam", Ub, yn))
    return y


class ReducedMinco:
    """Same call sequence as orc::Minco<S> (setConditions/setParameters/.../propogateGrad)."""

    def __init__(self, S):
        self.S = S
        self.D = 2 * S
        H, Q, W = hermite_constants(S, exact=True)      # exact rationals -> nearest doubles
        self.H = np.array(H.tolist(), dtype=float)
        self.Q = np.array(Q.tolist(), dtype=float)
        self.W = np.array(W.tolist(), dtype=float)

    def set_conditions(self, head, tail, N):
        self.head = np.asarray(head, float).reshape(self.S, 3)
        self.tail = np.asarray(tail, float).reshape(self.S, 3)
        self.N = N

    def _lam(self, T):
        return np.array([T ** d for d in range(self.S)] * 2)

    def _dlam(self, T):
        return np.array([d * T ** (d - 1) if d > 0 else 0.0 for d in range(self.S)] * 2)

    def _Wi(self, T):
        lam = self._lam(T)
        return T ** (1 - 2 * self.S) * (lam[:, None] * self.W * lam[None, :])

    def set_parameters(self, q, T):
        S, N, D = self.S, self.N, self.D
        q = np.asarray(q, float).reshape(max(N - 1, 0), 3)
        T = np.asarray(T, float)
        self.T = T
        P = np.vstack([self.head[0:1], q, self.tail[0:1]])          # N+1 positions
        b = S - 1
        ia = list(range(1, S))            # start unknown rows of s
        ib = list(range(S + 1, 2 * S))    # end unknown rows of s
        Wl = [self._Wi(T[i]) for i in range(N)]
        nj = N - 1
        Y = np.zeros((N + 1, b, 3))
        Y[0] = self.head[1:]
        Y[N] = self.tail[1:]
        self.fact = None
        if nj > 0:
            Db = np.zeros((nj, b, b)); Ub = np.zeros((nj, b, b)); Lb = np.zeros((nj, b, b)); R = np.zeros((nj, b, 3))
            for j in range(1, N):          # junction j between piece j-1 and piece j
                Wp, Wn = Wl[j - 1], Wl[j]
                Db[j - 1] = Wp[np.ix_(ib, ib)] + Wn[np.ix_(ia, ia)]
                Ub[j - 1] = Wn[np.ix_(ia, ib)]
                Lb[j - 1] = Wp[np.ix_(ib, ia)]
                # W[:,0] == -W[:,S] exactly (the energy sees positions only through p1-p0):
                # use waypoint differences, never absolute positions, to avoid cancellation.
                r = -(np.outer(Wp[ib, S], P[j] - P[j - 1]) + np.outer(Wn[ia, S], P[j + 1] - P[j]))
                if j == 1:
                    r -= Wp[np.ix_(ib, ia)] @ Y[0]
                if j == N - 1:
                    r -= Wn[np.ix_(ia, ib)] @ Y[N]
                R[j - 1] = r
            X, self.fact = sweep_solve(Lb, Db, Ub, R)
            Y[1:N] = X
        self.P, self.Y, self.Wl = P, Y, Wl
        # boundary states and coefficients
        self.s = np.zeros((N, D, 3))
        self.c = np.zeros((N, D, 3))
        for i in range(N):
            self.s[i, 0] = P[i]; self.s[i, 1:S] = Y[i]
            self.s[i, S] = P[i + 1]; self.s[i, S + 1:] = Y[i + 1]
            lam = self._lam(T[i])
            sh = lam[:, None] * self.s[i]
            sh0 = sh.copy(); sh0[0] = 0.0; sh0[S] = P[i + 1] - P[i]   # Hhat[k,0] == -Hhat[k,S] for k >= S
            chat = self.H @ sh0
            chat[0] = P[i]
            self.c[i] = chat / (T[i] ** np.arange(D))[:, None]
        return self.c.reshape(N * D, 3)

    def propagate_grad(self, gdC, gdT):
        S, N, D = self.S, self.N, self.D
        G = np.asarray(gdC, float).reshape(N, D, 3)
        gdT = np.asarray(gdT, float)
        T = self.T
        b = S - 1
        z = np.zeros((N, D, 3)); gs = np.zeros((N, D, 3))
        for i in range(N):
            gam = 1.0 / T[i] ** np.arange(D)
            z[i] = self.H.T @ (gam[:, None] * G[i])
            gs[i] = self._lam(T[i])[:, None] * z[i]
        gy = np.zeros((max(N - 1, 0), b, 3)); gp = np.zeros((max(N - 1, 0), 3))
        for j in range(1, N):
            gy[j - 1] = gs[j - 1, S + 1:] + gs[j, 1:S]
            gp[j - 1] = gs[j - 1, S] + gs[j, 0]
        MU = np.zeros((N + 1, b, 3))
        if N > 1:
            MU[1:N] = sweep_resolve(self.fact, gy)
        gq = gp.copy(); gT = np.zeros(N)
        for i in range(N):
            m = np.zeros((D, 3))
            m[1:S] = MU[i]; m[S + 1:] = MU[i + 1]
            lam, dlam = self._lam(T[i]), self._dlam(T[i])
            t5 = T[i] ** (1 - 2 * S)
            wm = self.W @ (lam[:, None] * m)
            ws = self.W @ (lam[:, None] * self.s[i])
            if i >= 1:
                gq[i - 1] -= t5 * wm[0]
            if i <= N - 2:
                gq[i] -= t5 * wm[S]
            k = np.arange(D)[:, None]
            through_H = -(k * G[i] * self.c[i]).sum() / T[i] + (z[i] * (dlam[:, None] * self.s[i])).sum()
            mWs = t5 * (-(2 * S - 1) / T[i] * ((lam[:, None] * m) * ws).sum()
                        + ((dlam[:, None] * m) * ws).sum() + ((dlam[:, None] * self.s[i]) * wm).sum())
            gT[i] = gdT[i] + through_H - mWs
        return gq, gT
