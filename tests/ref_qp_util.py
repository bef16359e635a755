"""Helpers shared by the CPU and GPU tests that use tests/golden/ref_qp_*.npz -- the Q, A, b, G, h (and
autograd Jacobians dQ/dT, dA/dT) built by the reference's own network/utils/min_traj_opt.py, see
tests/golden/make_ref_qp_fixtures.py."""
import os

import numpy as np

REF_QP_CASES = ["n2", "n3", "n5", "n8"]


def load_ref_qp(name):
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_qp_{name}.npz"))
    return {k: f[k] for k in f.files}


def ref_qp_problem(c, seed):
    """Inputs of the fixture in this repository's layouts + seeded inner waypoints."""
    st = c["state"]                                  # {9,2}: [px,vx,ax,py,...] x {start, goal}
    head = st[:, 0].reshape(3, 3).T.copy()           # rows P,V,A ; cols axis
    tail = st[:, 1].reshape(3, 3).T.copy()
    T = c["times"].copy(); N = len(T)
    rng = np.random.default_rng(seed)
    w = np.linspace(0.0, 1.0, N + 1)[1:-1, None]
    q = head[0] * (1 - w) + tail[0] * w + rng.normal(scale=0.4, size=(N - 1, 3))
    return head, tail, q, T


def solve_ld(M, r):
    """Gaussian elimination with partial pivoting in 80-bit long double (so that the KKT solve itself does
    not limit the comparison at 1e-9)."""
    M = M.astype(np.longdouble).copy(); r = r.astype(np.longdouble).copy(); n = len(r)
    for k in range(n):
        p = k + int(np.argmax(np.abs(M[k:, k])))
        if p != k:
            M[[k, p]] = M[[p, k]]; r[[k, p]] = r[[p, k]]
        f = M[k + 1:, k] / M[k, k]
        M[k + 1:, k:] -= f[:, None] * M[k, k:][None, :]
        r[k + 1:] -= f * r[k]
    x = np.zeros(n, np.longdouble)
    for k in range(n - 1, -1, -1):
        x[k] = (r[k] - M[k, k + 1:] @ x[k + 1:]) / M[k, k]
    return x.astype(np.float64)




def ref_qp_kkt(c, q):
    """Solve min 1/2 z^T Q z s.t. A z = b and the waypoint rows (start position of piece i+1 = q_i, selected with
    the reference's own zero_A row 0).  -> z, multipliers of A rows, multipliers of the waypoint rows."""
    Q, A, b = c["Q"], c["A"], c["b"]
    N = len(c["times"]); d = 6; nv = N * 3 * d
    W = np.zeros((3 * (N - 1), nv)); wq = np.zeros(3 * (N - 1))
    for i in range(N - 1):
        for a in range(3):
            W[3 * i + a, (i + 1) * 3 * d + a * d:(i + 1) * 3 * d + a * d + d] = c["zero_A"][0]
            wq[3 * i + a] = q[i, a]
    Cm = np.vstack([A, W]); rhs = np.concatenate([b, wq]); m = Cm.shape[0]
    KKT = np.block([[Q, Cm.T], [Cm, np.zeros((m, m))]])
    sol = solve_ld(KKT, np.concatenate([np.zeros(nv), rhs]))
    return sol[:nv], sol[nv:nv + A.shape[0]], sol[nv + A.shape[0]:]
