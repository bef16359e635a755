"""Seeded synthetic corridor problems for the batched MINCO optimizer.

The reference ships no dataset and no benchmark inputs; its
corridors come from OMPL + FIRI at run time (`gcopter/sfc_gen.hpp:116-186`).
This generator follows SURVEY.md §8(d): a random waypoint chain inside the map
box of `launch/learning_planning.launch:9-14`, one polytope per piece
(`planner/qp_solver.hpp:126,255-259`: `seg = hPolys.size()`), K half-planes per
polytope in GCOPTER sign `n.p + d <= 0` (`gcopter/geo_utils.hpp:41-42`), initial
durations from the trapezoid rule of `network/utils/min_traj_opt.py:195-206`
with `MaxVelBox/MaxAccBox` of `config/planner.yaml:17,19`.

Every draw is a pure function of (seed, problem index, draw index) through a
splitmix64 hash, so problem p is identical no matter how the batch is sharded
across ranks and can be regenerated anywhere (CPU oracle, GPU box, C).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED = 0xA110C000
BOX_LO = np.array([-10.0, -10.0, 0.0])
BOX_HI = np.array([10.0, 10.0, 5.0])
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


class _Draws:
    """u(p, k) in [0,1): counter-based uniform doubles."""

    def __init__(self, seed: int, pidx: np.ndarray):
        with np.errstate(over="ignore"):
            self.base = _splitmix64(np.uint64(seed) ^ pidx.astype(np.uint64))

    def u(self, k: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            h = _splitmix64(self.base + np.uint64(k) * np.uint64(0xD1342543DE82EF95))
        return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def uniform(self, k: int, lo: float, hi: float) -> np.ndarray:
        return lo + (hi - lo) * self.u(k)

    def sphere(self, k: int) -> np.ndarray:
        z = 2.0 * self.u(k) - 1.0
        phi = 2.0 * np.pi * self.u(k + 1)
        r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
        return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1)


@dataclass
class ProblemBatch:
    """Host-side (numpy, fp64, C-contiguous) batch in the C-ABI layouts of include/mincob.h."""

    S: int
    N: int
    K: int
    head: np.ndarray    # [B][S][3]   rows P,V,A[,J]  (== Eigen col-major 3xS iniPVA)
    tail: np.ndarray    # [B][S][3]
    hpolys: np.ndarray  # [B][N][K][4] (nx,ny,nz,d), n.p + d <= 0
    hrows: np.ndarray   # [B][N] int32, rows used per polytope (<= K)
    q0: np.ndarray      # [B][N-1][3] initial inner waypoints
    T0: np.ndarray      # [B][N]      initial durations

    @property
    def B(self) -> int:
        return self.head.shape[0]

    @property
    def nvars(self) -> int:
        return self.N + 3 * (self.N - 1)

    def x0(self) -> np.ndarray:
        """Decision vector x = [tau(N); q(3(N-1))] with tau = backwardT(T0)."""
        return np.concatenate([backward_t(self.T0), self.q0.reshape(self.B, -1)], axis=1).copy()

    def slice(self, lo: int, hi: int) -> "ProblemBatch":
        c = np.ascontiguousarray
        return ProblemBatch(self.S, self.N, self.K, c(self.head[lo:hi]), c(self.tail[lo:hi]),
                            c(self.hpolys[lo:hi]), c(self.hrows[lo:hi]), c(self.q0[lo:hi]), c(self.T0[lo:hi]))


def forward_t(tau: np.ndarray) -> np.ndarray:
    tau = np.asarray(tau, dtype=np.float64)
    pos = (0.5 * tau + 1.0) * tau + 1.0
    neg = 1.0 / ((0.5 * tau - 1.0) * tau + 1.0)
    return np.where(tau > 0.0, pos, neg)


def backward_t(T: np.ndarray) -> np.ndarray:
    T = np.asarray(T, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        big = np.sqrt(np.maximum(2.0 * T - 1.0, 0.0)) - 1.0
        small = 1.0 - np.sqrt(np.maximum(2.0 / T - 1.0, 0.0))
    return np.where(T > 1.0, big, small)


def make_problems(B: int, N: int = 8, K: int = 16, S: int = 3, *, seed: int = SEED, first: int = 0,
                  v_max: float = 4.0, a_max: float = 6.0, time_scale: float = 1.5,
                  rest_to_rest: bool = False, ragged_rows: bool = False) -> ProblemBatch:
    """Problems first .. first+B-1 of the seeded stream.

    K = 0 gives the energy-only configuration (no corridor rows).  `ragged_rows`
    keeps a per-polytope prefix of hrows in [min(6,K), K] (rows beyond are zero,
    as the reference zero-pads polytopes: planner/learning_planner.hpp:157-168).
    """
    if N < 1 or K < 0 or S not in (3, 4):
        raise ValueError("need N>=1, K>=0, S in {3,4}")
    pidx = np.arange(first, first + B, dtype=np.uint64)
    dr = _Draws(seed, pidx)
    lo, hi = BOX_LO, BOX_HI

    w = np.empty((B, N + 1, 3))
    for a in range(3):
        w[:, 0, a] = dr.uniform(a, lo[a] + 2.0 if hi[a] - lo[a] > 4.0 else lo[a] + 0.25 * (hi[a] - lo[a]),
                                hi[a] - 2.0 if hi[a] - lo[a] > 4.0 else hi[a] - 0.25 * (hi[a] - lo[a]))
    v0 = np.stack([dr.uniform(3 + a, -1.0, 1.0) for a in range(3)], axis=-1)
    if rest_to_rest:
        v0[:] = 0.0
    direction = dr.sphere(6)
    seglen = np.empty((B, N))
    hpolys = np.zeros((B, N, max(K, 1), 4))
    hrows = np.zeros((B, N), dtype=np.int32)

    for i in range(N):
        # draw indices of piece i: [k0, k0 + stride).  Up to 40 rows the plane draws (k0+16 .. k0+16+3(K-6)) and the ragged-row
        # draw stay inside 128 indices (and the committed fixtures depend on that numbering); above, a wider stride keeps
        # neighbouring pieces' draws disjoint
        stride = 128 if K <= 40 else 512
        k0 = 64 + stride * i
        ragged_at = k0 + (120 if K <= 40 else 500)
        if i > 0:  # rotate the previous direction by at most 60 degrees
            r = dr.sphere(k0 + 1)
            theta = dr.uniform(k0 + 3, 0.0, np.pi / 3.0)
            perp = r - np.sum(r * direction, axis=1, keepdims=True) * direction
            pn = np.linalg.norm(perp, axis=1, keepdims=True)
            perp = np.where(pn > 1e-9, perp / np.maximum(pn, 1e-300), 0.0)
            direction = np.cos(theta)[:, None] * direction + np.sin(theta)[:, None] * perp
            direction /= np.linalg.norm(direction, axis=1, keepdims=True)
        ell = dr.uniform(k0, 1.5, 3.5)
        nxt = w[:, i] + ell[:, None] * direction
        for a in range(3):  # reflect at the box faces
            over = nxt[:, a] > hi[a]
            under = nxt[:, a] < lo[a]
            nxt[:, a] = np.where(over, 2.0 * hi[a] - nxt[:, a], nxt[:, a])
            nxt[:, a] = np.where(under, 2.0 * lo[a] - nxt[:, a], nxt[:, a])
            direction[:, a] = np.where(over | under, -direction[:, a], direction[:, a])
        w[:, i + 1] = nxt
        seglen[:, i] = np.linalg.norm(nxt - w[:, i], axis=1)

        if K > 0:
            a_pt, b_pt = w[:, i], w[:, i + 1]
            mn, mx = np.minimum(a_pt, b_pt), np.maximum(a_pt, b_pt)
            rows = []
            for f in range(min(6, K)):  # AABB faces, each with its own inflation
                infl = dr.uniform(k0 + 5 + f, 0.6, 1.5)
                ax, sgn = f // 2, (1.0 if f % 2 == 0 else -1.0)
                n = np.zeros((B, 3))
                n[:, ax] = sgn
                d = -(mx[:, ax] + infl) if sgn > 0 else (mn[:, ax] - infl)
                rows.append(np.concatenate([n, d[:, None]], axis=1))
            for r_ in range(max(0, K - 6)):  # random supporting planes with margin
                n = dr.sphere(k0 + 16 + 3 * r_)
                margin = dr.uniform(k0 + 16 + 3 * r_ + 2, 0.3, 1.0)
                d = -np.maximum(np.sum(n * a_pt, axis=1), np.sum(n * b_pt, axis=1)) - margin
                rows.append(np.concatenate([n, d[:, None]], axis=1))
            hpolys[:, i, :K] = np.stack(rows, axis=1)
            hrows[:, i] = K
            if ragged_rows:
                keep = np.minimum(K, min(6, K) + np.floor(dr.u(ragged_at) * (K - min(6, K) + 1)).astype(np.int32))
                hrows[:, i] = keep
                mask = np.arange(K)[None, :] >= keep[:, None]
                hpolys[:, i, :K][mask] = 0.0

    T0 = np.maximum(seglen / v_max, np.sqrt(2.0 * seglen / a_max)) * time_scale
    head = np.zeros((B, S, 3))
    tail = np.zeros((B, S, 3))
    head[:, 0] = w[:, 0]
    head[:, 1] = v0
    tail[:, 0] = w[:, N]
    if K == 0:
        hpolys = np.zeros((B, N, 0, 4))
    c = np.ascontiguousarray
    return ProblemBatch(S, N, K, c(head), c(tail), c(hpolys), c(hrows), c(w[:, 1:N]), c(T0))
