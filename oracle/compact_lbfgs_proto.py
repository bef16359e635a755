"""Prototype for the next round (test infrastructure / design note, not used by the product).

The device kernel is bound by instruction supply (profiles/r01_v13_summary.md) and the largest single block of code
on its per-iteration path is the unrolled two-loop recursion of lbfgs.hpp:672-709 (16 serial group reductions).  The
same direction d = -H g can be computed from the Gram matrices of the history,

    S = [s_0 .. s_{b-1}],  Y = [y_0 .. y_{b-1}]  (oldest first),   R = triu(S^T Y),   D = diag(s_i . y_i),
    gamma = (s_{b-1} . y_{b-1}) / (y_{b-1} . y_{b-1}),
    H g = gamma g + [S  gamma Y] [[R^-T (D + gamma Y^T Y) R^-1,  -R^-T], [-R^-1, 0]] [S^T g; gamma Y^T g]

(Byrd, Nocedal, Schnabel 1994, eq. 3.1 -- the published compact representation of the BFGS matrix), which needs
2b INDEPENDENT reductions per iteration for S^T g, Y^T g plus 2b + 1 for the new row / column of S^T Y and Y^T Y
(they pipeline, unlike the recursion's serial chain), b x b scalar work that every lane can repeat, and one rolled
pass over the history for the final combination.  This module states both forms in numpy so that a device version
has an oracle: tests/test_compact_lbfgs_proto.py checks they agree to rounding on random and on real histories.
"""
from __future__ import annotations

import numpy as np


def two_loop_direction(S: np.ndarray, Y: np.ndarray, g: np.ndarray) -> np.ndarray:
    """lbfgs.hpp:672-709 as written there: S, Y are [b][n] with the OLDEST pair first; returns d = -H g."""
    b = S.shape[0]
    d = -g.copy()
    alpha = np.zeros(b)
    for i in range(b - 1, -1, -1):                      # newest first
        alpha[i] = S[i].dot(d) / Y[i].dot(S[i])
        d -= alpha[i] * Y[i]
    d *= Y[b - 1].dot(S[b - 1]) / Y[b - 1].dot(Y[b - 1])
    for i in range(b):                                  # oldest first
        beta = Y[i].dot(d) / Y[i].dot(S[i])
        d += (alpha[i] - beta) * S[i]
    return d


def compact_direction(S: np.ndarray, Y: np.ndarray, g: np.ndarray) -> np.ndarray:
    """The same direction from Gram matrices only (no vector operation between the reductions and the final sum)."""
    b = S.shape[0]
    SY = S @ Y.T                                        # (S^T Y)_{ij} = s_i . y_j
    YY = Y @ Y.T
    R = np.triu(SY)
    D = np.diag(np.diag(SY))
    gamma = SY[b - 1, b - 1] / YY[b - 1, b - 1]
    p = S @ g                                           # S^T g
    q = gamma * (Y @ g)                                 # gamma Y^T g
    Rinv_p = np.linalg.solve(R, p)                      # b x b triangular solves: scalar work
    top = np.linalg.solve(R.T, (D + gamma * YY) @ Rinv_p - q)
    bot = -Rinv_p
    Hg = gamma * g + S.T @ top + gamma * (Y.T @ bot)
    return -Hg
