"""oracle/compact_lbfgs_proto.py: the Gram-matrix (compact) form of the L-BFGS direction equals the two-loop recursion
of lbfgs.hpp:672-709 -- groundwork for a smaller device kernel, checked on CPU only."""
import numpy as np
import pytest

from oracle.compact_lbfgs_proto import compact_direction, two_loop_direction


@pytest.mark.parametrize("n,b,seed", [(29, 8, 0), (29, 3, 1), (61, 8, 2), (9, 1, 3)])
def test_compact_form_equals_two_loop_on_random_histories(n, b, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n)); A = A @ A.T + n * np.eye(n)        # an SPD "Hessian": y = A s keeps every y.s > 0
    S = rng.normal(size=(b, n)); Y = S @ A
    g = rng.normal(size=n)
    d1, d2 = two_loop_direction(S, Y, g), compact_direction(S, Y, g)
    assert np.abs(d1 - d2).max() <= 1e-10 * np.abs(d1).max()
    assert d1.dot(g) < 0.0                                            # a descent direction


def test_compact_form_on_a_real_history(oracle):
    """History taken from an actual run of the restated optimizer on a corridor problem (the pairs the device kernel
    would hold): consecutive iterates x_k, gradients g_k from the oracle's cost functional."""
    from allocnet_b200 import synth
    from allocnet_b200.params import default_params
    pb = synth.make_problems(1, N=8, K=16, S=3)
    xs, gs = [], []
    for it in range(20, 30):
        prm = default_params(3, max_iterations=it)
        r = oracle.optimize_batch(prm, pb)
        _, g = oracle.cost_batch(default_params(3), pb, r["x"])
        xs.append(r["x"][0]); gs.append(g[0])
    S = np.diff(np.array(xs), axis=0)[-8:]; Y = np.diff(np.array(gs), axis=0)[-8:]
    keep = (S * Y).sum(axis=1) > 0                                    # the cautious update only keeps such pairs
    S, Y = S[keep], Y[keep]
    assert S.shape[0] >= 3
    d1, d2 = two_loop_direction(S, Y, gs[-1]), compact_direction(S, Y, gs[-1])
    assert np.abs(d1 - d2).max() <= 1e-7 * np.abs(d1).max()
