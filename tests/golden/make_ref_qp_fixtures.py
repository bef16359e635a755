"""Golden matrices produced BY THE REFERENCE'S OWN CODE for the incumbent QP back-end.

Runs only where the read-only reference tree is present (this container, not the GPU box):

    python tests/golden/make_ref_qp_fixtures.py        # writes tests/golden/ref_qp_*.npz

It imports `/root/reference/network/utils/min_traj_opt.py` UNMODIFIED (the Python twin of
`src/planner/include/planner/qp_solver.hpp:119-360`).  Three of the module's
top-level imports are absent from this image and are never touched by the functions called here
(`cvxpy`, `osqp`, `memory_profiler`): they are satisfied with empty stand-in modules in `sys.modules`.
Then, for seeded inputs, `MinTrajOpt.update(state, hpolys, times, phase=2)` is executed, which calls the
reference's `fill_eq_obj` (:377-533 -> Q, A, b) and `fill_ineq` (:535-607 -> G1, h1, G2, h2).  Those
matrices are stored next to the inputs.  Nothing of this repository takes part in producing them.

`tests/test_oracle_minco.py::test_reference_qp_matrices_*` then demand that
  * the minimiser of 1/2 z^T Q z subject to A z = b plus waypoint rows (rows that select a piece's start
    position with the reference's own `zero_A[0]`) IS the oracle's MINCO_S3NU coefficient set (<= 1e-9), and
    1/2 z^T Q z == E / 2;
  * 2 Q z, z^T dQ/dT z and the KKT sensitivities (lambda^T dA/dT z, waypoint multipliers) equal the oracle's
    getEnergyPartialGradByCoeffs / ByTimes and propogateGrad (dQ/dT, dA/dT are torch.autograd Jacobians taken
    through the reference's fill_eq_obj);
  * G1 z - h1 and G2 z - h2 evaluated on oracle / device coefficients equal the corridor and box residuals
    sampled at the reference's `res` left-end points (layout pin for `idx = i*3*d + j*d + k`).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_NET = "/root/reference/network"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference_twin():
    if not os.path.exists(os.path.join(REF_NET, "utils", "min_traj_opt.py")):
        raise SystemExit("reference tree absent: fixtures can only be generated where /root/reference is")
    for name in ("cvxpy", "osqp"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "memory_profiler" not in sys.modules:
        try:
            __import__("memory_profiler")
        except Exception:
            m = types.ModuleType("memory_profiler")
            m.profile = lambda f: f
            sys.modules["memory_profiler"] = m
    if REF_NET not in sys.path:
        sys.path.insert(0, REF_NET)
    from utils import min_traj_opt  # noqa: the reference module, unmodified
    return min_traj_opt


def reference_params(order: int, res: int):
    # same keys as /root/reference/network/utils/params.yaml; limits of src/planner/config/planner.yaml:17,19
    return {
        "physical_limits": {"max_vel": 4.0, "max_acc": 6.0, "max_jerk": 12.0},
        "phase1_physical_limits": {"max_vel": 5.0, "max_acc": 8.0, "max_jerk": 10.0, "inf_dis": 0.1},
        "planning": {"order": order, "state_dim": 3, "dim": 3, "res": res, "seg": 5, "var_num": 120,
                     "use_time_factor": False},
    }


def make_case(mto, seed: int, N: int, K: int, res: int):
    import torch
    rng = np.random.default_rng(seed)
    # state tensor {9,2}: [px,vx,ax,py,vy,ay,pz,vz,az] x {start, goal} (learning_planner.hpp:147-155)
    start = np.zeros(9); goal = np.zeros(9)
    p0 = rng.uniform(-3, 3, 3); p1 = p0 + rng.uniform(2, 5, 3) * rng.choice([-1.0, 1.0], 3)
    start[0::3] = p0; start[1::3] = rng.uniform(-1, 1, 3); start[2::3] = rng.uniform(-0.5, 0.5, 3)
    goal[0::3] = p1; goal[1::3] = rng.uniform(-0.3, 0.3, 3); goal[2::3] = rng.uniform(-0.2, 0.2, 3)
    state = np.stack([start, goal], axis=1)
    times = rng.uniform(0.6, 2.0, N)
    # polytopes in the planner's [n, b] form (n.p <= b, unit normals; learning_planner.hpp:293-299), K rows each
    hp = np.zeros((K, 4, N))
    for i in range(N):
        n = rng.normal(size=(K, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True)
        c = p0 + (p1 - p0) * (i + 0.5) / N
        hp[:, :3, i] = n
        hp[:, 3, i] = n @ c + rng.uniform(1.0, 3.0, K)
    opt = mto.MinTrajOpt(reference_params(3, res))
    opt.update(torch.tensor(state), torch.tensor(hp), torch.tensor(times), phase=2, seq_len=N)
    Q, A, b, G1, h1, G2, h2 = [np.asarray(t.detach().numpy(), dtype=np.float64) for t in opt.params]
    assert opt.seg == N
    # d/dT of the reference's Q(T) and A(T) by torch.autograd THROUGH THE REFERENCE'S fill_eq_obj (it builds both
    # from `times` with differentiable torch ops; layers.py relies on exactly this)
    tt = torch.tensor(times, requires_grad=True)
    dQ, dA = torch.autograd.functional.jacobian(lambda t: tuple(opt.fill_eq_obj(t)[:2]), tt)
    dQ = np.moveaxis(np.asarray(dQ.detach().numpy(), dtype=np.float64), -1, 0)   # [N][nv][nv]
    dA = np.moveaxis(np.asarray(dA.detach().numpy(), dtype=np.float64), -1, 0)   # [N][eq][nv]
    return dict(dQ=dQ, dA=dA, state=state, times=times, hpolys=hp, Q=Q, A=A, b=b, G1=G1, h1=h1, G2=G2, h2=h2,
                zero_A=np.asarray(opt.zero_A.numpy(), dtype=np.float64), res=np.int64(res), order=np.int64(3))


CASES = [  # (name, seed, pieces, rows per polytope, res)
    ("n2", 11, 2, 4, 5),
    ("n3", 12, 3, 6, 5),
    ("n5", 13, 5, 8, 20),   # the planner's own size: ModelMaxSeg 5, ConstRes 20 (learning_planner.hpp:33, planner.yaml:21)
    ("n8", 14, 8, 4, 4),
]


def main():
    mto = import_reference_twin()
    for name, seed, N, K, res in CASES:
        case = make_case(mto, seed, N, K, res)
        out = os.path.join(HERE, f"ref_qp_{name}.npz")
        np.savez_compressed(out, **case)
        print(out, {k: getattr(v, "shape", v) for k, v in case.items()})


if __name__ == "__main__":
    main()
