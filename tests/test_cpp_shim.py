"""The C++ host side (include/mincob/*.hpp: minco::MINCO_S3NU, mincob::PolytopeSFC, mincob::solve)
driven the way the planner would drive it, against the CPU oracle.  The binaries are built by
__graft_entry__.build(); `test_shim_reflbfgs` additionally compiles the REFERENCE's own
gcopter/lbfgs.hpp (verbatim) and lets it drive the GPU costFunctional through the
lbfgs_evaluate_t callback ABI (lbfgs.hpp:200-202) -- the wiring BASELINE.json names."""
import json
import os
import subprocess

import numpy as np
import pytest

from allocnet_b200 import synth
from allocnet_b200.params import default_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_shim")
BIN_REF = os.path.join(ROOT, "oracle", "_ref", "test_shim_reflbfgs")


def _write_problem(path, pb, b):
    """One problem in the planner's own conventions: PVA columns, durations as float32, rows [n, b] with n.p <= b."""
    N, K = pb.N, pb.K
    with open(path, "w") as fh:
        fh.write(f"{N} {K}\n")
        fh.write(" ".join(repr(float(v)) for v in pb.head[b].ravel()) + "\n")
        fh.write(" ".join(repr(float(v)) for v in pb.tail[b].ravel()) + "\n")
        fh.write(" ".join(repr(float(np.float32(v))) for v in pb.T0[b]) + "\n")
        fh.write(" ".join(repr(float(v)) for v in pb.q0[b].ravel()) + "\n")
        for i in range(N):
            rows = int(pb.hrows[b, i])
            fh.write(f"{rows}\n")
            for r in range(rows):
                n = pb.hpolys[b, i, r]
                fh.write(" ".join(repr(float(v)) for v in (n[0], n[1], n[2], -n[3])) + "\n")


def test_binaries_built_and_fail_loudly_without_gpu(tmp_path):
    """Host headers compile (against the Eigen stand-in; the reference lbfgs.hpp variant too) and, on a box
    without a CUDA device, a compute call ends with an error message, not with a CPU result."""
    assert os.path.exists(BIN), "run __graft_entry__.build()"
    assert subprocess.run([BIN], capture_output=True).returncode == 2
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is covered on the CPU box")
    pb = synth.make_problems(1, N=5, K=8, S=3)
    p = tmp_path / "p.txt"
    _write_problem(p, pb, 0)
    r = subprocess.run([BIN, str(p)], capture_output=True, text=True)
    assert r.returncode == 3 and "mincob" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,seedoff", [(5, 16, 0), (8, 16, 3), (3, 9, 7)])
def test_cpp_host_side_matches_oracle(tmp_path, oracle, N, K, seedoff):
    pb = synth.make_problems(1, N=N, K=K, S=3, first=1000 + seedoff, ragged_rows=(K == 9))
    p = tmp_path / "p.txt"
    _write_problem(p, pb, 0)
    exe = BIN_REF if os.path.exists(BIN_REF) else BIN
    r = subprocess.run([exe, str(p)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    prm = default_params(3)
    T0 = pb.T0[0].astype(np.float32).astype(np.float64)          # the planner's times are float32 (learning_planner.hpp:179)
    tol = 1e-9
    # MINCO_S3NU call sequence
    ref = oracle.minco_forward(3, pb.head[0], pb.tail[0], pb.q0[0], T0)
    assert abs(out["energy"] - ref["energy"]) <= tol * abs(ref["energy"])
    c = np.array(out["coeffs_asc"]).reshape(6 * N, 3)
    assert np.abs(c - ref["coeffs"]).max() <= tol * np.abs(ref["coeffs"]).max()
    gq, gT = oracle.minco_propagate(3, pb.head[0], pb.tail[0], pb.q0[0], T0, ref["gdC"], ref["gdT"])
    assert np.abs(np.array(out["gradByPoints"]).reshape(N - 1, 3) - gq).max() <= tol * np.abs(gq).max()
    assert np.abs(np.array(out["gradByTimes"]) - gT).max() <= tol * np.abs(gT).max()
    assert np.abs(np.array(out["traj_desc"]).reshape(N, 3, 6) - ref["flat"]).max() <= tol * np.abs(ref["flat"]).max()
    # costFunctional through the lbfgs_evaluate_t signature
    pb32 = synth.ProblemBatch(3, N, K, pb.head, pb.tail, pb.hpolys, pb.hrows, pb.q0, T0[None, :])
    x0 = pb32.x0()
    np.testing.assert_allclose(np.array(out["x0"]), x0[0], rtol=1e-14, atol=1e-14)
    fo, go = oracle.cost_batch(prm, pb32, x0)
    assert abs(out["f0"] - fo[0]) <= tol * abs(fo[0])
    assert np.abs(np.array(out["g0"]) - go[0]).max() <= tol * np.abs(go[0]).max()
    # device-side lbfgs_optimize: reported cost is the oracle's cost at the returned x; success code
    assert out["dev_status"] >= 0
    xd = np.array(out["dev_x"])[None, :]
    fd, _ = oracle.cost_batch(prm, pb32, xd)
    assert abs(out["dev_f"] - fd[0]) <= tol * abs(fd[0])
    cpu = oracle.optimize_batch_ref(prm, pb32)
    assert abs(out["dev_f"] - cpu["f"][0]) <= 2e-2 * abs(cpu["f"][0])
    # the reference's lbfgs.hpp driving the GPU cost callback: same algorithm, same cost => same optimum
    if "host_lbfgs_ret" in out:
        assert out["host_lbfgs_ret"] >= 0
        assert abs(out["host_lbfgs_f"] - cpu["f"][0]) <= 2e-2 * abs(cpu["f"][0])
        fh, _ = oracle.cost_batch(prm, pb32, np.array(out["host_lbfgs_x"])[None, :])
        assert abs(out["host_lbfgs_f"] - fh[0]) <= tol * abs(fh[0])
    # drop-in solve(), TimeMode::Fixed (the literal replacement of qp_solver.solve: times are data): durations come back
    # bit-identical, the pieces use exactly those durations, boundary states and C^2 continuity hold, and the cost is the
    # oracle's fixed-time optimum
    assert out["solvefix_ok"] is True
    np.testing.assert_array_equal(np.array(out["solvefix_times"]), T0)
    ff = np.array(out["solvefix_flat"]).reshape(N, 3, 6)
    np.testing.assert_allclose(ff[0, :, 5], pb.head[0, 0], atol=1e-9)
    pw = T0[:, None] ** np.arange(5, -1, -1)[None, :]
    ends = (ff * pw[:, None, :]).sum(axis=2)
    np.testing.assert_allclose(ends[N - 1], pb.tail[0, 0], atol=1e-8)
    np.testing.assert_allclose(ends[:-1], ff[1:, :, 5], atol=1e-8)                      # position continuity at the junctions
    dcoef = ff[:, :, :5] * np.arange(5, 0, -1)[None, None, :]
    vend = (dcoef * pw[:, None, 1:]).sum(axis=2)
    np.testing.assert_allclose(vend[:-1], ff[1:, :, 4], atol=1e-7)                      # velocity continuity
    # drop-in solve(), TimeMode::Optimize: flat layout idx = i*3*6 + j*6 + k, descending powers; times written back
    assert out["solve_ok"] is True
    flat = np.array(out["solve_flat"]).reshape(N, 3, 6)
    Ts = np.array(out["solve_times"])
    assert (Ts > 0).all()
    np.testing.assert_allclose(flat[0, :, 5], pb.head[0, 0], atol=1e-9)                # p(0) of piece 0 = start
    end = (flat[N - 1] * (Ts[N - 1, None] ** np.arange(5, -1, -1))[None, :]).sum(axis=1)
    np.testing.assert_allclose(end, pb.tail[0, 0], atol=1e-4)                          # float32 durations
