"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of oracle/_ref.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs; never by allocnet_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
EVAL_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, _dp, _dp, C.c_int)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible: g++ only) and, when /root/reference is present, _ref."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "minco_oracle.hpp", "lbfgs_oracle.hpp")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    ref_missing = (not os.path.exists(os.path.join(_HERE, "_ref", "libref_lbfgs.so"))
                   or not os.path.exists(os.path.join(_HERE, "liboracle_strict.so")))
    if stale or ref_missing:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)


def _ptr(a, typ=_dp):
    return None if a is None else a.ctypes.data_as(typ)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self, strict: bool = False):
        """strict=True loads the -ffp-contract=off build (bit-comparable with oracle/_ref)."""
        build()
        self.lib = C.CDLL(os.path.join(_HERE, "liboracle_strict.so" if strict else "liboracle.so"))
        L = self.lib
        L.orc_cost_create.restype = C.c_void_p
        L.orc_cost_thunk_address.restype = C.c_void_p
        L.orc_cost_thunk.restype = C.c_double
        L.orc_forward_t.restype = L.orc_backward_t.restype = L.orc_backward_grad_t.restype = C.c_double
        L.orc_forward_t.argtypes = [C.c_double]
        L.orc_backward_t.argtypes = [C.c_double]
        L.orc_backward_grad_t.argtypes = [C.c_double, C.c_double]
        L.orc_smoothed_l1.argtypes = [C.c_double, C.c_double, _dp, _dp]
        L.orc_cost_destroy.argtypes = [C.c_void_p]
        L.orc_cost_thunk.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.orc_cost_flat.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.orc_lbfgs_optimize.argtypes = [C.c_int, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p, _ip, _ip]
        self.ref = None
        ref_so = os.path.join(_HERE, "_ref", "libref_lbfgs.so")
        if os.path.exists(ref_so):
            self.ref = C.CDLL(ref_so)
            self.ref.ref_lbfgs_optimize.argtypes = [C.c_int, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p, _ip, _ip]
            self.ref.ref_lbfgs_strerror.restype = C.c_char_p

    # ---- MINCO pieces ------------------------------------------------------
    def minco_forward(self, S, head, tail, inPs, ts):
        """-> dict(coeffs [2S*N][3], energy, gdC [2S*N][3], gdT [N], flat [N][3][2S])."""
        ts = _f64(ts); N = ts.shape[0]
        head, tail, inPs = _f64(head), _f64(tail), _f64(inPs if N > 1 else np.zeros((1, 3)))
        coeffs = np.zeros((2 * S * N, 3)); gdC = np.zeros((2 * S * N, 3)); gdT = np.zeros(N)
        flat = np.zeros((N, 3, 2 * S)); e = C.c_double(0.0)
        rc = self.lib.orc_minco_forward(S, N, _ptr(head), _ptr(tail), _ptr(inPs), _ptr(ts), _ptr(coeffs),
                                        C.byref(e), _ptr(gdC), _ptr(gdT), _ptr(flat))
        assert rc == 0
        return dict(coeffs=coeffs, energy=e.value, gdC=gdC, gdT=gdT, flat=flat)

    def minco_propagate(self, S, head, tail, inPs, ts, gdC, gdT):
        ts = _f64(ts); N = ts.shape[0]
        head, tail, inPs = _f64(head), _f64(tail), _f64(inPs if N > 1 else np.zeros((1, 3)))
        gdC, gdT = _f64(gdC), _f64(gdT)
        gq = np.zeros((max(N - 1, 1), 3)); gT = np.zeros(N)
        rc = self.lib.orc_minco_propagate(S, N, _ptr(head), _ptr(tail), _ptr(inPs), _ptr(ts), _ptr(gdC),
                                          _ptr(gdT), _ptr(gq), _ptr(gT))
        assert rc == 0
        return gq[: N - 1], gT

    def banded_solve(self, dense, p, q, b, adj=False):
        dense = _f64(dense); b = _f64(b).copy()
        n = dense.shape[0]; m = b.shape[1]
        self.lib.orc_banded_solve(n, p, q, _ptr(dense), _ptr(b), m, int(adj))
        return b

    def smoothed_l1(self, mu, x):
        f, df = C.c_double(0.0), C.c_double(0.0)
        hit = self.lib.orc_smoothed_l1(mu, x, C.byref(f), C.byref(df))
        return bool(hit), f.value, df.value

    # ---- batched cost / optimize ------------------------------------------
    def cost_batch(self, params, pb, x, nthreads=1):
        """f [B], g [B][n] of the cost functional at x (CPU, fp64)."""
        x = _f64(x); B, n = x.shape
        f = np.zeros(B); g = np.zeros((B, n))
        hp = pb.hpolys if pb.K > 0 else None
        rc = self.lib.orc_cost_batch(C.byref(params), B, pb.N, _ptr(pb.head), _ptr(pb.tail), _ptr(hp),
                                     _ptr(pb.hrows, _ip), pb.K, _ptr(x), _ptr(f), _ptr(g), nthreads)
        assert rc == 0
        return f, g

    def optimize_batch(self, params, pb, x0=None, nthreads=1):
        x = _f64(pb.x0() if x0 is None else x0).copy(); B, n = x.shape
        S, N = params.S, pb.N
        f = np.zeros(B); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
        evals = np.zeros(B, np.int32); coeffs = np.zeros((B, N, 3, 2 * S)); T = np.zeros((B, N))
        hp = pb.hpolys if pb.K > 0 else None
        rc = self.lib.orc_optimize_batch(C.byref(params), B, N, _ptr(pb.head), _ptr(pb.tail), _ptr(hp),
                                         _ptr(pb.hrows, _ip), pb.K, _ptr(x), _ptr(f), _ptr(status, _ip),
                                         _ptr(iters, _ip), _ptr(evals, _ip), _ptr(coeffs), _ptr(T), nthreads)
        assert rc == 0
        return dict(x=x, f=f, status=status, iters=iters, evals=evals, coeffs=coeffs, T=T)

    def optimize_batch_ref(self, params, pb, x0=None, nthreads=1):
        """optimize_batch with the L-BFGS driver of oracle/_ref (the reference's lbfgs.hpp compiled verbatim)
        around the restated MINCO cost functional.  Falls back to the restated driver if _ref is absent."""
        if self.ref is None:
            out = self.optimize_batch(params, pb, x0, nthreads)
            out["driver"] = "oracle/lbfgs_oracle.hpp (restated)"
            return out
        x = _f64(pb.x0() if x0 is None else x0).copy(); B, n = x.shape
        S, N = params.S, pb.N
        f = np.zeros(B); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
        evals = np.zeros(B, np.int32); coeffs = np.zeros((B, N, 3, 2 * S)); T = np.zeros((B, N))
        hp = pb.hpolys if pb.K > 0 else None
        drv = C.cast(self.ref.ref_lbfgs_optimize, C.c_void_p)
        fn = self.lib.orc_optimize_batch_with
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _dp, _ip, _ip, _ip,
                       _dp, _dp, C.c_int]
        rc = fn(drv, C.cast(C.byref(params), C.c_void_p), B, N, _ptr(pb.head), _ptr(pb.tail), _ptr(hp),
                _ptr(pb.hrows, _ip), pb.K, _ptr(x), _ptr(f), _ptr(status, _ip), _ptr(iters, _ip), _ptr(evals, _ip),
                _ptr(coeffs), _ptr(T), nthreads)
        assert rc == 0
        return dict(x=x, f=f, status=status, iters=iters, evals=evals, coeffs=coeffs, T=T,
                    driver="oracle/_ref (reference gcopter/lbfgs.hpp, verbatim)")

    def ref_qp_build(self, iniPVA, finPVA, polys, times32, order=3, res=20, vel_box=4.0, acc_box=6.0):
        """The matrices the reference's own QPSolver::solve (planner/qp_solver.hpp, compiled verbatim into
        oracle/_ref/libref_qp.so) hands to OSQP: dict(hessian [n][n], constraints [m][n], lower [m], upper [m]), or None
        when that library was not built (no reference tree).  iniPVA / finPVA 3x3 (rows axis, columns P,V,A);
        polys [seg][rows][4] rows [n, b]; times32 float32 [seg]."""
        so = os.path.join(_HERE, "_ref", "libref_qp.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        polys = _f64(polys); seg, rows = polys.shape[0], polys.shape[1]
        ini, fin = _f64(iniPVA), _f64(finPVA)
        t32 = np.ascontiguousarray(times32, dtype=np.float32)
        d = 2 * order
        n = seg * 3 * d
        m = (6 + order * (seg - 1)) * 3 + res * (rows * seg) + res * 12 * seg
        H = np.zeros((n, n)); A = np.zeros((m, n)); lo = np.zeros(m); up = np.zeros(m)
        nn, mm = C.c_int(0), C.c_int(0)
        rc = L.ref_qp_build(seg, rows, order, res, C.c_double(vel_box), C.c_double(acc_box), _ptr(ini), _ptr(fin), _ptr(polys),
                            t32.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nn), C.byref(mm), _ptr(H), _ptr(A), _ptr(lo), _ptr(up))
        assert rc == 0 and nn.value == n and mm.value == m, (rc, nn.value, n, mm.value, m)
        return dict(hessian=H, constraints=A, lower=lo, upper=up, n_eq=(6 + order * (seg - 1)) * 3)

    def hardware_threads(self):
        return int(self.lib.orc_hardware_threads())

    # ---- single-problem cost instance (lbfgs callback ABI) -------------------
    def cost_instance(self, params, pb, b):
        one = pb.slice(b, b + 1)
        hp = one.hpolys if one.K > 0 else None
        inst = self.lib.orc_cost_create(C.byref(params), one.N, _ptr(one.head), _ptr(one.tail), _ptr(hp),
                                        _ptr(one.hrows, _ip), one.K)
        return _CostInstance(self, inst, one)

    def lbfgs(self, n, x0, eval_ptr, inst, params, which="oracle"):
        """Run the restated ('oracle') or the verbatim reference ('ref') L-BFGS on a raw C callback."""
        x = _f64(x0).copy(); f = C.c_double(0.0); it = C.c_int(0); ev = C.c_int(0)
        fn = self.lib.orc_lbfgs_optimize if which == "oracle" else self.ref.ref_lbfgs_optimize
        ret = fn(n, _ptr(x), C.byref(f), eval_ptr, inst, C.cast(C.byref(params), C.c_void_p), C.byref(it), C.byref(ev))
        return dict(x=x, f=f.value, ret=int(ret), iters=it.value, evals=ev.value)


class _CostInstance:
    def __init__(self, orc, inst, one):
        self.orc, self.inst, self._keep = orc, inst, one
        self.thunk = orc.lib.orc_cost_thunk_address()

    def __call__(self, x):
        x = _f64(x); g = np.zeros_like(x)
        f = self.orc.lib.orc_cost_thunk(self.inst, _ptr(x), _ptr(g), x.shape[0])
        return f, g

    def close(self):
        if self.inst:
            self.orc.lib.orc_cost_destroy(self.inst)
            self.inst = None
