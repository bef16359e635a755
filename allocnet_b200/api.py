"""Host-side binding of the C-ABI in include/mincob.h (allocnet_b200/libmincob.so).

This is the Python mirror of the reference-side interface for the trajectory back-end slot
(`qp_solver.solve(...)` at src/planner/include/planner/learning_planner.hpp:196 and the
`lbfgs_optimize` / `lbfgs_evaluate_t` ABI of src/planner/include/gcopter/lbfgs.hpp:200-202,434-440),
batched over B problems.  Method names follow the upstream MINCO / GCOPTER API restated in
SURVEY.md Appendix A/B (setParameters, getEnergy, propogateGrad, costFunctional, optimize).

There is no CPU fallback: if the CUDA library is missing, or no device is present, the calls
raise.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .params import MincobParams, default_params

_HERE = os.path.dirname(os.path.abspath(__file__))
# MINCOB_LIBRARY selects another build of the same library (kernel A/B experiments, tools/variants.sh)
LIB_PATH = os.environ.get("MINCOB_LIBRARY") or os.path.join(_HERE, "libmincob.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

SYMBOLS = [
    "mincob_version", "mincob_default_params", "mincob_create", "mincob_destroy", "mincob_set_params",
    "mincob_set_stream", "mincob_synchronize", "mincob_last_error", "mincob_strerror", "mincob_lbfgs_strerror",
    "mincob_set_problems", "mincob_set_problems_device", "mincob_evaluate", "mincob_evaluate_device",
    "mincob_optimize", "mincob_optimize_device", "mincob_last_kernel_ms", "mincob_minco_forward",
    "mincob_minco_propagate", "mincob_nccl_unique_id", "mincob_comm_init", "mincob_allgather_device",
    "mincob_comm_destroy", "mincob_optimize_sharded", "mincob_host_alloc", "mincob_host_free",
    "mincob_check_feasibility", "mincob_check_feasibility_device", "mincob_measure_fp64_peak",
    "mincob_max_rates", "mincob_max_rates_device", "mincob_minco_forward_device", "mincob_minco_propagate_device",
    "mincob_optimize_sharded_local", "mincob_gathered_device", "mincob_last_mapping", "mincob_host_register",
    "mincob_host_unregister", "mincob_set_problems_async",
]


class MincobError(RuntimeError):
    pass


_lib = None
_libs = {}


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libmincob.so (or another build of it, e.g. libmincob_strict.so); raises (loudly) when it has not been built."""
    global _lib
    if path is None and _lib is not None:
        return _lib
    if path is not None and path in _libs:
        return _libs[path]
    lib_path = path or LIB_PATH
    if not os.path.exists(lib_path):
        raise MincobError(f"{lib_path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(lib_path)
    L.mincob_last_error.restype = C.c_char_p
    L.mincob_strerror.restype = C.c_char_p
    L.mincob_lbfgs_strerror.restype = C.c_char_p
    L.mincob_create.argtypes = [C.POINTER(_vp), C.POINTER(MincobParams), C.c_int]
    L.mincob_destroy.argtypes = [_vp]
    L.mincob_set_params.argtypes = [_vp, C.POINTER(MincobParams)]
    L.mincob_set_stream.argtypes = [_vp, _vp]
    L.mincob_synchronize.argtypes = [_vp]
    L.mincob_last_error.argtypes = [_vp]
    L.mincob_set_problems.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]
    L.mincob_set_problems_device.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]
    if hasattr(L, "mincob_set_problems_async"):
        L.mincob_set_problems_async.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]
    L.mincob_evaluate.argtypes = [_vp, _vp, _vp, _vp]
    L.mincob_evaluate_device.argtypes = [_vp, _vp, _vp, _vp]
    L.mincob_optimize.argtypes = [_vp] + [_vp] * 7
    L.mincob_optimize_device.argtypes = [_vp] + [_vp] * 7
    L.mincob_last_kernel_ms.argtypes = [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.mincob_minco_forward.argtypes = [_vp, C.c_int, C.c_int] + [_vp] * 9
    L.mincob_minco_propagate.argtypes = [_vp, C.c_int, C.c_int] + [_vp] * 8
    L.mincob_nccl_unique_id.argtypes = [_vp]
    L.mincob_comm_init.argtypes = [_vp, C.c_int, C.c_int, _vp]
    L.mincob_allgather_device.argtypes = [_vp, _vp, _vp, C.c_int64]
    L.mincob_comm_destroy.argtypes = [_vp]
    L.mincob_optimize_sharded.argtypes = [_vp] + [_vp] * 7
    L.mincob_optimize_sharded_local.argtypes = [_vp] + [_vp] * 7
    L.mincob_gathered_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]
    L.mincob_host_alloc.argtypes = [C.POINTER(_vp), C.c_uint64]
    L.mincob_host_free.argtypes = [_vp]
    if hasattr(L, "mincob_host_register"):
        L.mincob_host_register.argtypes = [_vp, C.c_uint64]
        L.mincob_host_unregister.argtypes = [_vp]
    L.mincob_check_feasibility.argtypes = [_vp, _vp, _vp, C.c_int, _vp]
    L.mincob_check_feasibility_device.argtypes = [_vp, _vp, _vp, C.c_int, _vp]
    L.mincob_measure_fp64_peak.argtypes = [_vp, C.POINTER(C.c_double)]
    L.mincob_max_rates.argtypes = [_vp, _vp, _vp, _vp]
    L.mincob_minco_forward_device.argtypes = [_vp, C.c_int, C.c_int] + [_vp] * 9
    L.mincob_minco_propagate_device.argtypes = [_vp, C.c_int, C.c_int] + [_vp] * 8
    L.mincob_max_rates_device.argtypes = [_vp, _vp, _vp, _vp]
    if hasattr(L, "mincob_last_mapping"):   # absent from older builds selected with MINCOB_LIBRARY (A/B runs)
        L.mincob_last_mapping.argtypes = [_vp, C.POINTER(C.c_int)]
    if path is None:
        _lib = L
    else:
        _libs[path] = L
    return L


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dev_ptr(t):
    """torch CUDA tensor (contiguous) -> raw device pointer."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise MincobError("device entry points need contiguous CUDA tensors")
    return C.c_void_p(t.data_ptr())


class MincoBatch:
    """Batched MINCO optimizer handle (one per GPU).  Mirrors GCOPTER_PolytopeSFC: setup -> optimize."""

    def __init__(self, params: MincobParams | None = None, device: int = 0, lib_path: str | None = None):
        self.L = load_library(lib_path)
        self.params = params if params is not None else default_params()
        self.h = _vp()
        rc = self.L.mincob_create(C.byref(self.h), C.byref(self.params), int(device))
        if rc != 0:
            raise MincobError(f"mincob_create failed: {self.L.mincob_strerror(rc).decode()} (rc={rc})")
        self.device = int(device)
        self.B = self.N = self.K = 0
        self._keep = None

    # -- plumbing -------------------------------------------------------------------------
    @staticmethod
    def _want(a, shape, name):
        if tuple(a.shape) != tuple(shape):
            raise MincobError(f"{name} must have shape {tuple(shape)}, got {tuple(a.shape)}")
        return a

    def _check(self, rc):
        if rc != 0:
            raise MincobError(f"{self.L.mincob_strerror(rc).decode()}: {self.L.mincob_last_error(self.h).decode()} (rc={rc})")

    def close(self):
        if self.h:
            self.L.mincob_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def S(self):
        return int(self.params.S)

    @property
    def nvars(self):
        return self.N + 3 * (self.N - 1)

    def set_params(self, params: MincobParams):
        self._check(self.L.mincob_set_params(self.h, C.byref(params)))
        self.params = params

    def set_stream(self, cuda_stream: int | None):
        """None = the handle's own non-blocking stream.  An integer is passed through as the cudaStream_t; note that
        torch's default stream reports 0, which the C-ABI also reads as "own stream": pass cudaStreamLegacy (0x1) or
        cudaStreamPerThread (0x2) to run on a default stream (allocnet_b200/autograd.py does)."""
        self._check(self.L.mincob_set_stream(self.h, _vp(cuda_stream) if cuda_stream is not None else None))

    def synchronize(self):
        self._check(self.L.mincob_synchronize(self.h))

    def last_kernel_ms(self):
        ms, n = C.c_float(0), C.c_int(0)
        self._check(self.L.mincob_last_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def last_mapping(self) -> int:
        """params.MAP_THROUGHPUT / MAP_LATENCY: what the last optimize call launched."""
        m = C.c_int(0)
        self._check(self.L.mincob_last_mapping(self.h, C.byref(m)))
        return int(m.value)

    # -- problems -------------------------------------------------------------------------
    def set_problems(self, pb):
        """pb: allocnet_b200.synth.ProblemBatch (host numpy arrays in the C-ABI layouts)."""
        if pb.S != self.S:
            raise MincobError(f"problem batch has S={pb.S}, handle has S={self.S}")
        B, N, K, S = int(pb.B), int(pb.N), int(pb.K), self.S
        # the C side reads exactly B*S*3 / B*N*K*4 doubles and B*N int32: coerce (dtype, contiguity) and check shapes here
        head = self._want(_f64(pb.head), (B, S, 3), "head")
        tail = self._want(_f64(pb.tail), (B, S, 3), "tail")
        hp = self._want(_f64(pb.hpolys), (B, N, K, 4), "hpolys") if K > 0 else None
        hr = self._want(np.ascontiguousarray(pb.hrows, dtype=np.int32), (B, N), "hrows") if K > 0 else None
        self._check(self.L.mincob_set_problems(self.h, B, N, K, _np_ptr(head), _np_ptr(tail), _np_ptr(hp), _np_ptr(hr)))
        self.B, self.N, self.K = B, N, K

    def set_problems_async(self, pb):
        """set_problems without waiting: pb's arrays must be page-locked, C-contiguous fp64 / int32 in the C-ABI layouts
        (pinned_empty) and untouched until the next optimize call has returned; that call overlaps the upload."""
        B, N, K, S = int(pb.B), int(pb.N), int(pb.K), self.S
        for name, a, shape, dt in (("head", pb.head, (B, S, 3), np.float64), ("tail", pb.tail, (B, S, 3), np.float64)) + \
                ((("hpolys", pb.hpolys, (B, N, K, 4), np.float64), ("hrows", pb.hrows, (B, N), np.int32)) if K > 0 else ()):
            if a.dtype != dt or not a.flags.c_contiguous or tuple(a.shape) != shape:
                raise MincobError(f"{name} must be a C-contiguous {np.dtype(dt).name} array of shape {shape}")
        self._check(self.L.mincob_set_problems_async(self.h, B, N, K, _np_ptr(pb.head), _np_ptr(pb.tail),
                                                     _np_ptr(pb.hpolys if K > 0 else None), _np_ptr(pb.hrows if K > 0 else None)))
        self.B, self.N, self.K = B, N, K

    def set_problems_device(self, B, N, K, head, tail, hpolys=None, hrows=None):
        self._check(self.L.mincob_set_problems_device(self.h, B, N, K, _dev_ptr(head), _dev_ptr(tail),
                                                      _dev_ptr(hpolys), _dev_ptr(hrows)))
        self.B, self.N, self.K = B, N, K
        self._keep = (head, tail, hpolys, hrows)

    # -- costFunctional ---------------------------------------------------------------------
    def evaluate(self, x):
        x = _f64(x)
        if x.shape != (self.B, self.nvars):
            raise MincobError(f"x must be [{self.B}][{self.nvars}]")
        f = np.empty(self.B); g = np.empty_like(x)
        self._check(self.L.mincob_evaluate(self.h, _np_ptr(x), _np_ptr(f), _np_ptr(g)))
        return f, g

    def evaluate_device(self, x, f, g):
        self._check(self.L.mincob_evaluate_device(self.h, _dev_ptr(x), _dev_ptr(f), _dev_ptr(g)))

    # -- lbfgs_optimize on costFunctional + getTrajectory -------------------------------------
    def optimize(self, x0, want_coeffs=True):
        x = _f64(x0).copy()
        if x.shape != (self.B, self.nvars):
            raise MincobError(f"x must be [{self.B}][{self.nvars}]")
        B, N, S = self.B, self.N, self.S
        f = np.zeros(B); status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); evals = np.zeros(B, np.int32)
        coeffs = np.zeros((B, N, 3, 2 * S)) if want_coeffs else None
        T = np.zeros((B, N)) if want_coeffs else None
        self._check(self.L.mincob_optimize(self.h, _np_ptr(x), _np_ptr(f), _np_ptr(status), _np_ptr(iters),
                                           _np_ptr(evals), _np_ptr(coeffs), _np_ptr(T)))
        return dict(x=x, f=f, status=status, iters=iters, evals=evals, coeffs=coeffs, T=T)

    def optimize_host_buffers(self, x, f, status, iters, evals, coeffs, T):
        """In-place variant on caller-owned (e.g. pinned) host arrays; used by bench.py's e2e leg."""
        self._check(self.L.mincob_optimize(self.h, _np_ptr(x), _np_ptr(f), _np_ptr(status), _np_ptr(iters),
                                           _np_ptr(evals), _np_ptr(coeffs), _np_ptr(T)))

    def optimize_device(self, x, f=None, status=None, iters=None, evals=None, coeffs=None, T=None):
        self._check(self.L.mincob_optimize_device(self.h, _dev_ptr(x), _dev_ptr(f), _dev_ptr(status), _dev_ptr(iters),
                                                  _dev_ptr(evals), _dev_ptr(coeffs), _dev_ptr(T)))

    # -- MINCO_S3NU / S4NU building blocks ---------------------------------------------------
    def minco_forward(self, head, tail, inPs, ts):
        """setParameters + getCoeffs/getEnergy/getEnergyPartialGradBy{Coeffs,Times}/getTrajectory for B problems."""
        ts = _f64(ts); B, N = ts.shape; S = self.S
        head, tail = _f64(head), _f64(tail)
        inPs = _f64(inPs) if N > 1 else np.zeros((B, 1, 3))
        coeffs = np.zeros((B, 2 * S * N, 3)); energy = np.zeros(B); gdC = np.zeros((B, 2 * S * N, 3))
        gdT = np.zeros((B, N)); flat = np.zeros((B, N, 3, 2 * S))
        self._check(self.L.mincob_minco_forward(self.h, B, N, _np_ptr(head), _np_ptr(tail), _np_ptr(inPs), _np_ptr(ts),
                                                _np_ptr(coeffs), _np_ptr(energy), _np_ptr(gdC), _np_ptr(gdT), _np_ptr(flat)))
        return dict(coeffs=coeffs, energy=energy, gdC=gdC, gdT=gdT, flat=flat)

    def minco_propagate(self, head, tail, inPs, ts, gdC, gdT):
        """propogateGrad (upstream spelling) for B problems -> gradByPoints [B][N-1][3], gradByTimes [B][N]."""
        ts = _f64(ts); B, N = ts.shape
        head, tail, gdC, gdT = _f64(head), _f64(tail), _f64(gdC), _f64(gdT)
        inPs = _f64(inPs) if N > 1 else np.zeros((B, 1, 3))
        gq = np.zeros((B, max(N - 1, 1), 3)); gT = np.zeros((B, N))
        self._check(self.L.mincob_minco_propagate(self.h, B, N, _np_ptr(head), _np_ptr(tail), _np_ptr(inPs), _np_ptr(ts),
                                                  _np_ptr(gdC), _np_ptr(gdT), _np_ptr(gq), _np_ptr(gT)))
        return gq[:, : N - 1], gT

    # -- feasibility report (sampled Piece::getMaxVelRate / getMaxAccRate / corridor residual) ------
    def check_feasibility(self, coeffs, T, samples: int = 64):
        """-> [B][4]: max |v|, max |a|, max |j|, max_k(n_k.p + d_k) over samples+1 points per piece."""
        coeffs = self._want(_f64(coeffs), (self.B, self.N, 3, 2 * self.S), "coeffs")
        T = self._want(_f64(T), (self.B, self.N), "T")
        rep = np.zeros((self.B, 4))
        self._check(self.L.mincob_check_feasibility(self.h, _np_ptr(coeffs), _np_ptr(T), int(samples), _np_ptr(rep)))
        return rep

    def check_feasibility_device(self, coeffs, T, samples, report):
        self._check(self.L.mincob_check_feasibility_device(self.h, _dev_ptr(coeffs), _dev_ptr(T), int(samples), _dev_ptr(report)))

    def minco_forward_device(self, B, N, head, tail, inPs, ts, coeffs_asc=None, energy=None, gdC=None, gdT=None, flat=None):
        """CUDA tensors in the layouts of minco_forward; None outputs are skipped; enqueues only."""
        self._check(self.L.mincob_minco_forward_device(self.h, int(B), int(N), _dev_ptr(head), _dev_ptr(tail), _dev_ptr(inPs),
                                                       _dev_ptr(ts), _dev_ptr(coeffs_asc), _dev_ptr(energy), _dev_ptr(gdC),
                                                       _dev_ptr(gdT), _dev_ptr(flat)))

    def minco_propagate_device(self, B, N, head, tail, inPs, ts, gdC, gdT, gradByPoints, gradByTimes):
        self._check(self.L.mincob_minco_propagate_device(self.h, int(B), int(N), _dev_ptr(head), _dev_ptr(tail), _dev_ptr(inPs),
                                                         _dev_ptr(ts), _dev_ptr(gdC), _dev_ptr(gdT), _dev_ptr(gradByPoints),
                                                         _dev_ptr(gradByTimes)))

    def max_rates(self, coeffs, T):
        """[B][3] exact max |v|, |a|, |j| per trajectory (Trajectory<D>::getMaxVelRate / getMaxAccRate, + jerk)."""
        coeffs = self._want(_f64(coeffs), (self.B, self.N, 3, 2 * self.S), "coeffs")
        T = self._want(_f64(T), (self.B, self.N), "T")
        rates = np.empty((self.B, 3), dtype=np.float64)
        self._check(self.L.mincob_max_rates(self.h, _np_ptr(coeffs), _np_ptr(T), _np_ptr(rates)))
        return rates

    def max_rates_device(self, coeffs, T, rates):
        self._check(self.L.mincob_max_rates_device(self.h, _dev_ptr(coeffs), _dev_ptr(T), _dev_ptr(rates)))

    def measure_fp64_peak(self) -> float:
        """TFLOP/s of independent DFMA chains on this device (the fp64-pipe ceiling bench.py reports against)."""
        v = C.c_double(0.0)
        self._check(self.L.mincob_measure_fp64_peak(self.h, C.byref(v)))
        return float(v.value)

    # -- multi-GPU ---------------------------------------------------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.L.mincob_nccl_unique_id(buf)
        if rc != 0:
            raise MincobError(f"mincob_nccl_unique_id: {self.L.mincob_strerror(rc).decode()}")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.L.mincob_comm_init(self.h, nranks, rank, buf))

    def allgather_device(self, send, recv, count_per_rank: int):
        self._check(self.L.mincob_allgather_device(self.h, _dev_ptr(send), _dev_ptr(recv), int(count_per_rank)))


    def optimize_sharded_local_host_buffers(self, x, f, status, iters, evals, coeffs_local, T):
        """Sharded job on a rank that does not consume the gathered coefficients: same collective, only this rank's own
        results come back to the host (mincob_optimize_sharded_local)."""
        self._check(self.L.mincob_optimize_sharded_local(self.h, _np_ptr(x), _np_ptr(f), _np_ptr(status), _np_ptr(iters),
                                                         _np_ptr(evals), _np_ptr(coeffs_local), _np_ptr(T)))

    def gathered_device(self):
        """(device pointer, count of doubles) of the coefficients gathered by the last sharded optimize call."""
        p = _vp(); n = C.c_int64(0)
        self._check(self.L.mincob_gathered_device(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def optimize_sharded_host_buffers(self, x, f, status, iters, evals, coeffs_all, T):
        """Config 5 on caller-owned host arrays: optimize this rank's shard, all-gather the coefficients."""
        self._check(self.L.mincob_optimize_sharded(self.h, _np_ptr(x), _np_ptr(f), _np_ptr(status), _np_ptr(iters),
                                                   _np_ptr(evals), _np_ptr(coeffs_all), _np_ptr(T)))


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array over page-locked host memory from mincob_host_alloc (freed when the array dies)."""
    import weakref
    L = load_library()
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    p = _vp()
    rc = L.mincob_host_alloc(C.byref(p), max(nbytes, 8))
    if rc != 0:
        raise MincobError(f"mincob_host_alloc({nbytes}): {L.mincob_strerror(rc).decode()}")
    buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, L.mincob_host_free, _vp(p.value))
    return arr


def host_register(arr: np.ndarray) -> None:
    """Page-lock an existing numpy buffer (mincob_host_register), e.g. one over multiprocessing.shared_memory."""
    L = load_library()
    rc = L.mincob_host_register(_vp(arr.ctypes.data), arr.nbytes)
    if rc != 0:
        raise MincobError(f"mincob_host_register({arr.nbytes}): {L.mincob_strerror(rc).decode()}")


def host_unregister(arr: np.ndarray) -> None:
    load_library().mincob_host_unregister(_vp(arr.ctypes.data))


def lbfgs_strerror(status: int) -> str:
    return load_library().mincob_lbfgs_strerror(int(status)).decode()
