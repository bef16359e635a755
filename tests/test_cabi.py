"""The C-ABI library must load on a CPU-only box and export every symbol include/mincob.h declares.
No compute call is made here (there is no CPU path: compute entry points need a CUDA device)."""
import ctypes as C
import os
import re

import pytest

from allocnet_b200 import api
from allocnet_b200.params import MincobParams, default_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from allocnet_b200.build import build_library
    build_library()
    return api.load_library()


def _declared():
    src = open(os.path.join(ROOT, "include", "mincob.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mincob_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mincob.h but not exported"
    assert sorted(api.SYMBOLS) == names


def test_params_struct_matches_oracle_and_defaults(lib, oracle):
    assert C.sizeof(MincobParams) == oracle.lib.orc_params_size()
    p = MincobParams()
    assert lib.mincob_default_params(C.byref(p), 3) == 0
    q = default_params(3)
    assert bytes(p) == bytes(q)
    assert lib.mincob_default_params(C.byref(p), 5) != 0


def test_error_strings(lib):
    assert b"CPU" in lib.mincob_strerror(-2)          # "no CPU path"
    from allocnet_b200 import params as P
    assert b"convergence" in lib.mincob_lbfgs_strerror(P.LBFGS_CONVERGENCE).lower()
    assert b"line search" in lib.mincob_lbfgs_strerror(P.LBFGSERR_MAXIMUMLINESEARCH).lower()
    assert lib.mincob_version() >= 100


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, not fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.MincobError):
        api.MincoBatch(default_params(3), device=0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "allocnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "minco_oracle.hpp" not in txt, f


def test_null_and_state_errors_without_a_device(lib):
    """Entry points that take a handle refuse NULL handles with MINCOB_E_INVALID before touching CUDA (what a binding in
    another language sees first), and the error strings are stable."""
    import ctypes as C
    E_INVALID = lib.mincob_evaluate(None, None, None, None)
    assert E_INVALID != 0
    assert lib.mincob_optimize(None, None, None, None, None, None, None, None) == E_INVALID
    assert lib.mincob_max_rates(None, None, None, None) == E_INVALID
    assert lib.mincob_check_feasibility(None, None, None, 8, None) == E_INVALID
    assert lib.mincob_optimize_sharded(None, None, None, None, None, None, None, None) == E_INVALID
    assert lib.mincob_optimize_sharded_local(None, None, None, None, None, None, None, None) == E_INVALID
    p = C.c_void_p(); n = C.c_int64(0)
    assert lib.mincob_gathered_device(None, C.byref(p), C.byref(n)) == E_INVALID
    v = C.c_double(0.0)
    assert lib.mincob_measure_fp64_peak(None, C.byref(v)) == E_INVALID
    assert b"invalid" in lib.mincob_strerror(E_INVALID).lower()
