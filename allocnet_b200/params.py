"""`mincob_params` — plain-C parameter block passed through the FFI (include/mincob.h).

Penalty defaults: upstream GCOPTER config values as recalled in SURVEY.md
Appendix B.1 (`SmoothingEps` survives in `config/planner.yaml:15`; `MaxVelBox`,
`MaxAccBox` in `config/planner.yaml:17,19`; `max_jerk` in
`network/utils/params.yaml:4`).  L-BFGS defaults: upstream `optimize()` settings
(`past=3, min_step=1e-32, g_epsilon=0, delta=1e-5`) on top of the
`lbfgs_parameter_t` defaults of `gcopter/lbfgs.hpp:15-129`; `mem_size` keeps the
header default 8 and `max_iterations` is the safety cap.
"""
from __future__ import annotations

import ctypes as C

# lbfgs return codes, gcopter/lbfgs.hpp:135-184
LBFGS_CONVERGENCE = 0
LBFGS_STOP = 1
LBFGS_CANCELED = 2
LBFGSERR_UNKNOWNERROR = -1024
(LBFGSERR_INVALID_N, LBFGSERR_INVALID_MEMSIZE, LBFGSERR_INVALID_GEPSILON, LBFGSERR_INVALID_TESTPERIOD,
 LBFGSERR_INVALID_DELTA, LBFGSERR_INVALID_MINSTEP, LBFGSERR_INVALID_MAXSTEP, LBFGSERR_INVALID_FDECCOEFF,
 LBFGSERR_INVALID_SCURVCOEFF, LBFGSERR_INVALID_MACHINEPREC, LBFGSERR_INVALID_MAXLINESEARCH,
 LBFGSERR_INVALID_FUNCVAL, LBFGSERR_MINIMUMSTEP, LBFGSERR_MAXIMUMSTEP, LBFGSERR_MAXIMUMLINESEARCH,
 LBFGSERR_MAXIMUMITERATION, LBFGSERR_WIDTHTOOSMALL, LBFGSERR_INVALIDPARAMETERS,
 LBFGSERR_INCREASEGRADIENT) = range(-1023, -1023 + 19)


# mincob_params.flags / .mapping (include/mincob.h)
FLAG_FREEZE_TIMES = 1     # durations are data (the fixed-time call of learning_planner.hpp:196), waypoints only
FLAG_PLANNER_ROWS = 2     # rows are [n, b] with n.p <= b (learning_planner.hpp:293-299) instead of n.p + d <= 0
MAP_AUTO, MAP_THROUGHPUT, MAP_LATENCY = 0, 1, 2


class MincobParams(C.Structure):
    _fields_ = [
        ("S", C.c_int32), ("kappa", C.c_int32),
        ("mu", C.c_double), ("w_pos", C.c_double), ("w_vel", C.c_double), ("w_acc", C.c_double),
        ("w_jerk", C.c_double), ("v_max", C.c_double), ("a_max", C.c_double), ("j_max", C.c_double),
        ("rho", C.c_double),
        ("mem_size", C.c_int32), ("past", C.c_int32), ("max_iterations", C.c_int32),
        ("max_linesearch", C.c_int32),
        ("g_epsilon", C.c_double), ("delta", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double),
        ("f_dec_coeff", C.c_double), ("s_curv_coeff", C.c_double), ("cautious_factor", C.c_double),
        ("machine_prec", C.c_double),
        ("flags", C.c_int32), ("mapping", C.c_int32),
    ]


def default_params(S: int = 3, **over) -> MincobParams:
    p = MincobParams(
        S=S, kappa=16, mu=1.0e-2, w_pos=1.0e4, w_vel=1.0e4, w_acc=1.0e4, w_jerk=1.0e4,
        v_max=4.0, a_max=6.0, j_max=12.0, rho=20.0,
        mem_size=8, past=3, max_iterations=1000, max_linesearch=64,
        g_epsilon=0.0, delta=1.0e-5, min_step=1.0e-32, max_step=1.0e20,
        f_dec_coeff=1.0e-4, s_curv_coeff=0.9, cautious_factor=1.0e-6, machine_prec=1.0e-16,
        flags=0, mapping=MAP_AUTO)
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(f"mincob_params has no field {k!r}")
        setattr(p, k, v)
    return p


def energy_only(p: MincobParams) -> MincobParams:
    """Config 2 of BASELINE.json: no corridor, no magnitude penalties (cost = E + rho*sum T)."""
    q = MincobParams.from_buffer_copy(p)
    q.w_pos = q.w_vel = q.w_acc = q.w_jerk = 0.0
    return q
