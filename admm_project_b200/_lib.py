"""ctypes binding of libadmm_b200.so (include/admm_b200.h).  The product path: if the CUDA
library is missing or no B200 is present this module FAILS LOUDLY -- there is no CPU fallback
(the CPU restatement lives in oracle/ and is test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadmm_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOTPOSDEF, ERR_COMM, ERR_UNSUPPORTED = range(7)
LASSO, BASISPURSUIT, TOTALVARIATION, SVM_HINGE, SVM_01, HUBERFIT, LAD, PROX_NONNEG, PROX_BOX, MODEL = range(1, 11)
STOP_STANDARD, STOP_HNORM, STOP_BOTH = 0, 1, 2
RUNNING, CONVERGED_STD, CONVERGED_HNORM, MAXITERS, DIVERGED_RETURN, CONVERGED_DVAL = range(6)
XSOLVE_INVFACTOR, XSOLVE_SUBST = 0, 1

_dp = C.POINTER(C.c_double)


class EngineError(RuntimeError):
    """A non-zero status from the C-ABI; carries admm_b200_last_error() (the reference raises
    MATLAB error(...) strings)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class Options(C.Structure):
    _fields_ = [("rho", C.c_double), ("relax", C.c_double), ("abstol", C.c_double),
                ("reltol", C.c_double), ("convtol", C.c_double), ("hnormtol", C.c_double),
                ("maxiters", C.c_int64), ("domaxiters", C.c_int32), ("stopcond", C.c_int32),
                ("nodualerror", C.c_int32), ("convtest", C.c_int32), ("objevals", C.c_int32),
                ("history", C.c_int32), ("xsolve", C.c_int32), ("check_every", C.c_int32),
                ("fast", C.c_int32), ("fasttype", C.c_int32), ("restart", C.c_double), ("dvaltol", C.c_double),
                ("graph", C.c_int32), ("reserved", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("steps", C.c_int64), ("status", C.c_int32), ("reserved", C.c_int32),
                ("objopt", C.c_double), ("setup_ms", C.c_double), ("loop_ms", C.c_double),
                ("xopt", _dp), ("zopt", _dp), ("uopt", _dp),
                ("pnorm", _dp), ("dnorm", _dp), ("perr", _dp), ("derr", _dp),
                ("hnormsq", _dp), ("objevals", _dp),
                ("xvals", _dp), ("zvals", _dp), ("uvals", _dp),
                ("dvals", _dp), ("avals", _dp), ("restarted", _dp)]


class Info(C.Structure):
    _fields_ = [("generation", C.c_int64), ("zero_cols", C.c_int64), ("diag_ratio", C.c_double),
                ("xsolve_effective", C.c_int32), ("p2p_ready", C.c_int32), ("nranks", C.c_int32), ("rank", C.c_int32)]


# every symbol include/admm_b200.h declares: name -> (restype, argtypes)
_i64, _i32, _d, _vp, _int = C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_int
SYMBOLS = {
    "admm_b200_version": (_int, []),
    "admm_b200_last_error": (C.c_char_p, []),
    "admm_b200_default_options": (None, [C.POINTER(Options)]),
    "admm_b200_create": (_int, [_int, C.POINTER(_vp)]),
    "admm_b200_destroy": (_int, [_vp]),
    "admm_b200_set_stream": (_int, [_vp, _vp]),
    "admm_b200_synchronize": (_int, [_vp]),
    "admm_b200_setup_lasso": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _d, _i32]),
    "admm_b200_setup_lasso_sharded": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _d, _i32]),
    "admm_b200_setup_unwrapped": (_int, [_vp, _i32, _i64, _i64, _i64, _vp, _i64, _vp, _d]),
    "admm_b200_setup_basispursuit": (_int, [_vp, _i64, _i64, _vp, _i64, _vp]),
    "admm_b200_setup_totalvariation": (_int, [_vp, _i64, _vp, _d]),
    "admm_b200_setup_quadratic": (_int, [_vp, _i32, _i64, _vp, _i64, _vp, _d, _d, _vp, _vp]),
    "admm_b200_setup_model": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _d]),
    "admm_b200_get_unique_id": (_int, [_vp]),
    "admm_b200_comm_init": (_int, [_vp, _int, _int, _vp]),
    "admm_b200_comm_destroy": (_int, [_vp]),
    "admm_b200_comm_ipc_export": (_int, [_vp, _int, _int, _vp]),
    "admm_b200_comm_ipc_attach": (_int, [_vp, _vp]),
    "admm_b200_allreduce": (_int, [_vp, _vp, _i64]),
    "admm_b200_set_lambda": (_int, [_vp, _d]),
    "admm_b200_set_init": (_int, [_vp, _vp, _vp, _vp]),
    "admm_b200_solve": (_int, [_vp, C.POINTER(Options), C.POINTER(Result)]),
    "admm_b200_solve_lasso_batch": (_int, [_vp, C.POINTER(Options), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _vp, C.POINTER(C.c_double)]),
    "admm_b200_solve_unwrapped_batch": (_int, [_vp, C.POINTER(Options), _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _vp, _vp, _vp, _vp, C.POINTER(C.c_double)]),
    "admm_b200_get_dims": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "admm_b200_get_factor": (_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "admm_b200_dgemm": (_int, [_vp, _int, _int, _i64, _i64, _i64, _d, _vp, _i64, _vp, _i64, _d, _vp, _i64, _int]),
    "admm_b200_gram": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _d, _d, _vp, _i64]),
    "admm_b200_potrf": (_int, [_vp, _i64, _vp, _i64, _vp, _i64]),
    "admm_b200_factor_solve": (_int, [_vp, _vp, _vp, _i32]),
    "admm_b200_iterate_raw": (_int, [_vp, C.POINTER(Options), _int, _int]),
    "admm_b200_launch_count": (_i64, [_vp]),
    "admm_b200_graph_replays": (_i64, [_vp]),
    "admm_b200_get_setup_phases": (_int, [_vp, C.POINTER(C.c_double)]),
    "admm_b200_get_info": (_int, [_vp, C.POINTER(Info)]),
    "admm_b200_slicemaker": (_int, [_i64, _i64, C.POINTER(_i64)]),
}

_lib = None


def load():
    """dlopen libadmm_b200.so and bind every symbol.  Raises EngineError when the library has not
    been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(ERR_CUDA, "libadmm_b200.so is not built (%s missing); run "
                          "__graft_entry__.build().  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != OK:
        raise EngineError(status, load().admm_b200_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array, an int (raw device pointer) or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


def fvec(v, n=None, name="vector"):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    if n is not None and a.size != n:
        raise EngineError(ERR_INVALID, "%s has %d entries, expected %d" % (name, a.size, n))
    return a


def fmat(a):
    """Column-major float64 view/copy of a 2-D array (MATLAB layout)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise EngineError(ERR_INVALID, "expected a matrix")
    return a if a.flags.f_contiguous else np.asfortranarray(a)
