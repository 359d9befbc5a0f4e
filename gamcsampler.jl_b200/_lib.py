"""ctypes binding of libgamc_b200.so (include/gamc_b200.h).  There is NO CPU fallback: if the CUDA
library is missing or cannot be loaded this module raises, and every call surfaces the C status code
as a Python exception that mirrors the reference's error behaviour (AssertionError for bad sampler
parameters, PosDefError for LAPACK's PosDefException, ...)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GAMC_LIB", os.path.join(_HERE, "libgamc_b200.so"))   # GAMC_LIB: debug builds only

GAMC_OK, GAMC_INVALID_ARG, GAMC_NONFINITE_INIT, GAMC_NOT_POSDEF, GAMC_CUDA_ERROR, GAMC_NCCL_ERROR, \
    GAMC_UNSUPPORTED_SHAPE, GAMC_NOT_READY = range(8)

# kernel families (gamc_kernel_family)
K_LOGLIK, K_LOGLIK_GRAD, K_METRIC, K_REDUCE, K_LINALG, K_ALLREDUCE, K_LAND, K_SMMALA_PRE, K_SMMALA_POST, K_AM_PRE, K_AM_POST, \
    K_HANDOVER = range(12)
KERNEL_FAMILIES = ("loglik", "loglik_grad", "metric", "reduce", "linalg", "allreduce", "land", "smmala_pre", "smmala_post",
                   "am_pre", "am_post", "handover")

# array ids (gamc_array_id)
ARR_VALUE, ARR_GRADLOGTARGET, ARR_TENSORLOGTARGET, ARR_LASTMEAN, ARR_SECONDLASTMEAN, ARR_OLDINVTENSOR, \
    ARR_CHOLINVTENSOR, ARR_OLDFIRSTTERM, ARR_PROPOSED_VALUE, ARR_GRADLOGLIKELIHOOD, ARR_GRADLOGPRIOR, \
    ARR_TENSORLOGLIKELIHOOD, ARR_TENSORLOGPRIOR = range(13)


class GamcError(RuntimeError):
    status = None


class PosDefError(GamcError):
    """Julia's PosDefException / SingularException (`chol`, `inv`, MvNormal constructor)."""


class NonFiniteInitError(AssertionError):
    """`@assert isfinite(pstate.logtarget)` & co, src/samplers/GAMC.jl:168-170."""


class CudaError(GamcError):
    pass


class NcclError(GamcError):
    pass


class UnsupportedShape(GamcError):
    pass


class NotReady(GamcError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("driftstep", C.c_double), ("minorscale", C.c_double), ("c", C.c_double),
        ("t0", C.c_int32), ("schedule", C.c_int32), ("sched_params", C.c_double * 3),
        ("total_tuner_kind", C.c_int32), ("targetrate", C.c_double), ("score_k", C.c_double),
        ("period", C.c_int32), ("verbose_total", C.c_int32), ("verbose_smmala", C.c_int32), ("verbose_am", C.c_int32),
        ("nsteps", C.c_int64), ("burnin", C.c_int64), ("am_mode", C.c_int32), ("sampler_kind", C.c_int32),
        ("reuse_tensor", C.c_int32), ("reserved_", C.c_int32),
    ]


class StateScalars(C.Structure):
    _fields_ = [
        ("count", C.c_int64), ("updatetensorcount", C.c_int64), ("step", C.c_double), ("sqrttunestep", C.c_double),
        ("total_accepted", C.c_int64), ("total_proposed", C.c_int64), ("total_totproposed", C.c_int64),
        ("total_rate", C.c_double),
        ("smmala_accepted", C.c_int64), ("smmala_proposed", C.c_int64), ("smmala_totproposed", C.c_int64),
        ("am_accepted", C.c_int64), ("am_proposed", C.c_int64), ("am_totproposed", C.c_int64),
        ("smmalafrequency", C.c_double), ("amfrequency", C.c_double),
        ("logtarget", C.c_double), ("ratio", C.c_double),
        ("presentupdatetensor", C.c_int32), ("pastupdatetensor", C.c_int32),
        ("error_flag", C.c_int32), ("error_pivot", C.c_int32), ("error_iteration", C.c_int64),
        ("chain_saved", C.c_int64),
    ]


_DP = C.POINTER(C.c_double)
_U8P = C.POINTER(C.c_uint8)
_H = C.c_void_p

# every symbol include/gamc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gamc_create": (C.c_int, [C.POINTER(_H), C.c_int32]),
    "gamc_destroy": (C.c_int, [_H]),
    "gamc_last_error_string": (C.c_char_p, [_H]),
    "gamc_set_stream": (C.c_int, [_H, C.c_void_p]),
    "gamc_sync": (C.c_int, [_H]),
    "gamc_version": (C.c_char_p, []),
    "gamc_target_logistic": (C.c_int, [_H, _DP, C.c_int64, _DP, C.c_int64, C.c_int32, C.c_double]),
    "gamc_target_logistic_device": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_double]),
    "gamc_target_logistic_synth": (C.c_int, [_H, C.c_uint64, C.c_int64, C.c_int64, C.c_int32, _DP, C.c_double]),
    "gamc_target_callback": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p]),
    "gamc_target_read_rows": (C.c_int, [_H, C.c_int64, C.c_int64, _DP, _DP]),
    "gamc_eval_logtarget": (C.c_int, [_H, _DP, _DP]),
    "gamc_eval_upto_tensor": (C.c_int, [_H, _DP, _DP, _DP, _DP]),
    "gamc_eval_partials": (C.c_int, [_H, _DP, _DP, _DP, _DP]),
    "gamc_comm_unique_id": (C.c_int, [_U8P]),
    "gamc_comm_init": (C.c_int, [_H, _U8P, C.c_int32, C.c_int32]),
    "gamc_comm_peer_path": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "gamc_sampler_init": (C.c_int, [_H, C.POINTER(Config), _DP]),
    "gamc_tape_set": (C.c_int, [_H, C.c_int64, _DP, _DP, _DP, _DP]),
    "gamc_tape_generate": (C.c_int, [_H, C.c_int64, C.c_uint64]),
    "gamc_run": (C.c_int, [_H, C.c_int64]),
    "gamc_chain_get": (C.c_int, [_H, C.c_int64, C.c_int64, _DP, _U8P]),
    "gamc_iterate": (C.c_int, [_H, C.c_int64, _DP, _DP, _DP, _DP, _DP, _U8P, C.POINTER(C.c_int64)]),
    "gamc_peek_kinds": (C.c_int, [_H, C.c_int64, _U8P]),
    "gamc_get_state": (C.c_int, [_H, C.POINTER(StateScalars)]),
    "gamc_get_array": (C.c_int, [_H, C.c_int32, _DP]),
    "gamc_batched_target_mvt": (C.c_int, [_H, C.c_int32, C.c_double, _DP, C.c_double]),
    "gamc_batched_target_rv": (C.c_int, [_H, _DP, _DP, _DP, C.c_int64, C.c_int32]),
    "gamc_batched_init": (C.c_int, [_H, C.POINTER(Config), C.c_int32, _DP, C.c_int32, C.c_double]),
    "gamc_batched_run": (C.c_int, [_H, C.c_int64, C.c_uint64, _DP, _DP, _DP, _DP]),
    "gamc_batched_set_chain_offset": (C.c_int, [_H, C.c_int64]),
    "gamc_batched_chain_get": (C.c_int, [_H, C.c_int32, C.c_int32, _DP, _U8P]),
    "gamc_batched_get_state": (C.c_int, [_H, C.c_int32, C.POINTER(StateScalars), _DP]),
    "gamc_profile_enable": (C.c_int, [_H, C.c_int32]),
    "gamc_profile_get": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_int64), _DP]),
    "gamc_profile_reset": (C.c_int, [_H]),
    "gamc_launch_count": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "gamc_target_device_callback": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p]),
    "gamc_comm_set_timeout": (C.c_int, [_H, C.c_double]),
    "gamc_set_metric_route": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "gamc_get_metric_route": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gamc_set_schedule_callback": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "gamc_set_state": (C.c_int, [_H, C.POINTER(StateScalars), _DP, _DP, _DP, _DP]),
    "gamc_get_monitor": (C.c_int, [_H, _DP, _DP]),
    "gamc_chain_get_logtarget": (C.c_int, [_H, C.c_int64, C.c_int64, _DP]),
    "gamc_tune_log_get": (C.c_int, [_H, _DP, C.c_int64, C.POINTER(C.c_int64)]),
    "gamc_probe_fp64_peak": (C.c_int, [_H, _DP, _DP]),
    "gamc_bigd_target_mvt": (C.c_int, [_H, C.c_int32, C.c_double, _DP, _DP]),
    "gamc_bigd_set_chain_offset": (C.c_int, [_H, C.c_int64]),
    "gamc_bigd_init": (C.c_int, [_H, C.POINTER(Config), C.c_int32, _DP, C.c_int32, C.c_double, C.c_int32]),
    "gamc_bigd_run": (C.c_int, [_H, C.c_int64, C.c_uint64, _DP, _DP, _DP, _DP]),
    "gamc_bigd_chain_get": (C.c_int, [_H, C.c_int32, C.c_int32, _DP, _U8P]),
    "gamc_bigd_get_state": (C.c_int, [_H, C.c_int32, C.POINTER(StateScalars), _DP]),
    "gamc_tape_get": (C.c_int, [_H, C.c_int64, C.c_int64, _DP, _DP, _DP, _DP]),
    "gamc_batched_target_logistic": (C.c_int, [_H, _DP, C.c_int64, _DP, C.c_int64, C.c_int32, C.c_double]),
}

_lib = None


# gamc_target_fn of include/gamc_b200.h
TARGET_FN = C.CFUNCTYPE(C.c_int32, C.POINTER(C.c_double), C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                        C.POINTER(C.c_double), C.c_void_p)


# gamc_target_enqueue_fn / gamc_schedule_fn of include/gamc_b200.h
TARGET_ENQUEUE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)
SCHEDULE_FN = C.CFUNCTYPE(C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_void_p)


def load():
    """Load libgamc_b200.so (once).  Raises ImportError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built.  Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C gamcsampler.jl_b200/csrc`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    """double* of a contiguous float64 array (or NULL for None)."""
    if a is None:
        return None
    assert a.dtype == np.float64
    return a.ctypes.data_as(_DP)


def u8ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint8
    return a.ctypes.data_as(_U8P)


def check(status, handle=None):
    if status == GAMC_OK:
        return
    msg = ""
    if handle is not None and _lib is not None:
        raw = _lib.gamc_last_error_string(handle)
        msg = raw.decode() if raw else ""
    if status == GAMC_INVALID_ARG:
        raise AssertionError(msg or "invalid argument")
    if status == GAMC_NONFINITE_INIT:
        raise NonFiniteInitError(msg)
    exc = {GAMC_NOT_POSDEF: PosDefError, GAMC_CUDA_ERROR: CudaError, GAMC_NCCL_ERROR: NcclError,
           GAMC_UNSUPPORTED_SHAPE: UnsupportedShape, GAMC_NOT_READY: NotReady}.get(status, GamcError)
    e = exc(msg or f"gamc status {status}")
    e.status = status
    raise e
