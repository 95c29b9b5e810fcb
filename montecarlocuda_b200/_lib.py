"""ctypes binding of libmcb200.so (include/mcb200.h).  Loads the in-tree CUDA library or fails
loudly: there is no Python, NumPy or CPU implementation of the pricing path behind this module."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_DIR = Path(__file__).resolve().parent / "lib"

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_OVERFLOW, ERR_UNSUPPORTED, ERR_ALIGNMENT, ERR_PEER_TIMEOUT = range(8)
OPT_TIMING, OPT_OVERLAP = 1, 2
PEER_WAIT, PEER_PUSH = 0, 1
F32, F64 = 0, 1
VANILLA, BASKET, CVA = 1, 2, 3
ACC_WORDS = 12
LANES = 5
MAX_ASSETS = 256
MAX_DATES = 1 << 20


class Mcb200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"mcb200 status {status}: {message}")
        self.status = status


class OptionT(C.Structure):
    """mcb200_option_t  <- OptionData (reference DP/MonteCarlo.h:32-38)"""
    _fields_ = [(k, C.c_double) for k in ("s", "k", "r", "v", "t")]


class BasketT(C.Structure):
    """mcb200_basket_t  <- MultiOptionData with a runtime width (reference DP/MonteCarlo.h:41-50)"""
    _fields_ = [("n", C.c_int)] + [(k, C.POINTER(C.c_double)) for k in ("s", "v", "p", "d", "w")] + \
               [(k, C.c_double) for k in ("k", "t", "r")]


class CvaT(C.Structure):
    """mcb200_cva_t  <- CVA (reference DP/MonteCarlo.h:57-65)"""
    _fields_ = [("def_int", C.c_double), ("lgd", C.c_double), ("option", OptionT),
                ("n_dates", C.c_int), ("grid_mode", C.c_int)]


class ResultT(C.Structure):
    _fields_ = [("n_paths", C.c_uint64)] + [(k, C.c_double) for k in (
        "sum", "sumsq", "mean", "expected", "confidence", "std_error", "kernel_ms")]


class JobT(C.Structure):
    """mcb200_job_t: one job of mcb200_price_batch"""
    _fields_ = [("workload", C.c_int), ("precision", C.c_int), ("params", C.c_void_p), ("n_paths", C.c_uint64), ("seed", C.c_uint64)]


class PlanT(C.Structure):
    _fields_ = [("workload", C.c_int), ("precision", C.c_int), ("total_paths", C.c_uint64),
                ("unit_paths", C.c_int), ("rounds", C.c_int), ("total_units", C.c_uint64),
                ("chunk_units", C.c_uint64), ("n_chunks", C.c_uint64), ("scale_exp_sum", C.c_int),
                ("scale_exp_sumsq", C.c_int), ("discount", C.c_double)]


PEER_HANDLE_BYTES = 64
_P = C.POINTER
_CTX = C.c_void_p
_SIGNATURES = {
    "mcb200_device_count": (C.c_int, []),
    "mcb200_create": (C.c_int, [_P(_CTX), C.c_int]),
    "mcb200_destroy": (C.c_int, [_CTX]),
    "mcb200_device": (C.c_int, [_CTX]),
    "mcb200_sm_count": (C.c_int, [_CTX]),
    "mcb200_strerror": (C.c_char_p, [C.c_int]),
    "mcb200_last_error": (C.c_char_p, [_CTX]),
    "mcb200_launch_count": (C.c_uint64, [_CTX]),
    "mcb200_peer_create": (C.c_int, [_CTX, C.c_int, C.c_int, _P(C.c_void_p), C.c_void_p]),
    "mcb200_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mcb200_peer_connect_local": (C.c_int, [_P(C.c_void_p), C.c_int]),
    "mcb200_peer_attach": (C.c_int, [_CTX, C.c_void_p]),
    "mcb200_peer_destroy": (C.c_int, [C.c_void_p]),
    "mcb200_peer_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "mcb200_peer_set_timeout_ms": (C.c_int, [C.c_void_p, C.c_double]),
    "mcb200_peer_pull": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcb200_set_option": (C.c_int, [_CTX, C.c_int, C.c_int]),
    "mcb200_get_option": (C.c_int, [_CTX, C.c_int]),
    "mcb200_set_basket_engine": (C.c_int, [C.c_int]),
    "mcb200_get_basket_engine": (C.c_int, []),
    "mcb200_vanilla": (C.c_int, [_CTX, C.c_int, _P(OptionT), C.c_uint64, C.c_uint64, _P(ResultT)]),
    "mcb200_basket": (C.c_int, [_CTX, C.c_int, _P(BasketT), C.c_uint64, C.c_uint64, _P(ResultT)]),
    "mcb200_cva": (C.c_int, [_CTX, C.c_int, _P(CvaT), C.c_uint64, C.c_uint64, _P(ResultT)]),
    "mcb200_vanilla_multi": (C.c_int, [_P(_CTX), C.c_int, C.c_int, _P(OptionT), C.c_uint64, C.c_uint64, _P(ResultT)]),
    "mcb200_basket_multi": (C.c_int, [_P(_CTX), C.c_int, C.c_int, _P(BasketT), C.c_uint64, C.c_uint64, _P(ResultT)]),
    "mcb200_cva_multi": (C.c_int, [_P(_CTX), C.c_int, C.c_int, _P(CvaT), C.c_uint64, C.c_uint64, _P(ResultT)]),
    "mcb200_price_batch": (C.c_int, [_CTX, C.c_int, _P(JobT), _P(ResultT), _P(C.c_int)]),
    "mcb200_plan_vanilla": (C.c_int, [C.c_int, _P(OptionT), C.c_uint64, _P(PlanT)]),
    "mcb200_plan_basket": (C.c_int, [C.c_int, _P(BasketT), C.c_uint64, _P(PlanT)]),
    "mcb200_plan_cva": (C.c_int, [C.c_int, _P(CvaT), C.c_uint64, _P(PlanT)]),
    "mcb200_shard_range": (C.c_int, [_P(PlanT), C.c_int, C.c_int, _P(C.c_uint64), _P(C.c_uint64)]),
    "mcb200_vanilla_launch": (C.c_int, [_CTX, _P(PlanT), _P(OptionT), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mcb200_basket_launch": (C.c_int, [_CTX, _P(PlanT), _P(BasketT), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mcb200_cva_launch": (C.c_int, [_CTX, _P(PlanT), _P(CvaT), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mcb200_finalize": (C.c_int, [_P(PlanT), _P(C.c_uint64), _P(ResultT)]),
    "mcb200_vanilla_paths": (C.c_int, [_CTX, C.c_int, _P(OptionT), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "mcb200_basket_paths": (C.c_int, [_CTX, C.c_int, _P(BasketT), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "mcb200_cva_paths": (C.c_int, [_CTX, C.c_int, _P(CvaT), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "mcb200_debug_philox": (C.c_int, [_CTX, C.c_uint64, C.c_void_p, _P(C.c_uint32), C.c_void_p]),
    "mcb200_debug_normals": (C.c_int, [_CTX, C.c_int, C.c_uint64, C.c_void_p, _P(C.c_uint32), C.c_void_p]),
    "mcb200_debug_math64": (C.c_int, [_CTX, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mcb200_debug_reduce": (C.c_int, [_CTX, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def library_path() -> Path:
    # MCB200_LIBRARY: an explicitly named build of the same library (A/B timing of kernel variants)
    override = os.environ.get("MCB200_LIBRARY")
    return Path(override) if override else LIB_DIR / "libmcb200.so"


def load() -> C.CDLL:
    """dlopen the in-tree libmcb200.so and bind every symbol of include/mcb200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise Mcb200Error(ERR_CUDA, f"{path} is missing: build it with `python -m montecarlocuda_b200.build` "
                                    "(nvcc, sm_100a); there is no CPU fallback")
    lib = C.CDLL(str(path))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(status: int, ctx=None):
    if status == OK:
        return
    lib = load()
    msg = lib.mcb200_strerror(status).decode()
    if ctx:
        detail = lib.mcb200_last_error(ctx).decode()
        if detail:
            msg += f" ({detail})"
    raise Mcb200Error(status, msg)
