"""Host-side mirror of the reference interface, bound to the CUDA library through its C ABI.

Names follow the reference (`/root/reference/double_precision/MonteCarlo.h:32-65`,
`MonteCarloKernel.cu:483-532`): OptionData, MultiOptionData, CVA, OptionValue, dev_vanillaOpt,
dev_basketOpt, dev_cvaEquityOption.  `Engine` is the persistent context the reference lacks.
Every number comes from libmcb200.so; nothing here computes a price.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import F32, F64, Mcb200Error  # noqa: F401  (re-exported)

DEFAULT_SEED = 0x6D63623230300001  # same constant as csrc/dropin.cpp


def _prec(precision) -> int:
    if precision in (F64, "f64", "fp64", "double", "dp", np.float64):
        return F64
    if precision in (F32, "f32", "fp32", "float", "single", "sp", np.float32):
        return F32
    raise ValueError(f"unknown precision {precision!r}")


@dataclass
class OptionData:
    """European option (reference OptionData, DP/MonteCarlo.h:32-38)."""
    s: float
    k: float
    r: float
    v: float
    t: float

    def _c(self) -> _lib.OptionT:
        return _lib.OptionT(self.s, self.k, self.r, self.v, self.t)

    def _snapshot(self):
        return (float(self.s), float(self.k), float(self.r), float(self.v), float(self.t))


@dataclass
class MultiOptionData:
    """Basket of n underlyings (reference MultiOptionData, DP/MonteCarlo.h:41-50) with a runtime
    width.  `p` is the row-major n x n Cholesky factor of the correlation matrix, as the reference
    expects on entry to dev_basketOpt (DP/basketOpt.cu:96-99)."""
    s: Sequence[float]
    v: Sequence[float]
    p: Sequence[Sequence[float]]
    d: Sequence[float]
    w: Sequence[float]
    k: float
    t: float
    r: float

    @property
    def n(self) -> int:
        return len(self.s)

    def _c(self) -> _lib.BasketT:
        n = self.n
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in (self.s, self.v, self.p, self.d, self.w)]
        if arrs[2].shape != (n, n) or any(a.shape != (n,) for a in (arrs[0], arrs[1], arrs[3], arrs[4])):
            raise ValueError("MultiOptionData: s, v, d, w must have n entries and p must be n x n")
        ptr = [a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs]
        c = _lib.BasketT(n, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], self.k, self.t, self.r)
        # the STRUCT owns the buffers it points at (a ctypes pointer field does not keep its array alive): every call
        # returns an independent struct, valid for as long as the caller holds it
        c._buffers = arrs
        return c

    def _snapshot(self):
        """The job's parameters by value (cache keys: the dataclass is mutable)."""
        return (tuple(np.asarray(a, dtype=np.float64).ravel().tolist() for a in (self.s, self.v, self.p, self.d, self.w)),
                float(self.k), float(self.t), float(self.r))


@dataclass
class CVA:
    """CVA of one call (reference CVA, DP/MonteCarlo.h:57-65): flat default intensity, loss given
    default, the option, n exposure dates.  grid_mode 0 = the reference's time grid."""
    defInt: float
    lgd: float
    option: OptionData
    n: int
    ns: int = 1
    grid_mode: int = 0

    def _c(self) -> _lib.CvaT:
        return _lib.CvaT(self.defInt, self.lgd, self.option._c(), int(self.n), int(self.grid_mode))

    def _snapshot(self):
        return (float(self.defInt), float(self.lgd), self.option._snapshot(), int(self.n), int(self.grid_mode))


@dataclass
class OptionValue:
    """Reference OptionValue (DP/MonteCarlo.h:52-55) plus what the extended API reports."""
    Expected: float
    Confidence: float
    std_error: float = 0.0
    n_paths: int = 0
    mean: float = 0.0
    sum: float = 0.0
    sumsq: float = 0.0
    kernel_ms: float = 0.0

    @staticmethod
    def _from(r: _lib.ResultT) -> "OptionValue":
        return OptionValue(r.expected, r.confidence, r.std_error, int(r.n_paths), r.mean, r.sum, r.sumsq, r.kernel_ms)


BASKET_TENSOR, BASKET_FFMA = 0, 1


def set_basket_engine(engine: int) -> None:
    """Process-wide: which kernel prices wide fp32 baskets (32 < n <= 64): BASKET_TENSOR (tcgen05, default)
    or BASKET_FFMA (the packed-FMA column sweep).  Mirrors mcb200_set_basket_engine (include/mcb200.h)."""
    _lib.check(_lib.load().mcb200_set_basket_engine(int(engine)))


def get_basket_engine() -> int:
    return int(_lib.load().mcb200_get_basket_engine())


class Engine:
    """Persistent pricing context on one CUDA device (stream, accumulator, events)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        _lib.check(self._lib.mcb200_create(C.byref(self._ctx), int(device)))
        self.device = int(device)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.mcb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        return self._ctx

    @property
    def sm_count(self) -> int:
        return self._lib.mcb200_sm_count(self._ctx)

    @property
    def launch_count(self) -> int:
        return int(self._lib.mcb200_launch_count(self._ctx))

    def set_timing(self, on: bool = True) -> None:
        """Report OptionValue.kernel_ms (CUDA events around the kernels of the blocking calls; off by default: two
        more stream operations per call)."""
        _lib.check(self._lib.mcb200_set_option(self._ctx, _lib.OPT_TIMING, int(bool(on))), self._ctx)

    def set_overlap(self, on: bool = True) -> None:
        """Programmatic dependent launch: consecutive launches of a stream overlap tail and start."""
        _lib.check(self._lib.mcb200_set_option(self._ctx, _lib.OPT_OVERLAP, int(bool(on))), self._ctx)

    # ---- one-call pricing (host structs in, host result out) ----
    def vanilla(self, opt: OptionData, n_paths: int, precision=F64, seed: int = DEFAULT_SEED) -> OptionValue:
        c, r = opt._c(), _lib.ResultT()
        _lib.check(self._lib.mcb200_vanilla(self._ctx, _prec(precision), C.byref(c), n_paths, seed, C.byref(r)), self._ctx)
        return OptionValue._from(r)

    def basket(self, opt: MultiOptionData, n_paths: int, precision=F64, seed: int = DEFAULT_SEED) -> OptionValue:
        c, r = opt._c(), _lib.ResultT()
        _lib.check(self._lib.mcb200_basket(self._ctx, _prec(precision), C.byref(c), n_paths, seed, C.byref(r)), self._ctx)
        return OptionValue._from(r)

    def cva(self, cva: CVA, n_paths: int, precision=F64, seed: int = DEFAULT_SEED) -> OptionValue:
        c, r = cva._c(), _lib.ResultT()
        _lib.check(self._lib.mcb200_cva(self._ctx, _prec(precision), C.byref(c), n_paths, seed, C.byref(r)), self._ctx)
        return OptionValue._from(r)

    def price_batch(self, jobs, seed: int = DEFAULT_SEED):
        """Price many jobs with one synchronisation.  jobs: iterable of (workload, params, n_paths, precision)
        with workload in {"vanilla", "basket", "cva"}.  Returns a list of OptionValue, in order."""
        jobs = list(jobs)
        codes = {"vanilla": _lib.VANILLA, "basket": _lib.BASKET, "cva": _lib.CVA}
        c_params = [params._c() for _, params, _, _ in jobs]          # the C structs (and their buffers) live until the call returns
        arr = (_lib.JobT * len(jobs))()
        for i, (workload, _, n_paths, precision) in enumerate(jobs):
            arr[i] = _lib.JobT(codes[workload], _prec(precision), C.cast(C.pointer(c_params[i]), C.c_void_p), n_paths, seed)
        out = (_lib.ResultT * len(jobs))()
        status = (C.c_int * len(jobs))()
        _lib.check(self._lib.mcb200_price_batch(self._ctx, len(jobs), arr, out, status), self._ctx)
        return [OptionValue._from(r) for r in out]

    # ---- per-path values (parity instrumentation) ----
    def _paths(self, fn, c_struct, precision, seed, first_path, n_paths) -> np.ndarray:
        p = _prec(precision)
        out = np.empty(n_paths, dtype=np.float64 if p == F64 else np.float32)
        _lib.check(fn(self._ctx, p, C.byref(c_struct), seed, first_path, n_paths, out.ctypes.data), self._ctx)
        return out

    def vanilla_paths(self, opt, first_path, n_paths, precision=F64, seed=DEFAULT_SEED):
        return self._paths(self._lib.mcb200_vanilla_paths, opt._c(), precision, seed, first_path, n_paths)

    def basket_paths(self, opt, first_path, n_paths, precision=F64, seed=DEFAULT_SEED):
        return self._paths(self._lib.mcb200_basket_paths, opt._c(), precision, seed, first_path, n_paths)

    def cva_paths(self, cva, first_path, n_paths, precision=F64, seed=DEFAULT_SEED):
        return self._paths(self._lib.mcb200_cva_paths, cva._c(), precision, seed, first_path, n_paths)

    def philox(self, counters: np.ndarray, key) -> np.ndarray:
        ctr = np.ascontiguousarray(counters, dtype=np.uint32).reshape(-1, 4)
        out = np.empty_like(ctr)
        k = (C.c_uint32 * 2)(int(key[0]), int(key[1]))
        _lib.check(self._lib.mcb200_debug_philox(self._ctx, len(ctr), ctr.ctypes.data, k, out.ctypes.data), self._ctx)
        return out

    def normals(self, counters: np.ndarray, key, precision=F64) -> np.ndarray:
        p = _prec(precision)
        ctr = np.ascontiguousarray(counters, dtype=np.uint32).reshape(-1, 4)
        out = np.empty((len(ctr), 4 if p == F64 else 6), dtype=np.float64 if p == F64 else np.float32)
        k = (C.c_uint32 * 2)(int(key[0]), int(key[1]))
        _lib.check(self._lib.mcb200_debug_normals(self._ctx, p, len(ctr), ctr.ctypes.data, k, out.ctypes.data), self._ctx)
        return out

    def math64(self, fn: int, x: np.ndarray) -> np.ndarray:
        """The kernels' fp64 special functions on an array (fn: 0 -2ln u, 1 sqrt, 2 1/x, 3 e^x, 4 cos/sin turn)."""
        v = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty((len(v), 2), dtype=np.float64)
        _lib.check(self._lib.mcb200_debug_math64(self._ctx, int(fn), len(v), v.ctypes.data, out.ctypes.data), self._ctx)
        return out

    def reduce_chunk(self, values: np.ndarray, unit_paths: int, rounds: int, accumulate_in_float: bool,
                     scale_exp_sum: int, scale_exp_sumsq: int) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        acc = np.zeros(_lib.ACC_WORDS, dtype=np.uint64)
        _lib.check(self._lib.mcb200_debug_reduce(self._ctx, v.ctypes.data, len(v), unit_paths, rounds,
                                                 int(accumulate_in_float), scale_exp_sum, scale_exp_sumsq,
                                                 acc.ctypes.data), self._ctx)
        return acc


# ---- planning / closing: pure host functions of the C ABI (no device needed) ----
def plan(workload: str, params, n_paths: int, precision=F64) -> _lib.PlanT:
    lib = _lib.load()
    p = _lib.PlanT()
    c = params._c()
    fn = {"vanilla": lib.mcb200_plan_vanilla, "basket": lib.mcb200_plan_basket, "cva": lib.mcb200_plan_cva}[workload]
    _lib.check(fn(_prec(precision), C.byref(c), n_paths, C.byref(p)))
    return p


def shard_range(p: _lib.PlanT, rank: int, world: int):
    lib = _lib.load()
    first, count = C.c_uint64(), C.c_uint64()
    _lib.check(lib.mcb200_shard_range(C.byref(p), rank, world, C.byref(first), C.byref(count)))
    return int(first.value), int(count.value)


def finalize(p: _lib.PlanT, acc: np.ndarray) -> OptionValue:
    lib = _lib.load()
    a = np.ascontiguousarray(acc, dtype=np.uint64)
    r = _lib.ResultT()
    _lib.check(lib.mcb200_finalize(C.byref(p), a.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(r)))
    return OptionValue._from(r)


# ---- the reference's three entry points (MonteCarloKernel.cu:483, :500, :517) ----
_default_engines: dict = {}


def _engine(device: int = 0) -> Engine:
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]


def _ref_paths(numBlocks: int, sims: int) -> int:
    if numBlocks <= 0 or sims <= 0:
        raise ValueError("numBlocks and sims must be positive")
    return numBlocks * (sims // numBlocks)  # the reference's integer arithmetic (:491, :508, :524)


def dev_vanillaOpt(opt: OptionData, numBlocks: int, numThreads: int, sims: int, precision=F64,
                   seed: int = DEFAULT_SEED, device: int = 0) -> OptionValue:
    return _engine(device).vanilla(opt, _ref_paths(numBlocks, sims), precision, seed)


def dev_basketOpt(option: MultiOptionData, numBlocks: int, numThreads: int, sims: int, precision=F64,
                  seed: int = DEFAULT_SEED, device: int = 0) -> OptionValue:
    return _engine(device).basket(option, _ref_paths(numBlocks, sims), precision, seed)


def dev_cvaEquityOption(cva: CVA, numBlocks: int, numThreads: int, sims: int, precision=F64,
                        seed: int = DEFAULT_SEED, device: int = 0) -> OptionValue:
    return _engine(device).cva(cva, _ref_paths(numBlocks, sims), precision, seed)
