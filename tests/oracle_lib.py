"""ctypes access to the CPU oracle (oracle/liboracle.so) and to the reference built into
oracle/_ref/.  Test infrastructure: imported by tests/, __graft_entry__.smoke() and bench.py's
CPU legs only."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
u32x4 = C.c_uint32 * 4
u32x2 = C.c_uint32 * 2
LANES = 5


class Oracle:
    def __init__(self, path=None):
        self.lib = lib = C.CDLL(str(path or ROOT / "oracle" / "liboracle.so"))
        d, f, u64, i, vp = C.c_double, C.c_float, C.c_uint64, C.c_int, C.c_void_p
        sig = {
            "orc_philox4x32_10": (None, [vp, vp, vp]),
            "orc_normals_f64": (None, [vp, vp]), "orc_normals_f32": (None, [vp, vp]),
            "orc_uniforms_f64": (None, [vp, vp]), "orc_uniforms_f32": (None, [vp, vp]),
            "orc_cnd_hastings_f64": (d, [d]), "orc_cnd_hastings_f32": (f, [f]),
            "orc_bs_call_hastings_f64": (d, [d] * 5), "orc_bs_call_hastings_f32": (f, [f] * 5),
            "orc_bs_call_exact": (d, [d] * 5),
            "orc_chol_f64": (None, [i, vp, vp]), "orc_chol_f32": (None, [i, vp, vp]),
            "orc_cva_grid_f64": (None, [d, i, vp, vp]), "orc_cva_grid_f32": (None, [f, i, vp, vp]),
            "orc_cva_closed_form": (d, [d] * 7 + [i, vp, i]),
            "orc_vanilla_payoffs_f64": (None, [d] * 5 + [u64] * 3 + [vp]),
            "orc_vanilla_payoffs_f32": (None, [f] * 5 + [u64] * 3 + [vp]),
            "orc_basket_payoffs_f64": (None, [i, vp, vp, vp, vp, vp, d, d, d, i, u64, u64, u64, vp]),
            "orc_basket_payoffs_f32": (None, [i, vp, vp, vp, vp, vp, f, f, f, i, u64, u64, u64, vp]),
            "orc_cva_path_values_f64": (None, [d] * 7 + [i, i, u64, u64, u64, vp]),
            "orc_cva_path_values_f32": (None, [f] * 7 + [i, i, u64, u64, u64, vp]),
            "orc_closing": (None, [d, d, u64, d, d, i, vp, vp]),
            "orc_chunk_rounds": (i, [u64]),
            "orc_lanes_add": (i, [d, i, vp]),
            "orc_lanes_to_double": (d, [vp, i]),
            "orc_chunk_reduce": (None, [vp, u64, i, i, i, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args

    # ---- generator ----
    def philox(self, ctr, key):
        out = u32x4()
        self.lib.orc_philox4x32_10(u32x4(*[int(x) for x in ctr]), u32x2(*[int(x) for x in key]), out)
        return np.array(out, dtype=np.uint32)

    def philox_many(self, counters, key):
        counters = np.asarray(counters, dtype=np.uint32).reshape(-1, 4)
        return np.stack([self.philox(c, key) for c in counters])

    def normals(self, words, precision="f64"):
        w = u32x4(*[int(x) for x in words])
        if precision == "f64":
            z = (C.c_double * 4)()
            self.lib.orc_normals_f64(w, z)
            return np.array(z, dtype=np.float64)
        z = (C.c_float * 6)()  # three Box-Muller pairs per block in single precision
        self.lib.orc_normals_f32(w, z)
        return np.array(z, dtype=np.float32)

    def uniforms(self, words, precision="f64"):
        w = u32x4(*[int(x) for x in words])
        if precision == "f64":
            z = (C.c_double * 4)()
            self.lib.orc_uniforms_f64(w, z)
            return np.array(z, dtype=np.float64)
        z = (C.c_float * 6)()
        self.lib.orc_uniforms_f32(w, z)
        return np.array(z, dtype=np.float32)

    # ---- closed forms / helpers ----
    def bs_call_hastings(self, s, k, r, v, t, precision="f64"):
        return (self.lib.orc_bs_call_hastings_f64 if precision == "f64" else self.lib.orc_bs_call_hastings_f32)(s, k, r, v, t)

    def bs_call_exact(self, s, k, r, v, t):
        return self.lib.orc_bs_call_exact(s, k, r, v, t)

    def cnd(self, d, precision="f64"):
        return (self.lib.orc_cnd_hastings_f64 if precision == "f64" else self.lib.orc_cnd_hastings_f32)(d)

    def chol(self, c, precision="f64"):
        dt = np.float64 if precision == "f64" else np.float32
        c = np.ascontiguousarray(c, dtype=dt)
        a = np.zeros_like(c)
        (self.lib.orc_chol_f64 if precision == "f64" else self.lib.orc_chol_f32)(c.shape[0], c.ctypes.data, a.ctypes.data)
        return a

    def cva_grid(self, T, n, precision="f64"):
        dt = np.float64 if precision == "f64" else np.float32
        tau = np.zeros(n, dtype=dt)
        keep = np.zeros(n, dtype=np.int32)
        (self.lib.orc_cva_grid_f64 if precision == "f64" else self.lib.orc_cva_grid_f32)(T, n, tau.ctypes.data, keep.ctypes.data)
        return tau, keep

    def cva_closed_form(self, s, k, r, v, T, lam, lgd, n, keep, lagged=False):
        keep = np.ascontiguousarray(keep, dtype=np.int32)
        return self.lib.orc_cva_closed_form(s, k, r, v, T, lam, lgd, n, keep.ctypes.data, int(lagged))

    # ---- per-path values ----
    def vanilla_payoffs(self, s, k, r, v, t, seed, first, n, precision="f64"):
        out = np.empty(n, dtype=np.float64 if precision == "f64" else np.float32)
        (self.lib.orc_vanilla_payoffs_f64 if precision == "f64" else self.lib.orc_vanilla_payoffs_f32)(
            s, k, r, v, t, seed, first, n, out.ctypes.data)
        return out

    def basket_payoffs(self, s, v, p, d, w, k, t, r, seed, first, n, precision="f64", ref_dp_host_bug=False):
        dt = np.float64 if precision == "f64" else np.float32
        arrs = [np.ascontiguousarray(a, dtype=dt) for a in (s, v, p, d, w)]
        out = np.empty(n, dtype=dt)
        (self.lib.orc_basket_payoffs_f64 if precision == "f64" else self.lib.orc_basket_payoffs_f32)(
            len(arrs[0]), *[a.ctypes.data for a in arrs], k, t, r, int(ref_dp_host_bug), seed, first, n, out.ctypes.data)
        return out

    def cva_path_values(self, s, k, r, v, t, lam, lgd, n_grid, seed, first, n, precision="f64", lagged_spot=False):
        out = np.empty(n, dtype=np.float64 if precision == "f64" else np.float32)
        (self.lib.orc_cva_path_values_f64 if precision == "f64" else self.lib.orc_cva_path_values_f32)(
            s, k, r, v, t, lam, lgd, n_grid, int(lagged_spot), seed, first, n, out.ctypes.data)
        return out

    def closing(self, total, total_sq, n, r, t, discount=True):
        e, c = C.c_double(), C.c_double()
        self.lib.orc_closing(total, total_sq, n, r, t, int(discount), C.byref(e), C.byref(c))
        return e.value, c.value

    # ---- order-free combine ----
    def chunk_rounds(self, total_units):
        return self.lib.orc_chunk_rounds(total_units)

    def lanes_add(self, value, scale_exp, lanes):
        return self.lib.orc_lanes_add(value, scale_exp, lanes.ctypes.data)

    def lanes_to_double(self, lanes, scale_exp):
        lanes = np.ascontiguousarray(lanes, dtype=np.uint64)
        return self.lib.orc_lanes_to_double(lanes.ctypes.data, scale_exp)

    def chunk_reduce(self, values, unit_paths, rounds, accumulate_in_float=False):
        v = np.ascontiguousarray(values, dtype=np.float64)
        s, s2 = C.c_double(), C.c_double()
        self.lib.orc_chunk_reduce(v.ctypes.data, len(v), unit_paths, rounds, int(accumulate_in_float), C.byref(s), C.byref(s2))
        return s.value, s2.value

    def accumulate(self, values, plan, first_chunk=0, n_chunks=None):
        """Accumulator block (uint64[12]) of chunks [first_chunk, first_chunk + n_chunks) of a job whose
        per-path values (job-global indexing, from path 0) are `values`: the CPU restatement of what
        one shard launch leaves in device memory."""
        acc = np.zeros(12, dtype=np.uint64)
        chunk_paths = int(plan.chunk_units) * int(plan.unit_paths)
        total = int(plan.total_paths)
        if n_chunks is None:
            n_chunks = int(plan.n_chunks) - first_chunk
        in_float = values.dtype == np.float32
        for c in range(first_chunk, first_chunk + n_chunks):
            lo = c * chunk_paths
            hi = min(lo + chunk_paths, total)
            s, s2 = self.chunk_reduce(values[lo:hi], int(plan.unit_paths), int(plan.rounds), in_float)
            bad = self.lanes_add(s, int(plan.scale_exp_sum), acc[0:5]) | self.lanes_add(s2, int(plan.scale_exp_sumsq), acc[5:10])
            acc[10] += np.uint64(hi - lo)
            acc[11] += np.uint64(bad)
        return acc


class Reference:
    """The unmodified reference compiled into oracle/_ref/libref_<dp|sp>_n<N>.so (oracle/Makefile)."""

    def __init__(self, precision="dp", n=3):
        path = ROOT / "oracle" / "_ref" / f"libref_{precision}_n{n}.so"
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference is mounted")
        self.lib = lib = C.CDLL(str(path))
        self.n = n
        real = C.c_double if precision == "dp" else C.c_float
        self.real = real
        self.np_real = np.float64 if precision == "dp" else np.float32

        class OptionData(C.Structure):
            _fields_ = [(k, real) for k in "skrvt"]

        class MultiOptionData(C.Structure):
            _fields_ = [("s", real * n), ("v", real * n), ("p", (real * n) * n), ("d", real * n), ("w", real * n),
                        ("k", real), ("t", real), ("r", real)]

        class OptionValue(C.Structure):
            _fields_ = [("Expected", real), ("Confidence", real)]

        class CVA(C.Structure):
            _fields_ = [("defInt", real), ("lgd", real), ("ns", C.c_int), ("option", OptionData), ("n", C.c_int)]

        self.OptionData, self.MultiOptionData, self.OptionValue, self.CVA = OptionData, MultiOptionData, OptionValue, CVA
        lib.host_bsCall.restype, lib.host_bsCall.argtypes = real, [OptionData]
        lib.host_vanillaOpt.restype, lib.host_vanillaOpt.argtypes = OptionValue, [OptionData, C.c_int]
        lib.host_basketOpt.restype, lib.host_basketOpt.argtypes = OptionValue, [C.POINTER(MultiOptionData), C.c_int]
        lib.host_cvaEquityOption.restype, lib.host_cvaEquityOption.argtypes = OptionValue, [C.POINTER(CVA), C.c_int]
        lib.Chol.restype, lib.Chol.argtypes = None, [C.c_void_p, C.c_void_p]
        lib.dev_vanillaOpt.restype, lib.dev_vanillaOpt.argtypes = OptionValue, [C.POINTER(OptionData), C.c_int, C.c_int, C.c_int]
        lib.dev_basketOpt.restype, lib.dev_basketOpt.argtypes = OptionValue, [C.POINTER(MultiOptionData), C.c_int, C.c_int, C.c_int]
        lib.dev_cvaEquityOption.restype, lib.dev_cvaEquityOption.argtypes = OptionValue, [C.POINTER(CVA), C.c_int, C.c_int, C.c_int]
        lib.ref_set_time.restype, lib.ref_set_time.argtypes = None, [C.c_long, C.c_int]

    def set_seed(self, value, auto_advance=False):
        self.lib.ref_set_time(int(value), int(auto_advance))

    def option(self, s, k, r, v, t):
        return self.OptionData(s, k, r, v, t)

    def multi(self, s, v, p, d, w, k, t, r):
        n = self.n
        m = self.MultiOptionData()
        for i in range(n):
            m.s[i], m.v[i], m.d[i], m.w[i] = s[i], v[i], d[i], w[i]
            for j in range(n):
                m.p[i][j] = p[i][j]
        m.k, m.t, m.r = k, t, r
        return m

    def cva(self, lam, lgd, opt, n_dates):
        return self.CVA(lam, lgd, 1, opt, n_dates)

    def chol(self, c):
        n = self.n
        c = np.ascontiguousarray(c, dtype=self.np_real)
        a = np.zeros((n, n), dtype=self.np_real)
        self.lib.Chol(c.ctypes.data, a.ctypes.data)
        return a
