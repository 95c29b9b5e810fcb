"""Accuracy of the hand-built fp64 special functions (montecarlocuda_b200/csrc/device_math64.cuh).

CPU part: the SAME header compiled for the host (tests/hostmath.cpp, MUFU seeds emulated at 20 bits)
against libm / long-double references.  GPU part (-m gpu): the device build through
mcb200_debug_math64.  Bars, written here: exp (both argument conventions), sqrt, 1/x <= 2 ulp; cos/sin <= 3e-16 absolute;
-2 ln u <= 2.5e-16 absolute + 3 ulp relative.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def ulps(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def cases():
    rng = np.random.default_rng(42)
    k = rng.integers(0, 2 ** 52, 200_000, dtype=np.uint64)
    u = 2.0 - (k | np.uint64(0x3FF0000000000000)).view(np.float64)
    u = np.concatenate([u, [1.0, 1.0 - 2.0 ** -52, 1.0 - 2.0 ** -51, 2.0 ** -52, 0.5, 0.5 - 2.0 ** -53, 0.999999, 0.99]])
    kk = np.concatenate([rng.integers(0, 2 ** 52, 200_000, dtype=np.uint64),
                         np.array([0, 1 << 49, 1 << 50, (1 << 49) - 1, (1 << 52) - 1, 1 << 51, 3 << 49, 5 << 49, 7 << 49], dtype=np.uint64)])
    return {
        "exp": np.concatenate([rng.uniform(-40, 40, 200_000), rng.uniform(-1, 1, 50_000), rng.uniform(-700, 700, 50_000)]),
        "sqrt": np.concatenate([rng.uniform(0, 80, 200_000), 10.0 ** rng.uniform(-300, 300, 50_000)]),
        "rcp": np.concatenate([rng.uniform(1, 10, 200_000), 10.0 ** rng.uniform(-100, 100, 50_000)]),
        "u": u, "k": kk,
    }


def check(run):
    """run(fn_id, array) -> (n, 2) array of results"""
    c = cases()
    assert ulps(run(3, c["exp"])[:, 0], np.exp(c["exp"])).max() <= 2
    edge = run(3, np.array([-700.0, 700.0, 0.0]))[:, 0]                          # the validated domain is |x| <= 700
    assert ulps(edge, np.exp(np.array([-700.0, 700.0, 0.0]))).max() <= 2
    assert np.isnan(run(3, np.array([np.nan]))[0, 0])
    # the same exponential taking its argument in table units: 2^(y/256), |y| <= 700 * 256 / ln 2
    y = np.concatenate([c["exp"] * 369.3299304675746, np.array([0.0, 0.5, -0.5, 1.5, 255.5, 256.0, -258000.0, 258000.0]),
                        np.arange(-600.0, 600.0, 0.25)])
    assert ulps(run(7, y)[:, 0], np.exp2(y.astype(np.longdouble) / 256).astype(np.float64)).max() <= 2
    assert ulps(run(8, y)[:, 0], np.exp2(y.astype(np.longdouble) / 256).astype(np.float64)).max() <= 2      # table entry last
    assert ulps(run(1, c["sqrt"])[:, 0], np.sqrt(c["sqrt"])).max() <= 2
    assert ulps(run(6, c["sqrt"])[:, 0], np.sqrt(c["sqrt"])).max() <= 2      # short iteration (basket, CVA kernels)
    assert ulps(run(2, c["rcp"])[:, 0], 1.0 / c["rcp"]).max() <= 2
    got = run(0, c["u"])[:, 0]
    want = (-2 * np.log(c["u"].astype(np.longdouble))).astype(np.float64)
    err = np.abs(got - want)
    assert np.all(err <= 2.5e-16 + 3 * 2.0 ** -52 * want)       # table + polynomial: ~1 ulp of 1, plus 3 ulp relative
    assert got.min() >= -1e-16          # callers take |.|: never meaningfully negative
    assert got[c["u"] == 1.0][0] == 1e-300     # never exactly 0: the square root after it has no zero guard
    k = c["k"]
    out = run(4, k.view(np.float64))
    ang = 2 * np.longdouble("3.14159265358979323846264338327950288") * (k.astype(np.longdouble) * np.longdouble(2.0) ** -52)
    assert np.abs(out[:, 0] - np.cos(ang).astype(np.float64)).max() <= 3e-16
    assert np.abs(out[:, 1] - np.sin(ang).astype(np.float64)).max() <= 3e-16
    k20 = np.concatenate([np.arange(0, 1 << 20, 37, dtype=np.uint64), np.array([0, 1, 1023, 1024, (1 << 20) - 1, 1 << 19, 3 << 18], dtype=np.uint64)])
    out = run(5, k20.view(np.float64))
    ang = 2 * np.longdouble("3.14159265358979323846264338327950288") * (k20.astype(np.longdouble) * np.longdouble(2.0) ** -20)
    assert np.abs(out[:, 0] - np.cos(ang).astype(np.float64)).max() <= 3e-16      # two-level table: four rounded entries
    assert np.abs(out[:, 1] - np.sin(ang).astype(np.float64)).max() <= 3e-16


@pytest.fixture(scope="module")
def hostmath(tmp_path_factory):
    so = tmp_path_factory.mktemp("hostmath") / "libhostmath.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(ROOT / "tests" / "hostmath.cpp")], check=True)
    return C.CDLL(str(so))


def test_host_build_of_device_math(hostmath):
    names = {0: "hm_neg2log", 1: "hm_sqrt", 2: "hm_rcp", 3: "hm_exp", 6: "hm_sqrt_short", 7: "hm_exp_units", 8: "hm_exp_units_late"}

    def run(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros((len(x), 2))
        if fn == 5:
            k = np.ascontiguousarray(x.view(np.uint64) & np.uint64(0xFFFFFFFF), dtype=np.uint64).astype(np.uint32)
            cs, sn = np.empty(len(x)), np.empty(len(x))
            hostmath.hm_sincos20(C.c_void_p(k.ctypes.data), C.c_void_p(cs.ctypes.data), C.c_void_p(sn.ctypes.data), C.c_long(len(x)))
            out[:, 0], out[:, 1] = cs, sn
        elif fn == 4:
            k = x.view(np.uint64)
            hi = (k >> np.uint64(32)).astype(np.uint32)
            lo = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            cs, sn = np.empty(len(x)), np.empty(len(x))
            hostmath.hm_sincos(C.c_void_p(hi.ctypes.data), C.c_void_p(lo.ctypes.data), C.c_void_p(cs.ctypes.data),
                               C.c_void_p(sn.ctypes.data), C.c_long(len(x)))
            out[:, 0], out[:, 1] = cs, sn
        else:
            res = np.empty(len(x))
            getattr(hostmath, names[fn])(C.c_void_p(x.ctypes.data), C.c_void_p(res.ctypes.data), C.c_long(len(x)))
            out[:, 0] = res
        return out

    check(run)


def test_scaled_log_is_the_log_with_the_scale_folded_in(hostmath):
    """scaled_log_unit(u, k, k ln2) = k ln(u): with k = -2 it IS neg2log_unit, bit for bit; with k = -2 b^2 (the
    diffusion scale folded under the Box-Muller square root, csrc/device_math.cuh polar_from_words) it stays within a
    few ulp of the exactly scaled value."""
    rng = np.random.default_rng(3)
    u = np.concatenate([2.0 - (1.0 + rng.integers(0, 1 << 44, 200_000) * 2.0 ** -44), [1.0, 2.0 ** -44, 0.5, 0.999999999999]])
    u = np.ascontiguousarray(u, dtype=np.float64)

    def scaled(k):
        out = np.empty(len(u))
        hostmath.hm_scaled_log(C.c_void_p(u.ctypes.data), C.c_double(k), C.c_double(k * 0.69314718055994530942),
                               C.c_void_p(out.ctypes.data), C.c_long(len(u)))
        return out

    ref = np.empty(len(u))
    hostmath.hm_neg2log(C.c_void_p(u.ctypes.data), C.c_void_p(ref.ctypes.data), C.c_long(len(u)))
    assert np.array_equal(scaled(-2.0), ref)
    for b in (0.2, 0.2 * 1.4426950408889634, 0.0282842712474619, 1.7):
        k = -2.0 * b * b
        want = (k * np.log(u.astype(np.longdouble))).astype(np.float64)
        err = np.abs(scaled(k) - want)
        # absolute part: neg2log's 1.25e-16 per unit of |k|, plus the rounding of the constant k ln2 (<= 2^-53 |k ln2|
        # per unit of exponent; it matters next to u = 1, where e ln2 and -ln c cancel)
        assert np.all(err <= abs(k) * 2.5e-16 + 4 * 2.0 ** -52 * np.abs(want) + 1e-299), (b, err.max())
    assert np.all(scaled(0.0) == 1e-300)       # zero volatility: the radius collapses to the guard value


@pytest.mark.gpu
def test_device_math_on_gpu(engine):
    check(lambda fn, x: engine.math64(fn, x))
