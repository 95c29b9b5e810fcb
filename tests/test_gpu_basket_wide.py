"""Baskets wider than the widest register template (64 < n <= 256): the generic route of kernels_basket.cu
(normals in a local-memory array, column sweep over row blocks of 16 accumulators, factor from device memory).

The reference's N is a compile-time macro with no upper bound (double_precision/MonteCarlo.h:16); its kernel keeps
g[], bt[], s[] in local memory for every N (DP/MonteCarloKernel.cu:74-101).  Bars: per-path payoffs against the oracle
as for the register templates (1e-10 fp64, 9e-3 fp32); the accumulator bit for bit against the oracle's restatement
applied to the kernel's own per-path values; and -- the route keeps the templates' summation order -- a NARROW basket
forced through it (MCB200_BASKET_WIDE=1, a separate process) gives the register templates' values bit for bit.
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import montecarlocuda_b200 as m
from montecarlocuda_b200 import _lib as _lib_mod
from test_gpu_parity import make_basket

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n_assets", [65, 100, 256])
def test_wide_basket_paths_match_oracle(engine, oracle, prec, n_assets):
    opt = make_basket(oracle, n_assets, prec)
    n, first, seed = 1024, 1000, 99
    got = engine.basket_paths(opt, first, n, prec, seed).astype(np.float64)
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, seed, first, n, prec).astype(np.float64)
    tol = 1e-10 if prec == "f64" else 3e-5 * 300
    assert np.max(np.abs(got - want)) < tol
    assert got.max() > 0 and (got == 0).any()          # both sides of the strike were seen


def test_wide_basket_full_matrix_factor(engine, oracle):
    # a caller whose p is not triangular (the reference multiplies the full matrix, MonteCarloKernel.cu:79-84)
    rng = np.random.default_rng(6)
    n = 70
    p = rng.uniform(-0.2, 0.2, (n, n))
    opt = m.MultiOptionData(list(rng.uniform(80, 120, n)), list(rng.uniform(0.1, 0.3, n)), p, list(rng.uniform(-0.02, 0.02, n)),
                            list(rng.uniform(0.0, 2.0 / n, n)), 98.0, 0.75, 0.03)
    got = engine.basket_paths(opt, 0, 512, "f64", 5)
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, 5, 0, 512, "f64")
    assert np.max(np.abs(got - want)) < 1e-10


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_wide_basket_price_sums_its_paths_exactly(engine, oracle, prec):
    opt = make_basket(oracle, 100, prec)
    n, seed = 20_000, 7
    p = m.plan("basket", opt, n, prec)
    r = engine.basket(opt, n, prec, seed)
    assert r == m.finalize(p, oracle.accumulate(engine.basket_paths(opt, 0, n, prec, seed), p))
    # 100 equicorrelated assets: between the 64-asset value (8.13) and the infinitely diversified limit
    big = engine.basket(opt, 1 << 20, prec, seed)
    assert big.n_paths == 1 << 20 and 7.6 < big.Expected < 8.4 and big.std_error < 0.02


def test_wider_than_the_wide_route_is_refused(engine, oracle):
    opt = make_basket(oracle, 257)
    with pytest.raises(_lib_mod.Mcb200Error) as err:
        engine.basket(opt, 1024, "f64")
    assert err.value.status == _lib_mod.ERR_UNSUPPORTED


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_wide_dropin_library(engine, oracle, precision):
    """dev_basketOpt of a reference build with `#define N 100` (libmcb200_{dp,sp}_n100.so): the struct the reference
    driver fills, by pointer, in; the extended API's price out."""
    import ctypes as C
    n = 100
    real = C.c_double if precision == "dp" else C.c_float
    lib = C.CDLL(str(ROOT / "montecarlocuda_b200" / "lib" / f"libmcb200_{precision}_n{n}.so"))

    class OptionValue(C.Structure):
        _fields_ = [("Expected", real), ("Confidence", real)]

    class MultiOptionData(C.Structure):
        _fields_ = [("s", real * n), ("v", real * n), ("p", (real * n) * n), ("d", real * n), ("w", real * n), ("k", real), ("t", real), ("r", real)]

    lib.dev_basketOpt.restype, lib.dev_basketOpt.argtypes = OptionValue, [C.POINTER(MultiOptionData), C.c_int, C.c_int, C.c_int]
    prec = {"dp": "f64", "sp": "f32"}[precision]
    opt = make_basket(oracle, n, prec)
    mo = MultiOptionData()
    for i in range(n):
        mo.s[i], mo.v[i], mo.d[i], mo.w[i] = opt.s[i], opt.v[i], opt.d[i], opt.w[i]
        for j in range(n):
            mo.p[i][j] = opt.p[i][j]
    mo.k, mo.t, mo.r = opt.k, opt.t, opt.r
    got = lib.dev_basketOpt(mo, 512, 128, 1 << 18)
    # the SP library receives float fields: give our side the same rounded parameters
    rounded = m.MultiOptionData([real(x).value for x in opt.s], [real(x).value for x in opt.v],
                                np.array([[real(x).value for x in row] for row in opt.p]), [real(x).value for x in opt.d],
                                [real(x).value for x in opt.w], real(opt.k).value, real(opt.t).value, real(opt.r).value)
    ours = engine.basket(rounded, 1 << 18, prec)
    assert got.Expected == real(ours.Expected).value and got.Confidence == real(ours.Confidence).value


_CHILD = """
import json, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import montecarlocuda_b200 as m
from oracle_lib import Oracle
from test_gpu_parity import make_basket
oracle = Oracle()
m.set_basket_engine(m.BASKET_FFMA)
out = {{}}
with m.Engine(0) as eng:
    for n_assets in (3, 10, 64):
        for prec in ("f32", "f64"):
            opt = make_basket(oracle, n_assets, prec)
            vals = eng.basket_paths(opt, 0, 2048, prec, 11)
            r = eng.basket(opt, 50_000, prec, 11)
            out[f"{{n_assets}}_{{prec}}"] = [vals.tobytes().hex(), r.Expected.hex(), r.Confidence.hex()]
print("WIDE " + json.dumps(out))
"""


def test_wide_route_gives_the_register_templates_bits(engine, oracle):
    """Same summation order (columns 0, 1, 2, ... into each exponent, assets in order into the payoff), same normals,
    same exponential: forced through the wide route, 3, 10 and 64 assets give the register templates' per-path values
    and prices bit for bit, in both precisions (fp32: the packed-FMA engine; the tensor-core engine rounds its mat-vec
    differently by design)."""
    env = dict(os.environ, MCB200_BASKET_WIDE="1")
    res = subprocess.run([sys.executable, "-c", _CHILD.format(root=str(ROOT))], env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in res.stdout.splitlines() if l.startswith("WIDE ")]
    assert line, res.stderr[-2000:]
    wide = json.loads(line[-1][5:])
    keep = m.get_basket_engine()
    m.set_basket_engine(m.BASKET_FFMA)
    try:
        for n_assets in (3, 10, 64):
            for prec in ("f32", "f64"):
                opt = make_basket(oracle, n_assets, prec)
                vals = engine.basket_paths(opt, 0, 2048, prec, 11)
                r = engine.basket(opt, 50_000, prec, 11)
                assert wide[f"{n_assets}_{prec}"] == [vals.tobytes().hex(), r.Expected.hex(), r.Confidence.hex()], (n_assets, prec)
    finally:
        m.set_basket_engine(keep)
