"""GPU parity against closed forms beyond the BASELINE parameter point: a moneyness / volatility /
maturity sweep of the European call in both precisions, baskets that collapse to a single asset
(one underlying; perfectly correlated identical underlyings), and CVA cases whose value is known
exactly (no loss given default, no default intensity).  Bars: 4 standard errors per point (3 SE with
a dozen comparisons would fail one run in four by chance) and exact zeros where the value is zero."""
import math

import numpy as np
import pytest

import montecarlocuda_b200 as m

pytestmark = pytest.mark.gpu


def black_scholes_call(s, k, r, v, t):
    d1 = (math.log(s / k) + (r + 0.5 * v * v) * t) / (v * math.sqrt(t))
    d2 = d1 - v * math.sqrt(t)
    cnd = lambda x: 0.5 * math.erfc(-x / math.sqrt(2.0))  # noqa: E731
    return s * cnd(d1) - k * math.exp(-r * t) * cnd(d2)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_vanilla_sweep_against_black_scholes(engine, prec):
    n = 1 << 24
    for k in (60.0, 90.0, 110.0, 150.0):
        for v, t in ((0.1, 0.25), (0.5, 2.0), (0.2, 1.0)):
            opt = m.OptionData(100.0, k, 0.03, v, t)
            r = engine.vanilla(opt, n, prec, seed=int(k) * 1000 + int(100 * v))
            exact = black_scholes_call(100.0, k, 0.03, v, t)
            # fp32: MUFU-accurate exponentials move a price by a few 1e-6 relative on top of the sampling error
            slack = 1e-9 if prec == "f64" else 2e-5 * max(exact, 1.0)   # 1e-9: a deep out-of-the-money call prices to exactly 0
            assert abs(r.Expected - exact) < 4 * r.std_error + slack, (k, v, t, r.Expected, exact, r.std_error)
            assert r.n_paths == n


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_one_asset_basket_is_a_vanilla_call(engine, prec):
    # n = 1 runs on the 3-wide template with two zero-weight padding assets
    b = m.MultiOptionData([100.0], [0.2], np.array([[1.0]]), [0.0], [1.0], 100.0, 1.0, 0.05)
    r = engine.basket(b, 1 << 24, prec)
    assert abs(r.Expected - black_scholes_call(100.0, 100.0, 0.05, 0.2, 1.0)) < 4 * r.std_error + 2e-4


@pytest.mark.parametrize("n_assets,prec", [(10, "f64"), (10, "f32"), (64, "f32"), (64, "f64")])
def test_perfectly_correlated_identical_assets_collapse_to_one(engine, n_assets, prec):
    """Correlation 1 between identical underlyings: the Cholesky factor has a single non-zero column, every asset
    follows the first normal and the basket is one asset -- closed form.  Exercises the column sweep (and, for the wide
    fp32 basket, the tensor-core mat-vec; for the wide fp64 basket the two-pass sweep) with a rank-1 factor."""
    factor = np.zeros((n_assets, n_assets))
    factor[:, 0] = 1.0
    b = m.MultiOptionData([100.0] * n_assets, [0.25] * n_assets, factor, [0.0] * n_assets, [1.0 / n_assets] * n_assets,
                          95.0, 0.5, 0.04)
    paths = 1 << 22 if n_assets == 64 else 1 << 24
    r = engine.basket(b, paths, prec)
    exact = black_scholes_call(100.0, 95.0, 0.04, 0.25, 0.5)
    slack = 1e-9 if prec == "f64" else 3e-4   # fp32: 64 MUFU exponentials and a TF32x3 mat-vec per path
    assert abs(r.Expected - exact) < 4 * r.std_error + slack, (r.Expected, exact, r.std_error)


def test_cva_is_exactly_zero_without_loss_or_default(engine):
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    for prec in ("f32", "f64"):
        no_loss = engine.cva(m.CVA(0.03, 0.0, opt, 50), 1 << 16, prec)
        assert no_loss.Expected == 0.0 and no_loss.Confidence == 0.0
        no_default = engine.cva(m.CVA(0.0, 0.6, opt, 50), 1 << 16, prec)
        assert no_default.Expected == 0.0 and no_default.Confidence == 0.0


def test_cva_scales_linearly_in_loss_given_default(engine):
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    a = engine.cva(m.CVA(0.03, 0.3, opt, 50), 1 << 20, "f64", seed=7)
    b = engine.cva(m.CVA(0.03, 0.6, opt, 50), 1 << 20, "f64", seed=7)
    # same paths, weights doubled: every product w_j * ee doubles exactly, so do the sums (up to the limb rounding 2^-80)
    assert b.Expected == pytest.approx(2.0 * a.Expected, rel=1e-13)
    assert b.std_error == pytest.approx(2.0 * a.std_error, rel=1e-12)
