"""GPU parity of the tensor-core basket engine (csrc/basket_tc.cuh, tcgen05 3xTF32) against the CPU oracle,
against the FFMA engine on the same Philox stream, and its order-free combine.

Runs on the B200 (`pytest -m gpu`) through the C ABI (include/mcb200.h: mcb200_set_basket_engine,
mcb200_basket_paths, mcb200_basket, mcb200_basket_launch).  The estimator is brownianVect + basketPayoff of
the reference (DP/MonteCarloKernel.cu:74-101).
"""
import numpy as np
import pytest

import montecarlocuda_b200 as m
from test_gpu_parity import _shard_accumulators, make_basket

pytestmark = pytest.mark.gpu

# fp32 payoffs of S_T ~ 100..300: the MUFU chain (lg2, sqrt, sin, cos, ex2) is good to ~3e-5 relative; the
# 3xTF32 mat-vec adds < 1e-5 absolute to an exponent of order 1, i.e. < 1e-5 relative to the payoff
TOL_F32 = 3e-5 * 300


@pytest.fixture
def tensor_engine(engine):
    m.set_basket_engine(m.BASKET_TENSOR)
    yield engine
    m.set_basket_engine(m.BASKET_TENSOR)


@pytest.mark.parametrize("n_assets", [33, 48, 64])
def test_tensor_paths_match_oracle_and_ffma(tensor_engine, oracle, n_assets):
    opt = make_basket(oracle, n_assets, "f32")
    n, first, seed = 4096 + 77, 1000, 99  # not a multiple of the 256-path CTA round: the ragged tail is exercised
    assert m.get_basket_engine() == m.BASKET_TENSOR
    got = tensor_engine.basket_paths(opt, first, n, "f32", seed).astype(np.float64)
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, seed, first, n, "f32").astype(np.float64)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want)) < TOL_F32
    m.set_basket_engine(m.BASKET_FFMA)
    ffma = tensor_engine.basket_paths(opt, first, n, "f32", seed).astype(np.float64)
    m.set_basket_engine(m.BASKET_TENSOR)
    # same stream, same normals: only the rounding of the mat-vec differs
    assert np.max(np.abs(got - ffma)) < 2e-3
    assert not np.array_equal(got, ffma) or n_assets < 0  # two different kernels really ran


def test_tensor_mat_vec_is_fp32_grade(tensor_engine, oracle):
    """3xTF32 keeps fp32-grade accuracy: the distance to the fp32 oracle (libm, same normals' bits) is no worse than
    the FFMA engine's (both dominated by the MUFU chain, not by the mat-vec), and the operand split adds no bias."""
    opt = make_basket(oracle, 64, "f32")
    n, seed = 8192, 7
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, seed, 0, n, "f32").astype(np.float64)
    tens = tensor_engine.basket_paths(opt, 0, n, "f32", seed).astype(np.float64)
    m.set_basket_engine(m.BASKET_FFMA)
    ffma = tensor_engine.basket_paths(opt, 0, n, "f32", seed).astype(np.float64)
    m.set_basket_engine(m.BASKET_TENSOR)
    err_t, err_f = np.abs(tens - want), np.abs(ffma - want)
    assert err_t.max() < TOL_F32
    assert np.sqrt((err_t ** 2).mean()) < 2.0 * np.sqrt((err_f ** 2).mean()) + 1e-5
    # mean signed difference between the engines far below one standard error of 2^30 paths (4.6e-4)
    assert abs((tens - ffma).mean()) < 5e-5


def test_tensor_full_matrix_factor(tensor_engine, oracle):
    # a caller whose p is not triangular (the reference multiplies the full matrix, MonteCarloKernel.cu:79-84)
    rng = np.random.default_rng(11)
    n = 40
    p = rng.uniform(-0.15, 0.15, (n, n))
    opt = m.MultiOptionData(list(rng.uniform(80, 120, n)), list(rng.uniform(0.1, 0.3, n)), p, list(rng.uniform(-0.02, 0.02, n)),
                            list(np.full(n, 1.0 / n)), 98.0, 0.75, 0.03)
    got = tensor_engine.basket_paths(opt, 0, 2048, "f32", 5).astype(np.float64)
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, 5, 0, 2048, "f32").astype(np.float64)
    assert np.max(np.abs(got - want)) < TOL_F32


def test_tensor_price_matches_ffma_and_sums_its_paths(tensor_engine, oracle):
    opt = make_basket(oracle, 64, "f32")
    n = (1 << 20) + 999
    t = tensor_engine.basket(opt, n, "f32", 2024)
    m.set_basket_engine(m.BASKET_FFMA)
    f = tensor_engine.basket(opt, n, "f32", 2024)
    m.set_basket_engine(m.BASKET_TENSOR)
    assert t.n_paths == n and f.n_paths == n
    # same paths, payoffs differ by rounding only: the two prices are far closer than one standard error
    assert abs(t.Expected - f.Expected) < 0.02 * t.std_error
    assert t.Confidence == pytest.approx(f.Confidence, rel=1e-4)
    # the accumulated sum is the sum of the engine's own per-path values (fp32 runs of <= 64 terms, then exact)
    vals = tensor_engine.basket_paths(opt, 0, 1 << 16, "f32", 2024).astype(np.float64)
    part = tensor_engine.basket(opt, 1 << 16, "f32", 2024)
    assert part.sum == pytest.approx(vals.sum(), rel=2e-6)
    assert part.sumsq == pytest.approx((vals * vals).sum(), rel=2e-6)


@pytest.mark.parametrize("n_assets,n_paths,full", [(64, 70_001, False), (64, 1 << 18, False), (48, 33_333, False), (33, 20_000, False), (40, 25_001, True)])
def test_tensor_accumulator_is_bit_exact_against_the_oracle_restatement(tensor_engine, oracle, n_assets, n_paths, full):
    """Like every other pricing kernel (test_pricing_kernels_sum_exactly_their_path_kernels_values): the accumulator of the
    tensor-core kernel equals, bit for bit, the oracle's restatement of the chunk reduction (fp32 runs per thread, fp64
    butterfly, warps in order, exact limb split) applied to the engine's own per-path values -- whole and ragged chunks,
    triangular and full factors."""
    if full:
        rng = np.random.default_rng(11)
        opt = m.MultiOptionData(list(rng.uniform(80, 120, n_assets)), list(rng.uniform(0.1, 0.3, n_assets)), rng.uniform(-0.15, 0.15, (n_assets, n_assets)),
                                list(rng.uniform(-0.02, 0.02, n_assets)), list(np.full(n_assets, 1.0 / n_assets)), 98.0, 0.75, 0.03)
    else:
        opt = make_basket(oracle, n_assets, "f32")
    p, acc = _shard_accumulators(tensor_engine, "basket", opt, n_paths, "f32", 5, 1)
    vals = tensor_engine.basket_paths(opt, 0, n_paths, "f32", 5)
    assert vals.dtype == np.float32
    assert np.array_equal(acc[0], oracle.accumulate(vals, p))


@pytest.mark.parametrize("n_paths", [300_001, 1 << 21])
def test_tensor_virtual_ranks_bit_identical(tensor_engine, oracle, n_paths):
    opt = make_basket(oracle, 64, "f32")
    results = []
    for world in (1, 2, 3, 4, 8):
        p, acc = _shard_accumulators(tensor_engine, "basket", opt, n_paths, "f32", 2024, world)
        total = acc.sum(axis=0)
        results.append(total)
        assert total[10] == n_paths and total[11] == 0
    for t in results[1:]:
        assert np.array_equal(t, results[0])
    one = tensor_engine.basket(opt, n_paths, "f32", 2024)
    fin = m.finalize(p, results[0])
    assert (one.Expected, one.Confidence, one.sum, one.sumsq) == (fin.Expected, fin.Confidence, fin.sum, fin.sumsq)


def test_engine_switch_only_touches_wide_fp32(tensor_engine, oracle):
    # n <= 32 and fp64 never go to the tensor cores: identical bits whatever the switch says
    for n_assets, prec in ((10, "f32"), (32, "f32"), (64, "f64")):
        opt = make_basket(oracle, n_assets, prec)
        m.set_basket_engine(m.BASKET_TENSOR)
        a = tensor_engine.basket_paths(opt, 0, 1024, prec, 3)
        m.set_basket_engine(m.BASKET_FFMA)
        b = tensor_engine.basket_paths(opt, 0, 1024, prec, 3)
        m.set_basket_engine(m.BASKET_TENSOR)
        assert np.array_equal(a, b)
    with pytest.raises(m.Mcb200Error):
        m.set_basket_engine(7)
