"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libmcb200*.so), against the
CPU oracle on the same Philox stream, the committed golden fixtures, closed forms and -- where it
travelled with the repo -- the reference's own GPU kernels rebuilt for sm_100a (oracle/_ref).

Bars: bit-exact for the integer stages (Philox words, bit-stuffed uniforms, limb accumulators);
for floating point the tolerance is written next to each assertion.
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import montecarlocuda_b200 as m
from montecarlocuda_b200 import _lib as _lib_mod

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
ROOT = Path(__file__).resolve().parents[1]

VAN = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
BS_EXACT = 10.450583572185565


def basket_data(n, rho=0.3):
    c = np.full((n, n), rho)
    np.fill_diagonal(c, 1.0)
    v = [0.3 if i % 2 == 0 else 0.2 for i in range(n)]
    return c, v


def make_basket(oracle, n, precision="f64"):
    c, v = basket_data(n)
    a = oracle.chol(c, precision).astype(np.float64)
    return m.MultiOptionData([100.0] * n, v, a, [0.0] * n, [1.0 / n] * n, 100.0, 1.0, 0.048790164)


CVA50 = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), 50)


# ------------------------------------------------------------------------------------------------
# integer stages: bit-exact
# ------------------------------------------------------------------------------------------------
def test_philox_known_answers_on_device(engine):
    for kat in json.loads((GOLD / "philox_kat.json").read_text()):
        out = engine.philox(np.array([kat["ctr"]], dtype=np.uint32), kat["key"])
        assert out[0].tolist() == kat["out"]


def test_philox_matches_oracle_bit_for_bit(engine, oracle):
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2 ** 32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    key = [0xDEADBEEF, 0x01234567]
    assert np.array_equal(engine.philox(ctr, key), oracle.philox_many(ctr, key))


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_normals_match_oracle(engine, oracle, prec):
    rng = np.random.default_rng(2)
    ctr = rng.integers(0, 2 ** 32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    key = [7, 9]
    words = oracle.philox_many(ctr, key)
    want = np.stack([oracle.normals(w, prec) for w in words]).astype(np.float64)
    got = engine.normals(ctr, key, prec).astype(np.float64)
    # fp64: libm vs device log/sincospi/sqrt, a few ulp of a value of order 1..8
    # fp32: MUFU.LG2/SQRT/SIN/COS approximations (abs error ~2^-21 on sin/cos, 2^-22 rel on lg2/sqrt)
    tol = 2e-14 if prec == "f64" else 4e-6
    assert np.max(np.abs(got - want)) < tol * 8


@pytest.mark.parametrize("in_float", [False, True])
@pytest.mark.parametrize("unit_paths,rounds,n_valid", [(1, 1, 256), (4, 2, 2048), (2, 4, 2048), (1, 8, 1500), (4, 1, 3), (1, 1, 0)])
def test_chunk_reduction_bit_exact(engine, oracle, in_float, unit_paths, rounds, n_valid):
    rng = np.random.default_rng(unit_paths * 100 + rounds)
    vals = np.abs(rng.standard_normal(n_valid)) * 17.0
    if in_float:
        vals = vals.astype(np.float32).astype(np.float64)
    acc = engine.reduce_chunk(vals, unit_paths, rounds, in_float, 73, 66)
    s, s2 = oracle.chunk_reduce(vals, unit_paths, rounds, in_float)
    want = np.zeros(12, dtype=np.uint64)
    assert oracle.lanes_add(s, 73, want[0:5]) == 0 and oracle.lanes_add(s2, 66, want[5:10]) == 0
    want[10] = n_valid
    assert np.array_equal(acc, want)


@pytest.mark.parametrize("value,scale_sum,scale_sumsq", [
    (17.25, 73, 66),                  # the common case: three limbs in the middle of the window
    (3.0, 0, 0), (0.75, 0, 1),        # p < 0: bits below the window's lsb are truncated (sum: floor(3), sumsq: 9 / floor(1.125))
    (2.0 ** -40, 20, 20),             # entirely below the lsb: zero limbs, no error
    (1.0, 158, 105), (1.5, 157, 52),  # top limb: the last values that fit (2^158 resp. the square's 2.25 * 2^52)
    (1.0, 159, 0),                    # 2^159 does not fit the 160-bit window with its headroom bit: error flag
    (2.0 ** 60, 64, -70),             # limb boundaries: off = 12 -> the spill into the third limb starts
    (5e-324, 1074, 0), (2.0 ** -1060, 1100, 1000),   # denormals (the high-word clamp leaves such values behind)
    (1e300, -900, -1900), (1e-300, 1000, 2000),      # far ends of the exponent range
])
def test_limb_split_edge_cases(engine, oracle, value, scale_sum, scale_sumsq):
    """lanes_add runs in integer arithmetic on the device (device_common.cuh) and in floating point in the oracle
    (ldexp, three conversions): the two must agree bit for bit everywhere, including truncation below the window,
    its top limb, denormals, and on WHEN a value is refused."""
    vals = np.array([value])
    acc = engine.reduce_chunk(vals, 1, 1, False, scale_sum, scale_sumsq)
    want = np.zeros(12, dtype=np.uint64)
    bad = oracle.lanes_add(value, scale_sum, want[0:5]) + oracle.lanes_add(value * value, scale_sumsq, want[5:10])
    want[10], want[11] = 1, bad
    assert np.array_equal(acc, want), (acc, want)


@pytest.mark.parametrize("value", [-1.0, float("nan"), float("inf"), -0.0])
def test_limb_split_refuses_what_the_oracle_refuses(engine, oracle, value):
    acc = engine.reduce_chunk(np.array([value]), 1, 1, False, 73, 66)
    want = np.zeros(12, dtype=np.uint64)
    sq = value * value
    bad = oracle.lanes_add(value, 73, want[0:5]) + oracle.lanes_add(sq, 66, want[5:10])
    want[10], want[11] = 1, bad
    assert np.array_equal(acc, want), (acc, want)
    assert (bad > 0) == (not (value >= 0.0) or np.isinf(value))


# ------------------------------------------------------------------------------------------------
# per-path values vs the oracle on the same stream
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_vanilla_paths_match_oracle(engine, oracle, prec):
    n, first, seed = 1 << 14, 4092, 4242  # first path on a draw-unit boundary (6 paths in fp32, 4 in fp64)
    got = engine.vanilla_paths(VAN, first, n, prec, seed).astype(np.float64)
    want = oracle.vanilla_payoffs(VAN.s, VAN.k, VAN.r, VAN.v, VAN.t, seed, first, n, prec).astype(np.float64)
    # fp64: exp(ln S0 + x) vs S0*exp(x): relative 1e-13 of S_T (<= ~300) ; fp32: MUFU chain, 3e-5 relative of S_T
    tol = 1e-10 if prec == "f64" else 3e-5 * 300
    assert np.max(np.abs(got - want)) < tol
    assert (got > 0).sum() == pytest.approx((want > 0).sum(), abs=3)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n_assets", [3, 5, 10, 64])
def test_basket_paths_match_oracle(engine, oracle, prec, n_assets):
    opt = make_basket(oracle, n_assets, prec)
    n, first, seed = 2048, 1000, 99
    got = engine.basket_paths(opt, first, n, prec, seed).astype(np.float64)
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, seed, first, n, prec).astype(np.float64)
    tol = 1e-10 if prec == "f64" else 3e-5 * 300
    assert np.max(np.abs(got - want)) < tol


def test_basket_full_matrix_factor(engine, oracle):
    # a caller whose p is not triangular (the reference multiplies the full matrix, MonteCarloKernel.cu:79-84)
    rng = np.random.default_rng(5)
    n = 4
    p = rng.uniform(-0.5, 0.5, (n, n))
    opt = m.MultiOptionData([100.0, 90.0, 110.0, 95.0], [0.2, 0.25, 0.3, 0.15], p, [0.01, -0.02, 0.0, 0.03], [0.4, 0.3, 0.2, 0.1],
                            98.0, 0.75, 0.03)
    got = engine.basket_paths(opt, 0, 1024, "f64", 5)
    want = oracle.basket_payoffs(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r, 5, 0, 1024, "f64")
    assert np.max(np.abs(got - want)) < 1e-10


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n_dates", [1, 2, 3, 7, 25, 50, 75])   # whole draw blocks, tails of every length, no block at all
def test_cva_paths_match_oracle(engine, oracle, prec, n_dates):
    cva = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), n_dates)
    n, first, seed = 2048, 512, 31337
    got = engine.cva_paths(cva, first, n, prec, seed).astype(np.float64)
    want = oracle.cva_path_values(100.0, 100.0, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, seed, first, n, prec).astype(np.float64)
    # path CVA is ~0.2 (up to ~2): fp64 differs by hoisted constants only; fp32 accumulates n_dates MUFU-accurate
    # exposures and the float time grid, hence the looser bar
    tol = 1e-11 if prec == "f64" else 2e-4
    assert np.max(np.abs(got - want)) < tol


@pytest.mark.parametrize("prec,n_dates", [("f64", 1025), ("f64", 1500), ("f32", 1100), ("f64", 4098)])
def test_cva_long_grid_reads_its_dates_from_device_memory(engine, oracle, prec, n_dates):
    """More kept dates than the 1024 of the constant table (the reference has no limit: cva->n is a plain int,
    DP/MonteCarloKernel.cu:247): same kernel, date table in device memory.  Per-path values against the oracle,
    the accumulator against the oracle's restatement applied to them bit for bit, the price against the closed form."""
    cva = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), n_dates)
    n, seed = 1024, 99
    got = engine.cva_paths(cva, 0, n, prec, seed)
    want = oracle.cva_path_values(100.0, 100.0, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, seed, 0, n, prec).astype(np.float64)
    tol = 2e-11 if prec == "f64" else 2e-3      # the error budget of test_cva_paths_match_oracle, for 20-80 times the dates
    assert np.max(np.abs(got.astype(np.float64) - want)) < tol
    p = m.plan("cva", cva, n, prec)
    r = engine.cva(cva, n, prec, seed)
    assert r == m.finalize(p, oracle.accumulate(got, p))
    if prec == "f64":
        r = engine.cva(cva, 1 << 18, prec, seed)
        _, keep = oracle.cva_grid(1.0, n_dates, prec)
        closed = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, keep)
        assert abs(r.Expected - closed) < 4 * r.std_error + 1e-6
    # a short grid right after it: the constant-table path is untouched by the long one
    short = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), 50)
    a = engine.cva(short, 4096, prec, seed)
    b = engine.price_batch([("cva", short, 4096, prec), ("cva", cva, 512, prec), ("cva", short, 4096, prec)], seed)
    assert b[0] == a and b[2] == a and b[1] == engine.cva(cva, 512, prec, seed)


# ------------------------------------------------------------------------------------------------
# prices: closed forms, golden fixtures, reference GPU kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_vanilla_price_vs_black_scholes(engine, prec):
    r = engine.vanilla(VAN, 1 << 24, prec)
    assert r.n_paths == 1 << 24
    assert abs(r.Expected - BS_EXACT) < 3 * r.std_error          # north_star: within 3 standard errors
    assert r.Confidence == pytest.approx(1.96 * r.std_error * np.exp(0.05), rel=1e-12)   # Q5: undiscounted half-width


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_accumulated_sums_match_oracle(engine, oracle, prec):
    # the whole pipeline (kernel + limbs + closing) against the oracle's per-path values summed in numpy
    n, seed = 200_000, 17
    r = engine.vanilla(VAN, n, prec, seed)
    pay = oracle.vanilla_payoffs(VAN.s, VAN.k, VAN.r, VAN.v, VAN.t, seed, 0, n, prec).astype(np.float64)
    rel = 1e-12 if prec == "f64" else 2e-6
    assert r.sum == pytest.approx(pay.sum(), rel=rel)
    assert r.sumsq == pytest.approx((pay * pay).sum(), rel=rel)
    e, c = oracle.closing(r.sum, r.sumsq, n, VAN.r, VAN.t)
    assert r.Expected == pytest.approx(e, rel=1e-13) and r.Confidence == pytest.approx(c, rel=1e-9)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_cva_price_vs_closed_form(engine, oracle, prec):
    r = engine.cva(CVA50, 1 << 22, prec)
    _, keep = oracle.cva_grid(1.0, 50, prec)
    closed = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, 50, keep)
    # the Hastings cnd biases each exposure by < 1e-6 relative; 3 SE at 2^22 paths is ~2e-4
    assert abs(r.Expected - closed) < 3 * r.std_error + 1e-6
    assert r.Expected == pytest.approx(r.mean)  # the CVA is not discounted


def test_cva_exact_grid_keeps_maturity(engine, oracle):
    cva = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), 50, grid_mode=1)
    r = engine.cva(cva, 1 << 22, "f64")
    closed = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, 50, np.ones(50))
    assert abs(r.Expected - closed) < 3 * r.std_error + 1e-6


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_basket_price_vs_golden_reference_host(engine, oracle, prec):
    # SP host basket is the sound reference CPU path (the DP host drops the volatility, Q1)
    gold = [e for e in json.loads((GOLD / "ref_host_mc.json").read_text()) if e["workload"] == "basket" and e["precision"] == "sp"]
    for entry in gold:
        if entry["corr"] != "equicorr_0.3":
            continue
        opt = make_basket(oracle, entry["n"], prec)
        r = engine.basket(opt, 1 << 22, prec)
        ref_se = entry["Confidence"] / 1.96 * np.exp(-0.048790164)
        assert abs(r.Expected - entry["Expected"]) < 3 * np.hypot(r.std_error, ref_se)


def _reference(precision, n):
    from oracle_lib import Reference
    try:
        return Reference(precision, n)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_against_reference_gpu_kernels(engine, oracle, precision, capfd):
    """basket and CVA agree with the reference's own GPU path within 3 combined standard errors."""
    prec = {"dp": "f64", "sp": "f32"}[precision]
    sims = 1 << 22
    ref = _reference(precision, 3)
    rv = ref.lib.dev_vanillaOpt(ref.option(100, 100, 0.05, 0.2, 1.0), 512, 128, sims)
    ours = engine.vanilla(VAN, sims, prec)
    assert abs(ours.Expected - rv.Expected) < 3 * np.hypot(ours.Confidence, rv.Confidence) / 1.96
    ref10 = _reference(precision, 10)
    opt = make_basket(oracle, 10, prec)
    rb = ref10.lib.dev_basketOpt(ref10.multi(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r), 512, 128, sims)
    ob = engine.basket(opt, sims, prec)
    assert abs(ob.Expected - rb.Expected) < 3 * np.hypot(ob.Confidence, rb.Confidence) / 1.96
    rc = ref.lib.dev_cvaEquityOption(ref.cva(0.03, 0.6, ref.option(100, 100, 0.05, 0.2, 1.0), 50), 1024, 128, 1 << 20)
    oc = engine.cva(CVA50, 1 << 20, prec)
    assert abs(oc.Expected - rc.Expected) < 3 * np.hypot(oc.Confidence, rc.Confidence) / 1.96
    capfd.readouterr()  # the reference prints timing lines on every call


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_wide_basket_against_reference_gpu_kernel(engine, oracle, precision, capfd):
    """BASELINE config 5's width: the 64-asset basket against the reference's own GPU kernel (N = 64 build) at 2^22
    simulations, within 3 combined standard errors -- fp32 through the tensor-core engine AND the packed-FMA engine."""
    prec = {"dp": "f64", "sp": "f32"}[precision]
    sims = 1 << 22
    ref = _reference(precision, 64)
    opt = make_basket(oracle, 64, prec)
    rb = ref.lib.dev_basketOpt(ref.multi(opt.s, opt.v, opt.p, opt.d, opt.w, opt.k, opt.t, opt.r), 512, 128, sims)
    capfd.readouterr()
    ours = [engine.basket(opt, sims, prec)]
    if prec == "f32":
        m.set_basket_engine(m.BASKET_FFMA)
        ours.append(engine.basket(opt, sims, prec))
        m.set_basket_engine(m.BASKET_TENSOR)
    for ob in ours:
        assert abs(ob.Expected - rb.Expected) < 3 * np.hypot(ob.Confidence, rb.Confidence) / 1.96


@pytest.mark.parametrize("precision", ["dp", "sp"])
@pytest.mark.parametrize("n_dates", [25, 75, 250, 500])
def test_cva_grids_against_reference_gpu_kernel(engine, oracle, precision, n_dates, capfd):
    """The other grids of the reference's cvaOpt sweep (cvaOpt.cu:70), each with its own keep / drop of the last
    date (Q3: 1 - n * (1/n) rounds differently per n and per precision): ours against the reference's GPU kernel
    within 3 combined standard errors, and the kept-date pattern against the oracle's closed form."""
    prec = {"dp": "f64", "sp": "f32"}[precision]
    sims = 1 << 22          # standard error 6.7e-5: the two last-date patterns are 3.8e-4 (n = 500) to 7.7e-3 (n = 25) apart
    ref = _reference(precision, 3)
    rc = ref.lib.dev_cvaEquityOption(ref.cva(0.03, 0.6, ref.option(100, 100, 0.05, 0.2, 1.0), n_dates), 1024, 128, sims)
    capfd.readouterr()
    oc = engine.cva(m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), n_dates), sims, prec)
    assert abs(oc.Expected - rc.Expected) < 3 * np.hypot(oc.Confidence, rc.Confidence) / 1.96
    _, keep = oracle.cva_grid(1.0, n_dates, prec)
    kept = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, keep)
    flipped = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, np.concatenate([keep[:-1], [1 - keep[-1]]]))
    # the two patterns differ by the last date's weight x its exposure (~ dp_n x intrinsic value): the estimate must side
    # with the grid's own rounding, and clearly so
    assert abs(oc.Expected - kept) < 3 * oc.std_error + 2e-5
    assert abs(oc.Expected - flipped) > abs(oc.Expected - kept)


# ------------------------------------------------------------------------------------------------
# order-free combine: any partition of the chunks gives the same bits
# ------------------------------------------------------------------------------------------------
def _shard_accumulators(engine, workload, params, n_paths, prec, seed, world):
    import torch
    from montecarlocuda_b200 import distributed as D
    p = m.plan(workload, params, n_paths, prec)
    acc = torch.zeros((world, 12), dtype=torch.int64, device="cuda:0")
    for rank in range(world):
        first, count = m.shard_range(p, rank, world)
        D._launch(engine, workload, p, params, seed, first, count, acc[rank], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return p, acc.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("workload,prec,n_paths", [("vanilla", "f32", (1 << 22) + 12345), ("vanilla", "f64", 1 << 22),
                                                   ("basket", "f64", 300_001), ("cva", "f64", 1 << 18), ("cva", "f32", 77_777)])
def test_virtual_ranks_bit_identical(engine, oracle, workload, prec, n_paths):
    params = {"vanilla": VAN, "basket": make_basket(oracle, 10), "cva": CVA50}[workload]
    results = []
    for world in (1, 2, 3, 4, 8):
        p, acc = _shard_accumulators(engine, workload, params, n_paths, prec, 2024, world)
        total = acc.sum(axis=0)
        results.append(total)
        assert total[10] == n_paths and total[11] == 0
    for t in results[1:]:
        assert np.array_equal(t, results[0])
    one = getattr(engine, workload)(params, n_paths, prec, 2024)
    fin = m.finalize(p, results[0])
    assert (one.Expected, one.Confidence, one.sum, one.sumsq) == (fin.Expected, fin.Confidence, fin.sum, fin.sumsq)


def test_accumulator_matches_oracle_restatement(engine, oracle):
    # chunk geometry + limb split restated on the CPU from the DEVICE's per-path values: bit-exact
    n = 70_000
    p, acc = _shard_accumulators(engine, "vanilla", VAN, n, "f64", 5, 1)
    vals = engine.vanilla_paths(VAN, 0, n, "f64", 5)
    assert np.array_equal(acc[0], oracle.accumulate(vals, p))
    p, acc = _shard_accumulators(engine, "vanilla", VAN, n, "f32", 5, 1)
    vals = engine.vanilla_paths(VAN, 0, n, "f32", 5)
    assert np.array_equal(acc[0], oracle.accumulate(vals, p))


@pytest.mark.parametrize("workload,n_assets,prec", [("basket", 10, "f64"), ("basket", 16, "f64"), ("basket", 8, "f64"), ("basket", 3, "f64"),
                                                    ("basket", 10, "f32"), ("basket", 64, "f64"), ("basket", 40, "f64"), ("cva", 0, "f64"), ("cva", 0, "f32"), ("cva", 7, "f64"), ("cva", 3, "f32"),
                                                    ("cva", 2, "f64")])
def test_pricing_kernels_sum_exactly_their_path_kernels_values(engine, oracle, workload, n_assets, prec):
    """The pricing kernels run a different memory layout from the per-path kernels the oracle is compared with
    (bank-conflict-free replicated fp64 tables, sub-block CTAs, the fp64 factor in shared memory, the two-pass sweep of
    the 64-asset fp64 basket, whole draw blocks of CVA dates) but the SAME arithmetic: the accumulator of a job equals, bit for bit, the oracle's restatement
    of the chunk reduction applied to the per-path kernel's values."""
    # for the CVA the second parameter is the number of exposure dates (0: the 50 of BASELINE config 4)
    params = (make_basket(oracle, n_assets, prec) if workload == "basket" else
              CVA50 if n_assets == 0 else m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), n_assets))
    n = 40_000 if n_assets < 64 else 9_000
    p, acc = _shard_accumulators(engine, workload, params, n, prec, 5, 1)
    vals = getattr(engine, workload + "_paths")(params, 0, n, prec, 5)
    assert np.array_equal(acc[0], oracle.accumulate(vals, p))


# ------------------------------------------------------------------------------------------------
# the drop-in libraries: the reference's symbols, struct layouts and semantics
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_dropin_symbols(engine, oracle, precision):
    real = C.c_double if precision == "dp" else C.c_float
    lib = C.CDLL(str(ROOT / "montecarlocuda_b200" / "lib" / f"libmcb200_{precision}.so"))

    class OptionData(C.Structure):
        _fields_ = [(k, real) for k in "skrvt"]

    class OptionValue(C.Structure):
        _fields_ = [("Expected", real), ("Confidence", real)]

    class MultiOptionData(C.Structure):
        _fields_ = [("s", real * 3), ("v", real * 3), ("p", (real * 3) * 3), ("d", real * 3), ("w", real * 3), ("k", real), ("t", real), ("r", real)]

    class CVA(C.Structure):
        _fields_ = [("defInt", real), ("lgd", real), ("ns", C.c_int), ("option", OptionData), ("n", C.c_int)]

    for fn, arg in ((lib.dev_vanillaOpt, OptionData), (lib.dev_basketOpt, MultiOptionData), (lib.dev_cvaEquityOption, CVA)):
        fn.restype, fn.argtypes = OptionValue, [C.POINTER(arg), C.c_int, C.c_int, C.c_int]
    prec = {"dp": "f64", "sp": "f32"}[precision]
    # n = numBlocks * (sims / numBlocks): 512 * (1000000 // 512) = 999936 paths (reference :508)
    v = lib.dev_vanillaOpt(OptionData(100, 100, 0.05, 0.2, 1.0), 512, 128, 1_000_000)
    # the SP library receives float fields: give our side the same rounded parameters
    ours = engine.vanilla(m.OptionData(100.0, 100.0, real(0.05).value, real(0.2).value, 1.0), 999_936, prec)
    assert v.Expected == real(ours.Expected).value and v.Confidence == real(ours.Confidence).value
    mo = MultiOptionData()
    a = oracle.chol(np.array([[1, .3, .3], [.3, 1, .3], [.3, .3, 1.0]]), prec)
    for i in range(3):
        mo.s[i], mo.v[i], mo.d[i], mo.w[i] = 100, [0.2, 0.3, 0.2][i], 0, 1 / 3
        for j in range(3):
            mo.p[i][j] = a[i][j]
    mo.k, mo.t, mo.r = 100, 1, 0.048790164
    b = lib.dev_basketOpt(mo, 512, 128, 1 << 20)
    py = m.MultiOptionData([real(100).value] * 3, [real(x).value for x in (0.2, 0.3, 0.2)], a.astype(np.float64), [0.0] * 3,
                           [real(1 / 3).value] * 3, 100.0, 1.0, real(0.048790164).value)
    ob = engine.basket(py, 1 << 20, prec)
    assert b.Expected == real(ob.Expected).value
    c = lib.dev_cvaEquityOption(CVA(0.03, 0.6, 1, OptionData(100, 100, 0.05, 0.2, 1.0), 50), 1024, 128, 131072)
    oc = engine.cva(m.CVA(real(0.03).value, real(0.6).value, m.OptionData(100, 100, real(0.05).value, real(0.2).value, 1.0), 50), 131072, prec)
    assert c.Expected == real(oc.Expected).value


def test_python_mirror_of_reference_entry_points(engine):
    v = m.dev_vanillaOpt(VAN, 512, 128, 1 << 20)
    assert v.n_paths == 1 << 20 and abs(v.Expected - BS_EXACT) < 4 * v.std_error


# ------------------------------------------------------------------------------------------------
# edge cases and full-size properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n_paths", [1, 2, 3, 5, 255, 256, 257, 1023, 1025, 65_537])
def test_ragged_path_counts(engine, oracle, prec, n_paths):
    r = engine.vanilla(VAN, n_paths, prec, 3)
    pay = oracle.vanilla_payoffs(VAN.s, VAN.k, VAN.r, VAN.v, VAN.t, 3, 0, n_paths, prec).astype(np.float64)
    assert r.n_paths == n_paths
    assert r.sum == pytest.approx(pay.sum(), rel=3e-6, abs=1e-3)


def test_invalid_arguments_fail_loudly(engine):
    with pytest.raises(m.Mcb200Error):
        engine.vanilla(VAN, 0)
    with pytest.raises(m.Mcb200Error):
        engine.vanilla(m.OptionData(-1.0, 100, 0.05, 0.2, 1.0), 1000)
    with pytest.raises(m.Mcb200Error):
        engine.vanilla(m.OptionData(float("nan"), 100, 0.05, 0.2, 1.0), 1000)
    with pytest.raises(m.Mcb200Error):
        engine.cva(m.CVA(0.03, 0.6, VAN, (1 << 20) + 1), 1000)      # MCB200_MAX_DATES
    n = 257                                                         # MCB200_MAX_ASSETS + 1
    with pytest.raises(m.Mcb200Error):
        engine.basket(m.MultiOptionData([100.0] * n, [0.2] * n, np.eye(n), [0.0] * n, [1 / n] * n, 100.0, 1.0, 0.05), 1000)
    with pytest.raises(m.Mcb200Error):
        engine.vanilla_paths(VAN, 3, 16, "f32")   # not on a draw-unit boundary


def test_duplicate_contexts_are_rejected_not_deadlocked(engine):
    """mcb200_*_multi locks every context it is given: the same context twice used to deadlock the call on itself."""
    lib = _lib_mod.load()
    handles = (C.c_void_p * 2)(engine.handle, engine.handle)
    r, c = _lib_mod.ResultT(), VAN._c()
    assert lib.mcb200_vanilla_multi(handles, 2, _lib_mod.F64, C.byref(c), 1 << 16, 1, C.byref(r)) == _lib_mod.ERR_INVALID
    one = (C.c_void_p * 1)(engine.handle)
    assert lib.mcb200_vanilla_multi(one, 1, _lib_mod.F64, C.byref(c), 1 << 16, 1, C.byref(r)) == _lib_mod.OK   # the context is still usable
    assert r.n_paths == 1 << 16


def test_high_volatility_many_dates_is_a_valid_cva(engine, oracle):
    """The range check of the table-driven exponential bounds what a PATH can reach (|ln s/K| + |drift| T + 9 v sqrt(T)),
    not n times what one step can: 500 dates at 150 % volatility were refused before (n x 9 v sqrt(dt) = 9 v sqrt(n T))."""
    cva = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 1.5, 1.0), 500)
    r = engine.cva(cva, 1 << 16, "f64")
    _, keep = oracle.cva_grid(1.0, 500, "f64")
    closed = oracle.cva_closed_form(100, 100, 0.05, 1.5, 1.0, 0.03, 0.6, 500, keep)
    assert abs(r.Expected - closed) < 4 * r.std_error + 1e-5
    vals = engine.cva_paths(cva, 0, 2048, "f64")
    want = oracle.cva_path_values(100.0, 100.0, 0.05, 1.5, 1.0, 0.03, 0.6, 500, m.api.DEFAULT_SEED, 0, 2048, "f64")
    assert np.max(np.abs(vals - want) / (1.0 + np.abs(want))) < 1e-10
    with pytest.raises(m.Mcb200Error):      # ... and what a path really cannot do is still refused
        engine.cva(m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 80.0, 1.0), 10), 1000, "f64")


def test_kernel_time_is_reported_on_request_only(engine):
    a = engine.vanilla(VAN, 1 << 20, "f64")
    assert a.kernel_ms == 0.0
    engine.set_timing(True)
    try:
        b = engine.vanilla(VAN, 1 << 24, "f64")
        assert 0.01 < b.kernel_ms < 50.0
    finally:
        engine.set_timing(False)
    assert engine.vanilla(VAN, 1 << 20, "f64").Expected == a.Expected


def test_degenerate_parameters(engine):
    # zero volatility: every path pays max(S0 e^{rT} - K, 0) exactly; zero maturity: intrinsic value
    r = engine.vanilla(m.OptionData(100.0, 90.0, 0.05, 0.0, 1.0), 4096, "f64")
    assert r.Expected == pytest.approx(100.0 - 90.0 * np.exp(-0.05), rel=1e-13) and r.Confidence < 1e-7  # n*sum(x^2) - sum(x)^2 cancels to rounding noise
    r = engine.vanilla(m.OptionData(100.0, 90.0, 0.05, 0.3, 0.0), 4096, "f64")
    assert r.Expected == pytest.approx(10.0, rel=1e-13)
    r = engine.vanilla(m.OptionData(100.0, 1e6, 0.05, 0.2, 1.0), 4096, "f32")
    assert r.Expected == 0.0 and r.Confidence == 0.0


def test_full_size_vanilla_fp32_2pow32(engine):
    """BASELINE config 2 at full size: 2^32 paths, fp32, within 3 SE of closed-form Black-Scholes,
    and the path count is exact (every chunk ran exactly once)."""
    r = engine.vanilla(VAN, 1 << 32, "f32")
    assert r.n_paths == 1 << 32
    assert abs(r.Expected - BS_EXACT) < 3 * r.std_error
    assert r.std_error < 2.5e-4


def test_full_size_vanilla_fp64_2pow32(engine):
    """BASELINE config 2, fp64 (the headline): 2^32 paths within 3 SE of closed-form Black-Scholes (north_star)."""
    r = engine.vanilla(VAN, 1 << 32, "f64")
    assert r.n_paths == 1 << 32
    assert abs(r.Expected - BS_EXACT) < 3 * r.std_error and r.std_error < 2.5e-4


def test_full_size_cva_fp64_2pow26(engine, oracle):
    """BASELINE config 4: CVA, 50 exposure dates, 2^26 paths fp64: within 3 SE (1.7e-5) of the closed form
    E[CVA] = LGD C0 sum_j dp_j e^{r t_j} over the dates the reference's fp64 grid keeps."""
    r = engine.cva(CVA50, 1 << 26, "f64")
    _, keep = oracle.cva_grid(1.0, 50, "f64")
    closed = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, 50, keep)
    assert r.n_paths == 1 << 26
    assert abs(r.Expected - closed) < 3 * r.std_error + 1e-6     # + the Hastings cnd's own bias
    assert 1.5e-5 < r.std_error < 2.0e-5


def test_full_size_baskets_agree_across_precisions_and_engines(engine, oracle):
    """BASELINE configs 3 and 5 at full size.  No closed form exists for a basket: the anchors are the golden
    reference host runs (tests above) and, here, agreement between independent code paths on the full job:
    N=10 fp64 (2^28 paths) against its fp32 kernel, N=64 fp32 (2^30 paths) tensor-core engine against the FFMA
    engine -- same Philox positions, different arithmetic -- within 3 combined standard errors."""
    b10 = make_basket(oracle, 10)
    d = engine.basket(b10, 1 << 28, "f64")
    f = engine.basket(make_basket(oracle, 10, "f32"), 1 << 28, "f32")
    assert d.n_paths == f.n_paths == 1 << 28
    assert abs(d.Expected - f.Expected) < 3 * np.hypot(d.std_error, f.std_error)
    assert 8.5 < d.Expected < 8.8 and d.std_error < 8e-4
    b64 = make_basket(oracle, 64, "f32")
    m.set_basket_engine(m.BASKET_TENSOR)
    t = engine.basket(b64, 1 << 30, "f32")
    m.set_basket_engine(m.BASKET_FFMA)
    try:
        g = engine.basket(b64, 1 << 30, "f32")
    finally:
        m.set_basket_engine(m.BASKET_TENSOR)
    assert t.n_paths == g.n_paths == 1 << 30
    # the same paths with different rounding: far closer than the Monte Carlo error (3.2e-4)
    assert abs(t.Expected - g.Expected) < 0.1 * t.std_error
    assert 8.0 < t.Expected < 8.3 and t.std_error < 4e-4


def test_linearity_in_notional(engine):
    # payoff is positively homogeneous: scaling S0 and K by 2 (an exact power of two) scales every path
    # value by exactly 2, so sum doubles and sumsq quadruples bit for bit (fp64 and the limb split are
    # both exact under power-of-two scaling)
    a = engine.vanilla(VAN, 1 << 20, "f64", 9)
    b = engine.vanilla(m.OptionData(200.0, 200.0, 0.05, 0.2, 1.0), 1 << 20, "f64", 9)
    assert b.sum == pytest.approx(2 * a.sum, rel=1e-13) and b.sumsq == pytest.approx(4 * a.sumsq, rel=1e-13)
