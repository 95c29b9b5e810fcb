"""Many jobs per launch (csrc/device_common.cuh mc_accumulate_batch_kernel, mcb200_price_batch).

The reference's cvaOpt driver prices 5 time grids x 4 thread counts as 20 blocking calls
(/root/reference/double_precision/cvaOpt.cu:70-109).  Here the grids of one precision are ONE launch.  The bar is the
ORACLE, not our own one-call path: per-path values against the oracle's (tolerance written below), and every job's
accumulator bit for bit against the oracle's restatement of the chunk reduction and limb split."""
import numpy as np
import pytest

import montecarlocuda_b200 as m
from test_gpu_parity import VAN, make_basket

pytestmark = pytest.mark.gpu

OPT = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
GRIDS = (25, 50, 75, 250, 500)          # cvaOpt.cu:70
SIMS = 131072                           # cvaOpt.cu: PATH 2^17


def _check_job_against_oracle(engine, oracle, workload, params, n_paths, prec, got, seed=m.api.DEFAULT_SEED):
    """got (an OptionValue out of the batch) == closing(oracle's chunk reduction of the per-path values), bit for bit."""
    p = m.plan(workload, params, n_paths, prec)
    vals = getattr(engine, workload + "_paths")(params, 0, n_paths, prec, seed)
    want = m.finalize(p, oracle.accumulate(vals, p))
    assert (got.sum, got.sumsq, got.n_paths, got.Expected, got.Confidence) == (want.sum, want.sumsq, want.n_paths, want.Expected, want.Confidence)
    return vals


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_reference_cva_sweep_is_one_launch_and_matches_the_oracle(engine, oracle, prec):
    jobs = [("cva", m.CVA(0.03, 0.6, OPT, n), SIMS, prec) for n in GRIDS]
    engine.price_batch(jobs)                      # first use: table upload, function attributes
    before = engine.launch_count
    batch = engine.price_batch(jobs)
    assert engine.launch_count - before == 1      # 900 exposure dates of 5 jobs: one multi-job kernel
    for (workload, cva, n_paths, _), got in zip(jobs, batch):
        vals = _check_job_against_oracle(engine, oracle, workload, cva, n_paths, prec, got)
        # per-path values against the oracle's own arithmetic (libm, the reference's formulas) on the same stream
        first = 4096
        ref = oracle.cva_path_values(OPT.s, OPT.k, OPT.r, OPT.v, OPT.t, cva.defInt, cva.lgd, cva.n, m.api.DEFAULT_SEED, 0, first, prec)
        tol = 2e-11 if prec == "f64" else 3e-4     # fp64: hand-built exp / 1/x (2 ulp each) over <= 500 dates; fp32: MUFU chain
        assert np.max(np.abs(vals[:first].astype(np.float64) - ref.astype(np.float64))) < tol
        # and the estimate against the closed form of the device recursion, E[CVA] = LGD C0 sum_j dp_j e^{r t_j} over kept dates
        tau, keep = oracle.cva_grid(OPT.t, cva.n, prec)
        exact = oracle.cva_closed_form(OPT.s, OPT.k, OPT.r, OPT.v, OPT.t, cva.defInt, cva.lgd, cva.n, keep)
        assert abs(got.Expected - exact) < 3.5 * got.std_error + (2e-5 if prec == "f32" else 0.0)


def test_mixed_batch_groups_by_kernel(engine, oracle):
    """Strike sweep of European calls (one launch per precision), CVA grids in both precisions (one each), two
    baskets (one launch each, their factor sits at fixed offsets of a __constant__ table): 6 launches, 17 jobs."""
    strikes = (80.0, 90.0, 100.0, 110.0, 120.0)
    jobs = [("vanilla", m.OptionData(100.0, k, 0.05, 0.2, 1.0), (1 << 18) + 5 * i, "f64") for i, k in enumerate(strikes)]
    jobs += [("vanilla", m.OptionData(100.0, k, 0.05, 0.25, 0.5), 77_777 + i, "f32") for i, k in enumerate(strikes)]
    jobs += [("cva", m.CVA(0.03, 0.6, OPT, n), 40_000, "f64") for n in (3, 50)]
    jobs += [("cva", m.CVA(0.05, 0.4, OPT, n), 30_001, "f32") for n in (7, 25)]
    basket = make_basket(oracle, 10)
    jobs += [("basket", basket, 1 << 16, "f64"), ("basket", basket, 1 << 16, "f32")]   # the SAME object twice
    jobs.append(("vanilla", VAN, 1000, "f64"))   # fewer chunks than CTAs
    engine.price_batch(jobs)
    before = engine.launch_count
    batch = engine.price_batch(jobs)
    assert engine.launch_count - before == 6
    for (workload, params, n_paths, prec), got in zip(jobs, batch):
        _check_job_against_oracle(engine, oracle, workload, params, n_paths, prec, got)


def test_batch_larger_than_one_launch(engine, oracle):
    """30 European calls (more than the 24 jobs a launch carries) and CVA grids adding up to more than the 1024 dates
    the device table holds: split into several launches, every job still exact."""
    jobs = [("vanilla", m.OptionData(100.0, 70.0 + 2 * i, 0.03, 0.3, 2.0), 20_000 + 7 * i, "f64") for i in range(30)]
    jobs += [("cva", m.CVA(0.03, 0.6, OPT, n), 9_000, "f64") for n in (500, 400, 300, 200)]
    before = engine.launch_count
    batch = engine.price_batch(jobs)
    assert engine.launch_count - before == 2 + 2
    for (workload, params, n_paths, prec), got in zip(jobs, batch):
        _check_job_against_oracle(engine, oracle, workload, params, n_paths, prec, got)


def test_batch_reports_failures_per_job(engine):
    import ctypes as C
    from montecarlocuda_b200 import _lib
    lib = _lib.load()
    good, bad = VAN._c(), m.OptionData(-1.0, 100.0, 0.05, 0.2, 1.0)._c()
    arr = (_lib.JobT * 3)(_lib.JobT(_lib.VANILLA, _lib.F64, C.cast(C.pointer(good), C.c_void_p), 10_000, 1),
                          _lib.JobT(_lib.VANILLA, _lib.F64, C.cast(C.pointer(bad), C.c_void_p), 10_000, 1),
                          _lib.JobT(_lib.VANILLA, _lib.F64, C.cast(C.pointer(good), C.c_void_p), 20_000, 1))
    out, status = (_lib.ResultT * 3)(), (C.c_int * 3)()
    assert lib.mcb200_price_batch(engine.handle, 3, arr, out, status) == _lib.ERR_INVALID
    assert list(status) == [_lib.OK, _lib.ERR_INVALID, _lib.OK]
    assert out[0].n_paths == 10_000 and out[2].n_paths == 20_000 and out[1].n_paths == 0
    assert out[0].expected == engine.vanilla(VAN, 10_000, "f64", 1).Expected


def test_sharded_pricer_sees_in_place_parameter_changes(engine, oracle):
    """The parameter dataclasses are mutable: a sweep `opt.k = k; pricer.price(...)` must price the NEW strike each time
    (the pricer's job cache is keyed by value), also for nested fields and for array-valued baskets."""
    from montecarlocuda_b200.distributed import ShardedPricer
    pricer = ShardedPricer(engine=engine, device=engine.device)
    opt = m.OptionData(100.0, 90.0, 0.05, 0.2, 1.0)
    a = pricer.price("vanilla", opt, 1 << 18, "f64", 7)
    opt.k = 110.0
    b = pricer.price("vanilla", opt, 1 << 18, "f64", 7)
    assert a.Expected == engine.vanilla(m.OptionData(100.0, 90.0, 0.05, 0.2, 1.0), 1 << 18, "f64", 7).Expected
    assert b.Expected == engine.vanilla(m.OptionData(100.0, 110.0, 0.05, 0.2, 1.0), 1 << 18, "f64", 7).Expected
    assert a.Expected > b.Expected
    cva = m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), 25)
    c = pricer.price("cva", cva, 1 << 16, "f64", 7)
    cva.option.v = 0.4                      # nested mutation
    d = pricer.price("cva", cva, 1 << 16, "f64", 7)
    assert d.Expected == engine.cva(m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.4, 1.0), 25), 1 << 16, "f64", 7).Expected
    assert d.Expected > c.Expected
    basket = make_basket(oracle, 10)
    e = pricer.price("basket", basket, 1 << 16, "f64", 7)
    basket.w = [0.2] * 5 + [0.0] * 5        # a new list in an array-valued field
    f = pricer.price("basket", basket, 1 << 16, "f64", 7)
    assert f.Expected == engine.basket(basket, 1 << 16, "f64", 7).Expected and f.Expected != e.Expected
