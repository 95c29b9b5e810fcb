"""The reference's CPU helper API re-implemented in csrc/hostapi.c (libmcb200_hostapi_*.so), checked
against the reference's own outputs (tests/golden) and against the oracle.  No GPU needed."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def hostapi(precision, n=3):
    suffix = precision if n == 3 else f"{precision}_n{n}"
    lib = C.CDLL(str(ROOT / "montecarlocuda_b200" / "lib" / f"libmcb200_hostapi_{suffix}.so"))
    real = C.c_double if precision == "dp" else C.c_float

    class OptionData(C.Structure):
        _fields_ = [(k, real) for k in "skrvt"]

    class OptionValue(C.Structure):
        _fields_ = [("Expected", real), ("Confidence", real)]

    class MultiOptionData(C.Structure):
        _fields_ = [("s", real * n), ("v", real * n), ("p", (real * n) * n), ("d", real * n), ("w", real * n), ("k", real), ("t", real), ("r", real)]

    class CVA(C.Structure):
        _fields_ = [("defInt", real), ("lgd", real), ("ns", C.c_int), ("option", OptionData), ("n", C.c_int)]

    lib.host_bsCall.restype, lib.host_bsCall.argtypes = real, [OptionData]
    lib.host_vanillaOpt.restype, lib.host_vanillaOpt.argtypes = OptionValue, [OptionData, C.c_int]
    lib.host_basketOpt.restype, lib.host_basketOpt.argtypes = OptionValue, [C.POINTER(MultiOptionData), C.c_int]
    lib.host_cvaEquityOption.restype, lib.host_cvaEquityOption.argtypes = OptionValue, [C.POINTER(CVA), C.c_int]
    lib.Chol.restype, lib.Chol.argtypes = None, [C.c_void_p, C.c_void_p]
    lib.mcb200_chol.restype, lib.mcb200_chol.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_void_p]
    return lib, dict(OptionData=OptionData, OptionValue=OptionValue, MultiOptionData=MultiOptionData, CVA=CVA, real=real,
                     np_real=np.float64 if precision == "dp" else np.float32)


@pytest.fixture(scope="module", autouse=True)
def built(ensure_built):
    return ensure_built


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_host_bscall_matches_reference(precision):
    lib, t = hostapi(precision)
    for case in json.loads((GOLD / "ref_bscall.json").read_text())[precision]:
        got = lib.host_bsCall(t["OptionData"](*case["args"]))
        if precision == "dp":
            assert got == case["value"]                               # bit for bit with the reference's host_bsCall
        else:
            assert abs(got - case["value"]) <= 4e-6 * max(1.0, abs(case["value"]))   # the SP reference mixes in double


def test_chol_matches_reference_bit_for_bit():
    for case in json.loads((GOLD / "ref_chol.json").read_text()):
        lib, t = hostapi(case["precision"], case["n"])
        c = np.ascontiguousarray(case["c"], dtype=t["np_real"])
        a = np.zeros_like(c)
        lib.Chol(c.ctypes.data, a.ctypes.data)
        assert np.array_equal(a.astype(float), np.array(case["a"])), (case["n"], case["precision"], case["name"])


def test_robust_chol_reports_non_pd_input():
    lib, _ = hostapi("dp")
    pd = np.full((10, 10), 0.3)
    np.fill_diagonal(pd, 1.0)
    a = np.zeros_like(pd)
    assert lib.mcb200_chol(10, pd.ctypes.data, a.ctypes.data) == 0
    assert np.allclose(a, np.linalg.cholesky(pd), atol=1e-14)
    singular = np.array([[1, -.5, -.5], [-.5, 1, -.5], [-.5, -.5, 1.0]])      # the reference driver's default (rank 2)
    a3 = np.zeros_like(singular)
    assert lib.mcb200_chol(3, singular.ctypes.data, a3.ctypes.data) == 3      # third pivot is not positive
    indefinite = np.array([[1.0, 2.0], [2.0, 1.0]])
    assert lib.mcb200_chol(2, indefinite.ctypes.data, np.zeros((2, 2)).ctypes.data) == 2


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_host_estimators_walk_the_engine_stream(oracle, precision, monkeypatch):
    """host_vanillaOpt uses the engine's Philox stream: same paths as the oracle (and the GPU) for a seed."""
    monkeypatch.setenv("MCB200_SEED", "12345")
    lib, t = hostapi(precision)
    prec = {"dp": "f64", "sp": "f32"}[precision]
    n = 100_003
    v = lib.host_vanillaOpt(t["OptionData"](100, 100, 0.05, 0.2, 1.0), n)
    pay = oracle.vanilla_payoffs(100, 100, 0.05, 0.2, 1.0, 12345, 0, n, prec).astype(np.float64)
    e, c = oracle.closing(pay.sum(), (pay * pay).sum(), n, float(t["real"](0.05).value), 1.0)
    assert v.Expected == pytest.approx(e, rel=1e-12 if precision == "dp" else 2e-6)
    assert v.Confidence == pytest.approx(c, rel=1e-9 if precision == "dp" else 2e-6)
    cva = t["CVA"](0.03, 0.6, 1, t["OptionData"](100, 100, 0.05, 0.2, 1.0), 50)
    r = lib.host_cvaEquityOption(cva, 4096)
    val = oracle.cva_path_values(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, 50, 12345, 0, 4096, prec).astype(np.float64)
    assert r.Expected == pytest.approx(val.mean(), rel=1e-11 if precision == "dp" else 5e-5)


def test_host_basket_is_the_sound_estimator(oracle, monkeypatch):
    monkeypatch.setenv("MCB200_SEED", "7")
    lib, t = hostapi("dp", 10)
    c = np.full((10, 10), 0.3)
    np.fill_diagonal(c, 1.0)
    a = oracle.chol(c)
    vol = [0.3 if i % 2 == 0 else 0.2 for i in range(10)]
    m = t["MultiOptionData"]()
    for i in range(10):
        m.s[i], m.v[i], m.d[i], m.w[i] = 100, vol[i], 0, 0.1
        for j in range(10):
            m.p[i][j] = a[i][j]
    m.k, m.t, m.r = 100, 1, 0.048790164
    r = lib.host_basketOpt(m, 20000)
    pay = oracle.basket_payoffs([100] * 10, vol, a, [0] * 10, [0.1] * 10, 100.0, 1.0, 0.048790164, 7, 0, 20000, "f64")
    assert r.Expected == pytest.approx(np.exp(-0.048790164) * pay.mean(), rel=1e-11)
    # the volatility stays in the diffusion (the reference DP host drops it: 73.04 instead of 8.63)
    assert 8.0 < r.Expected < 9.3
