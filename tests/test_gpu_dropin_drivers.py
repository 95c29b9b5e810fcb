"""The reference's own driver programs (vanillaOpt.cu, basketOpt.cu, cvaOpt.cu + MonteCarloHost.c,
compiled UNMODIFIED by oracle/Makefile `drivers`) linked against libmcb200_{dp,sp}.so instead of the
reference's MonteCarloKernel.o, and run on the GPU: the drop-in claim of INTEGRATION.md section 1."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import montecarlocuda_b200 as m

pytestmark = pytest.mark.gpu
REF = Path(__file__).resolve().parents[1] / "oracle" / "_ref"


def run_driver(name, precision, stdin=""):
    exe = REF / f"{name}_{precision}_mcb200"
    if not exe.exists():
        pytest.skip(f"{exe.name} not built (needs /root/reference at build time)")
    res = subprocess.run([str(exe)], input=stdin, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def floats_after(text, marker, count):
    tail = text[text.index(marker) + len(marker):]
    return [float(x) for x in re.findall(r"-?\d+\.\d+", tail)[:count]]


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_reference_vanilla_driver_on_our_library(engine, precision):
    out = run_driver("vanillaOpt", precision, "8\n")         # 8 x 131072 = 2^20 paths (vanillaOpt.cu:51-53)
    bs = floats_after(out, "Prezzo Black & Scholes:", 1)[0]
    gpu = re.search(r"Simulated price for the option with GPU:.*?\n128 \n([-\d.]+) \n([-\d.]+) \n([-\d.]+) \n", out, re.S)
    price, conf, diff = (float(gpu.group(i)) for i in (1, 2, 3))
    # driver parameters: S=K=100, R=0.048790, V=0.2, T=1 (vanillaOpt.cu:22-26); 512 blocks x 128 threads
    rate, vol = (float(np.float32(0.048790)), float(np.float32(0.2))) if precision == "sp" else (0.048790, 0.2)
    ours = engine.vanilla(m.OptionData(100.0, 100.0, rate, vol, 1.0), 1 << 20,
                          "f64" if precision == "dp" else "f32")
    assert price == pytest.approx(ours.Expected, abs=1.5e-6)          # %f prints 6 decimals
    assert conf == pytest.approx(ours.Confidence, abs=1.5e-6)
    assert abs(price - bs) < 4 * conf / 1.96                          # the driver's own self-check
    if precision == "dp":
        assert "Numero di simulazioni" not in out                    # no per-call chatter from the engine


def test_reference_basket_driver_on_our_library(engine):
    out = run_driver("basketOpt", "sp", "8\n")                        # SP: the host path there is sound (Q1)
    cpu = floats_after(out, "Expected price, I.C., time", 2)
    gpu = re.search(r"Simulated price for the option with GPU:.*?\n128 \n([-\d.]+) \n([-\d.]+) \n", out, re.S)
    price, conf = float(gpu.group(1)), float(gpu.group(2))
    # GPU (ours) vs the reference's CPU estimator inside the same driver run: 3 combined standard errors
    assert abs(price - cpu[0]) < 3 * np.hypot(conf, cpu[1]) / 1.96


def test_reference_cva_driver_on_our_library(engine, oracle):
    out = run_driver("cvaOpt", "dp")                                  # 5 grids x 4 thread counts, 131072 paths each
    values = [float(x) for x in re.findall(r"CVA: \n([-\d.]+)", out)]
    assert len(values) == 20
    for k, n_dates in enumerate((25, 50, 75, 250, 500)):
        block = values[4 * k: 4 * k + 4]
        assert max(block) == min(block)       # numThreads no longer changes the result: same n, same stream
        _, keep = oracle.cva_grid(1.0, n_dates, "f64")
        closed = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, keep)
        assert abs(block[0] - closed) < 4 * 0.138 / np.sqrt(131072)   # per-path CVA sd ~0.138


def run_pure(name, precision, stdin=""):
    exe = REF / f"{name}_{precision}_pure"
    if not exe.exists():
        pytest.skip(f"{exe.name} not built (needs /root/reference at build time)")
    res = subprocess.run([str(exe)], input=stdin, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def test_reference_drivers_with_no_reference_object(engine, oracle):
    """The reference's driver sources linked against libmcb200_* + libmcb200_hostapi_* only: every symbol
    they import (dev_*, host_*, Chol, printOption, ...) is ours."""
    out = run_pure("vanillaOpt", "dp", "8\n")
    bs = floats_after(out, "Prezzo Black & Scholes:", 1)[0]
    cpu = floats_after(out, "Expected price, I.C., time", 2)
    gpu = re.search(r"Simulated price for the option with GPU:.*?\n128 \n([-\d.]+) \n([-\d.]+) \n", out, re.S)
    assert bs == pytest.approx(10.386271, abs=1.5e-5)                  # driver's r = 0.048790
    assert abs(cpu[0] - bs) < 4 * cpu[1] / 1.96 and abs(float(gpu.group(1)) - bs) < 4 * float(gpu.group(2)) / 1.96
    # same Philox stream on both sides of the driver: CPU and GPU estimates agree far inside Monte Carlo error
    assert abs(cpu[0] - float(gpu.group(1))) < 2e-6
    out = run_pure("basketOpt", "dp", "8\n")
    cpu = floats_after(out, "Expected price, I.C., time", 2)
    gpu = re.search(r"Simulated price for the option with GPU:.*?\n128 \n([-\d.]+) \n([-\d.]+) \n([-\d.]+) \n", out, re.S)
    # our host estimator keeps the volatility (the reference DP host prints ~65 here, Q1): |GPU - CPU| ~ 0
    assert float(gpu.group(3)) < 2e-6 and abs(cpu[0] - float(gpu.group(1))) < 2e-6
    out = run_pure("cvaOpt", "sp")
    assert len(re.findall(r"CVA: \n([-\d.]+)", out)) == 20
