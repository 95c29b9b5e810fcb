"""The non-interactive drivers (csrc/apps/*.c): argv instead of scanf, the same
API (MonteCarlo.h host_* / dev_*) and the same printed fields as the reference drivers."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LIB = Path(__file__).resolve().parents[1] / "montecarlocuda_b200" / "lib"


def run(exe, *args):
    res = subprocess.run([str(LIB / exe), *map(str, args)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    fields = {}
    for line in res.stdout.splitlines():
        parts = line.split()
        if len(parts) == 2:
            try:
                fields.setdefault(parts[0], []).append(float(parts[1]))
            except ValueError:
                pass
    return fields, res.stdout


@pytest.mark.parametrize("precision", ["dp", "sp"])
def test_vanilla_cli(engine, precision):
    f, out = run(f"mcb200_vanillaOpt_{precision}", "--sims", 1 << 22, "--cpu-sims", 1 << 18)
    assert "Underlying asset price" in out                           # printOption, as the reference driver
    bs, gpu, conf = f["black_scholes_price"][0], f["gpu_price"][0], f["gpu_confidence"][0]
    assert bs == pytest.approx(10.386271, abs=1.5e-5)    # Hastings cnd: 10.386262 vs exact 10.386271 (Q11)
    assert abs(gpu - bs) < 4 * conf / 1.96 and f["gpu_difference_from_bs"][0] == pytest.approx(abs(gpu - bs), abs=2e-6)
    assert abs(f["cpu_price"][0] - bs) < 4 * f["cpu_confidence"][0] / 1.96
    assert f["gpu_sims"][0] == 1 << 22 and f["speedup_per_path"][0] > 10


def test_basket_cli_widths(engine):
    f3, _ = run("mcb200_basketOpt_dp", "--sims", 1 << 20, "--cpu-sims", 1 << 16)
    assert abs(f3["gpu_price"][0] - f3["cpu_price"][0]) < 4 * np.hypot(f3["gpu_confidence"][0], f3["cpu_confidence"][0]) / 1.96
    f10, _ = run("mcb200_basketOpt_dp_n10", "--sims", 1 << 22, "--cpu-sims", 1 << 16)
    assert abs(f10["gpu_price"][0] - 8.6305) < 4 * 0.0056 + 4 * f10["gpu_confidence"][0] / 1.96      # NumPy fp64 MC anchor, SURVEY 8(c)
    f64, _ = run("mcb200_basketOpt_sp_n64", "--sims", 1 << 22, "--no-cpu")
    assert abs(f64["gpu_price"][0] - 8.1237) < 4 * 0.0102 + 4 * f64["gpu_confidence"][0] / 1.96
    _, out = run("mcb200_basketOpt_dp", "--reference-data", "--sims", 1 << 16, "--no-cpu")
    assert "not positive definite" in out                            # the reference driver's own matrix is singular (Q9)


def test_cva_cli_sweep(engine, oracle):
    f, _ = run("mcb200_cvaOpt_dp", "--sims", 1 << 20, "--grids", "25,50,75", "--cpu", "--cpu-sims", 1 << 14)
    assert f["exposure_dates"] == [25.0, 50.0, 75.0]
    for n_dates, gpu, conf, cpu, cpu_conf in zip((25, 50, 75), f["gpu_cva"], f["gpu_confidence"], f["cpu_cva"], f["cpu_confidence"]):
        _, keep = oracle.cva_grid(1.0, n_dates, "f64")
        closed = oracle.cva_closed_form(100, 100, 0.05, 0.2, 1.0, 0.03, 0.6, n_dates, keep)
        assert abs(gpu - closed) < 4 * conf / 1.96 + 2e-6
        assert abs(cpu - closed) < 4 * cpu_conf / 1.96 + 2e-6


def test_cli_multi_gpu_flag_is_bit_identical(engine):
    from montecarlocuda_b200 import _lib
    if _lib.load().mcb200_device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    one, _ = run("mcb200_vanillaOpt_dp", "--sims", 1 << 24, "--no-cpu")
    two, _ = run("mcb200_vanillaOpt_dp", "--sims", 1 << 24, "--no-cpu", "--gpus", 2)
    assert one["gpu_price"] == two["gpu_price"] and one["gpu_confidence"] == two["gpu_confidence"]
