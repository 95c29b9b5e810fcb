"""CPU checks of bench.py's bookkeeping: the roofline is EXECUTED work on the binding pipe against the measured pipe
rate, taken from an ncu capture that is tied to the loaded library by the sha256 of the kernel's SASS -- and refused
when they differ; the library override fails loudly on a wrong path."""
import importlib
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CLOCKS = {"sm_max_mhz": 1965.0, "sm_mhz": 1965.0}


@pytest.fixture()
def bench(ensure_built, monkeypatch):
    b = importlib.import_module("bench")
    monkeypatch.setattr(b, "_manifest_cache", {})
    return b


def test_build_manifest_describes_the_library(ensure_built, bench):
    manifest, why = bench.loaded_manifest()
    assert manifest is not None, why
    kernels = manifest["kernel_sass_sha256"]
    for k in ("mc_accumulate_kernel<Vanilla<double,2,1,1>>", "mc_accumulate_kernel<Vanilla<float,8,1,1>>", "mc_accumulate_kernel<Basket<double,10,0,1>>",
              "mc_accumulate_kernel<Cva<double,1,0>>", "basket_tc_accumulate_kernel<0>", "mc_accumulate_batch_kernel<Cva<double,1,0>>"):
        assert len(kernels[k]) == 64, k
    from montecarlocuda_b200 import build
    assert manifest["source_sha256"] == build.source_hash()      # the library in the tree was built from the sources in the tree


def test_roofline_is_executed_work_tied_to_the_loaded_sass(bench):
    table = json.loads(bench.KERNEL_WORK.read_text())
    baseline = [k for k in bench.WORKLOADS if k not in bench.EXTRA_WORKLOADS and k not in bench.SMALL_WORKLOADS]
    assert sorted(baseline) == sorted(table)
    for name in baseline:
        entry = table[name]
        assert (ROOT / entry["capture"]).exists(), name
        # units/s at which that capture ran: the fraction must then reproduce ncu's own pipe percentage
        rate = entry["units_per_launch"] / (entry["duration_ms"] * 1e-3)
        r = bench.roofline(bench.WORKLOADS[name], rate, CLOCKS, name)
        assert r["frac"] is not None, r.get("frac_unavailable")
        assert 0.3 < r["frac"] <= 1.0 and r["unit"] == "Ginstr/s" and r["peak"] > 0
        pipe = {"fp64-pipe": "fp64", "mufu-pipe": "xu_mufu"}[r["bound"]]
        assert abs(r["frac"] - r["pipe_active"][pipe]) < 0.05, (name, r["frac"], r["pipe_active"])
        assert r["traffic"] is not None and r["traffic"] < 1e7          # no HBM-resident data on this path
        assert r["work_per_unit_executed"] < bench.WORKLOADS[name]["work"] or r["bound"] == "mufu-pipe"
        assert r["canonical"]["frac_canonical"] > 0                     # SURVEY's yardstick stays next to it
    # the other precision of each config is described, not part of the default `also`
    for name, w in bench.EXTRA_WORKLOADS.items():
        assert "work_note" in w and name in bench.WORKLOADS


def test_counters_of_another_build_are_refused(bench, monkeypatch, tmp_path):
    table = json.loads(bench.KERNEL_WORK.read_text())
    table["vanilla_f64_2p32"]["sass_sha256"] = "0" * 64          # a capture of some other kernel binary
    stale = tmp_path / "kernel_work.json"
    stale.write_text(json.dumps(table))
    monkeypatch.setattr(bench, "KERNEL_WORK", stale)
    r = bench.roofline(bench.WORKLOADS["vanilla_f64_2p32"], 4.4e11, CLOCKS, "vanilla_f64_2p32")
    assert r["frac"] is None and r["achieved"] is None and r["work_per_unit_executed"] is None and "pipe_active" not in r
    assert "stale ncu capture" in r["frac_unavailable"]
    assert r["canonical"]["frac_canonical"] == pytest.approx(4.4e11 * 57 / (64 * 148 * 1965e6))
    ok = bench.roofline(bench.WORKLOADS["cva50_f64_2p26"], 2e11, CLOCKS, "cva50_f64_2p26")     # untouched entries still count
    assert ok["frac"] is not None
    # ... and a library without (or with somebody else's) manifest gets no counters at all
    monkeypatch.setattr(bench, "_manifest_cache", {})
    monkeypatch.setattr(bench, "_sha256", lambda path: "f" * 64)
    r = bench.roofline(bench.WORKLOADS["cva50_f64_2p26"], 2e11, CLOCKS, "cva50_f64_2p26")
    assert r["frac"] is None and "another binary" in r["frac_unavailable"]


def test_library_override_fails_loudly(tmp_path, monkeypatch):
    from montecarlocuda_b200 import _lib
    monkeypatch.setenv("MCB200_LIBRARY", str(tmp_path / "no_such_libmcb200.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    assert _lib.library_path() == tmp_path / "no_such_libmcb200.so"
    with pytest.raises(_lib.Mcb200Error):
        _lib.load()
    monkeypatch.delenv("MCB200_LIBRARY")
    assert _lib.library_path().name == "libmcb200.so" and _lib.library_path().parent.name == "lib"
