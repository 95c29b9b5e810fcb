"""CPU checks of bench.py's bookkeeping: every BASELINE workload has a committed ncu summary behind
roofline.traffic / roofline.pipe_active, and the library override fails loudly on a wrong path."""
import importlib
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_roofline_evidence_is_committed():
    bench = importlib.import_module("bench")
    baseline = [k for k in bench.WORKLOADS if k not in bench.EXTRA_WORKLOADS]
    assert sorted(baseline) == sorted(bench.NCU_SUMMARY)
    for name in baseline:
        assert (ROOT / bench.NCU_SUMMARY[name]).exists(), name
        r = bench.roofline(bench.WORKLOADS[name], 1e11, {"sm_max_mhz": 1965.0, "sm_mhz": 1965.0}, name)
        assert r["traffic"] is not None and r["traffic"] < 1e7          # no HBM-resident data on this path
        assert r["pipe_active"] and max(r["pipe_active"].values()) > 0.5
        assert 0 < r["frac"] and r["unit"] == "Ginstr/s" and r["peak"] > 0
    # the other precision of each config is described, not part of the default `also`
    for name, w in bench.EXTRA_WORKLOADS.items():
        assert "work_note" in w and name in bench.WORKLOADS


def test_library_override_fails_loudly(tmp_path, monkeypatch):
    from montecarlocuda_b200 import _lib
    monkeypatch.setenv("MCB200_LIBRARY", str(tmp_path / "no_such_libmcb200.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    assert _lib.library_path() == tmp_path / "no_such_libmcb200.so"
    with pytest.raises(_lib.Mcb200Error):
        _lib.load()
    monkeypatch.delenv("MCB200_LIBRARY")
    assert _lib.library_path().name == "libmcb200.so" and _lib.library_path().parent.name == "lib"
