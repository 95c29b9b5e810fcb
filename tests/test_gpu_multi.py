"""Real multi-GPU checks (need >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
Single-GPU boxes skip them; the partition-independence property itself is covered on one GPU by
test_virtual_ranks_bit_identical and on the CPU by test_distributed_gloo.py."""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import montecarlocuda_b200 as m
from montecarlocuda_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def device_count():
    return _lib.load().mcb200_device_count()


def test_single_process_multi_device_bit_identical(engine):
    n_dev = device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 CUDA devices")
    lib = _lib.load()
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    one = engine.vanilla(opt, (1 << 26) + 777, "f64", 11)
    engines = [m.Engine(d) for d in range(n_dev)]
    for use in range(2, n_dev + 1):
        handles = (C.c_void_p * use)(*[e.handle for e in engines[:use]])
        r = _lib.ResultT()
        c = opt._c()
        _lib.check(lib.mcb200_vanilla_multi(handles, use, _lib.F64, C.byref(c), (1 << 26) + 777, 11, C.byref(r)))
        assert (r.expected, r.confidence, r.sum, r.sumsq, r.n_paths) == (one.Expected, one.Confidence, one.sum, one.sumsq, one.n_paths)
    for e in engines:
        e.close()


def test_torchrun_ranks_bit_identical(engine):
    n_dev = device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 CUDA devices")
    prices = {}
    for world in sorted({1, 2, n_dev}):
        cmd = [sys.executable, str(ROOT / "bench.py"), "--gpus", str(world), "--steps", "2", "--warmup", "3", "--also", "cva50_f64_2p26",
               "--no-cpu-baseline"]
        if world > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29700 + world)] + cmd[1:]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
        assert out.returncode == 0, out.stderr[-3000:]
        line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        prices[world] = (line["price"], line["std_error"], line["also"]["cva50_f64_2p26"]["price"])
        assert line["n_gpus"] == world
    for world in prices:
        assert prices[world] == prices[1]          # bit-identical output across GPU counts (north_star)
