"""Kernel time of baskets beyond the register templates (the wide route of kernels_basket.cu) next to the 64-asset kernels."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import montecarlocuda_b200 as m
from oracle_lib import Oracle
from test_gpu_parity import make_basket
o = Oracle()
with m.Engine(0) as eng:
    eng.set_timing(True)
    for n, prec, paths in ((100, "f64", 1 << 22), (100, "f32", 1 << 22), (256, "f64", 1 << 20), (256, "f32", 1 << 20), (64, "f64", 1 << 22), (64, "f32", 1 << 22)):
        opt = make_basket(o, n, prec)
        for _ in range(2):
            r = eng.basket(opt, paths, prec)
        print(f"basket n={n} {prec} {paths} paths: {r.kernel_ms:.3f} ms  {paths / r.kernel_ms * 1e3:.3e} paths/s  price {r.Expected:.6f} +- {r.std_error:.6f}", flush=True)
