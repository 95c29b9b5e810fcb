#!/usr/bin/env python
"""Per-call latency of the blocking C-ABI entry points on small jobs (what a user of the reference's
dev_* functions sees): median wall time of one call, host structs in, price out."""
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import montecarlocuda_b200 as m  # noqa: E402

opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
cva = m.CVA(0.03, 0.6, opt, 50)
with m.Engine(0) as eng:
    for name, call in (("vanilla f64", lambda n: eng.vanilla(opt, n, "f64")), ("vanilla f32", lambda n: eng.vanilla(opt, n, "f32")),
                       ("cva50 f64", lambda n: eng.cva(cva, n, "f64"))):
        for n in (1 << 10, 1 << 16, 1 << 20, 1 << 24):
            for _ in range(20):
                call(n)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter()
                r = call(n)
                ts.append(time.perf_counter() - t0)
            print(f"{name:12s} {n:>9d} paths: median {1e6 * statistics.median(ts):8.1f} us  min {1e6 * min(ts):8.1f} us   price {r.Expected:.6f}", flush=True)
# the reference-named entry point (creates nothing per call: the shim keeps one context per process)
for n in (1 << 16, 1 << 20):
    for _ in range(10):
        m.dev_vanillaOpt(opt, 512, 128, n)
    ts = []
    for _ in range(100):
        t0 = time.perf_counter()
        r = m.dev_vanillaOpt(opt, 512, 128, n)
        ts.append(time.perf_counter() - t0)
    print(f"dev_vanillaOpt {n:>9d} sims: median {1e6 * statistics.median(ts):8.1f} us  min {1e6 * min(ts):8.1f} us", flush=True)
