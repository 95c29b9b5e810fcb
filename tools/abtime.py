#!/usr/bin/env python
"""A/B kernel timing of library variants on one box, without torch:
    python tools/abtime.py [--reps K] name=path ... -- workload ...
One subprocess per variant (MCB200_LIBRARY picks the build), every workload in it: 3 warm-up calls, K timed calls of the
blocking C-ABI entry point, kernel time from the engine's own CUDA events (mcb200_set_timing).  Prints min / median
kernel ms and the price (variants of one row must agree bit for bit unless they change the arithmetic)."""
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(workloads, reps):
    import bench  # workload table and synthetic parameters only
    import montecarlocuda_b200 as m

    out = {}
    with m.Engine(0) as eng:
        if hasattr(eng, "set_timing"):
            eng.set_timing(True)
        for name in workloads:
            w = bench.WORKLOADS[name]
            params = bench.make_params(w)
            call = getattr(eng, w["kind"])
            for _ in range(3):
                r = call(params, w["paths"], w["prec"])
            ms = []
            for _ in range(reps):
                r = call(params, w["paths"], w["prec"])
                ms.append(r.kernel_ms)
            out[name] = {"min": min(ms), "median": statistics.median(ms), "price": r.Expected}
    print("ABTIME " + json.dumps(out), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(args[2:], int(args[1]))
    reps = 10
    if args and args[0] == "--reps":
        reps, args = int(args[1]), args[2:]
    split = args.index("--")
    variants = [a.split("=", 1) for a in args[:split]]
    workloads = args[split + 1:]
    results = {}
    for name, path in variants:
        env = dict(os.environ)
        if path != "default":
            env["MCB200_LIBRARY"] = str(ROOT / path)
        res = subprocess.run([sys.executable, __file__, "--child", str(reps)] + workloads, env=env, capture_output=True, text=True)
        line = [l for l in res.stdout.splitlines() if l.startswith("ABTIME ")]
        if not line:
            print(name, "FAILED", res.stderr[-800:], flush=True)
            continue
        results[name] = json.loads(line[-1][7:])
    for w in workloads:
        for name, _ in variants:
            if name in results and w in results[name]:
                r = results[name][w]
                print(f"{w:20s} {name:14s} min {r['min']:9.4f} ms  median {r['median']:9.4f} ms  price {r['price']!r}", flush=True)


if __name__ == "__main__":
    main()
