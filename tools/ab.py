#!/usr/bin/env python
"""A/B timing of library variants on one box:  python tools/ab.py [--steps K] name=path ... -- workload ...
Each (variant, workload) is one `bench.py --also none --no-cpu-baseline` run; prints ms per job."""
import json
import os
import subprocess
import sys

args = sys.argv[1:]
steps = "10"
if args and args[0] == "--steps":
    steps, args = args[1], args[2:]
split = args.index("--")
variants = [a.split("=", 1) for a in args[:split]]
workloads = args[split + 1:]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for w in workloads:
    for rep in range(2):
        for name, path in variants:
            env = dict(os.environ)
            if path != "default":
                env["MCB200_LIBRARY"] = os.path.join(root, path)
            out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", w, "--also", "none", "--no-cpu-baseline",
                                  "--steps", steps], env=env, capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(f"{w:22s} {name:10s} run {rep}: {d['ms_per_step']:.3f} ms  price {d['price']!r}  sm_mhz {d['clocks']['sm_mhz']}", flush=True)
            except Exception as exc:
                print(w, name, "FAILED", exc, out.stderr[-500:], flush=True)
