"""Generate tests/golden/*.json from the UNMODIFIED reference compiled into oracle/_ref/.

Run in the build container (where /root/reference is mounted):
    make -C oracle ref && python tests/golden/make_golden.py
The reference ships no golden vectors; these fixtures are its own outputs, recorded
once so that the GPU box (which has neither the sources nor, necessarily, the same libm) can pin
the oracle against them.  Deterministic functions (host_bsCall, Chol) are stored exactly (Python
float repr round-trips); the Monte Carlo estimators are stored with the pinned srand() seed.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle_lib import Reference  # noqa: E402


def equicorr(n, rho=0.3):
    c = np.full((n, n), rho)
    np.fill_diagonal(c, 1.0)
    return c


def ref_generator_rho(n):
    """getRandomRho of the reference driver (DP/basketOpt.cu:160-177): off-diagonals +-0.5 by column parity."""
    c = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            r = 1.0 if i == j else (0.5 if j % 2 == 0 else -0.5)
            c[i, j] = c[j, i] = r
    return c


def sigma_vector(n):
    """getRandomSigma (DP/basketOpt.cu:147-158): 0.3, 0.2, 0.3, ..."""
    return [0.3 if i % 2 == 0 else 0.2 for i in range(n)]


def main():
    out = {}

    # ---- Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the Random123
    # distribution; also the values SURVEY.md 8(c) recomputed) ----
    philox = [
        {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]},
        {"ctr": [0xFFFFFFFF] * 4, "key": [0xFFFFFFFF] * 2, "out": [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]},
        {"ctr": [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], "key": [0xA4093822, 0x299F31D0],
         "out": [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]},
    ]
    (HERE / "philox_kat.json").write_text(json.dumps(philox, indent=1))

    # ---- host_bsCall (DP/MonteCarloHost.c:139-143), both precisions, exact ----
    rng = np.random.default_rng(20180206)
    cases = [(100, 100, 0.05, 0.2, 1.0), (100, 100, 0.048790, 0.2, 1.0), (100, 120, 0.01, 0.4, 2.0),
             (50, 40, 0.03, 0.15, 0.25), (100, 100, 0.05, 0.2, 0.02), (100, 300, 0.05, 0.2, 1.0),
             (100, 30, 0.05, 0.2, 1.0)]
    for _ in range(40):
        cases.append((float(rng.uniform(20, 200)), float(rng.uniform(20, 200)), float(rng.uniform(0, 0.1)),
                      float(rng.uniform(0.05, 0.6)), float(rng.uniform(0.05, 3.0))))
    bs = {"dp": [], "sp": []}
    for prec in ("dp", "sp"):
        ref = Reference(prec, 3)
        for c in cases:
            args = [float(np.float32(x)) for x in c] if prec == "sp" else list(c)
            bs[prec].append({"args": args, "value": float(ref.lib.host_bsCall(ref.option(*args)))})
    (HERE / "ref_bscall.json").write_text(json.dumps(bs, indent=1))

    # ---- Chol (DP/MonteCarloHost.c:90-105), exact ----
    chol = []
    mats = {3: [("driver_default_singular", np.array([[1, -.5, -.5], [-.5, 1, -.5], [-.5, -.5, 1.0]])),
                ("equicorr_0.3", equicorr(3))],
            10: [("equicorr_0.3", equicorr(10)), ("driver_generator_indefinite", ref_generator_rho(10))],
            64: [("equicorr_0.3", equicorr(64))]}
    for n, items in mats.items():
        for prec in ("dp", "sp"):
            ref = Reference(prec, n)
            for name, c in items:
                a = ref.chol(c)
                chol.append({"n": n, "precision": prec, "name": name, "c": np.asarray(c, dtype=ref.np_real).astype(float).tolist(),
                             "a": a.astype(float).tolist()})
    (HERE / "ref_chol.json").write_text(json.dumps(chol))

    # ---- the three host Monte Carlo estimators, srand() pinned through --wrap=time ----
    mc = []
    seed = 20180206
    for prec in ("dp", "sp"):
        ref = Reference(prec, 3)
        ref.set_seed(seed)
        v = ref.lib.host_vanillaOpt(ref.option(100, 100, 0.05, 0.2, 1.0), 1 << 20)
        mc.append({"workload": "vanilla", "precision": prec, "paths": 1 << 20, "seed": seed,
                   "params": [100, 100, 0.05, 0.2, 1.0], "Expected": float(v.Expected), "Confidence": float(v.Confidence)})
        # driver default basket (DP/basketOpt.cu:34-61), factor by the reference's own Chol
        c = np.array([[1, -.5, -.5], [-.5, 1, -.5], [-.5, -.5, 1.0]])
        a = ref.chol(c)
        m = ref.multi([100] * 3, [0.2, 0.3, 0.2], a, [0] * 3, [1 / 3] * 3, 100.0, 1.0, 0.048790164)
        ref.set_seed(seed)
        v = ref.lib.host_basketOpt(m, 1 << 20)
        mc.append({"workload": "basket", "precision": prec, "n": 3, "paths": 1 << 20, "seed": seed,
                   "corr": "driver_default_singular", "Expected": float(v.Expected), "Confidence": float(v.Confidence)})
        for n_dates in (25, 50, 75):
            cva = ref.cva(0.03, 0.6, ref.option(100, 100, 0.05, 0.2, 1.0), n_dates)
            ref.set_seed(seed)
            v = ref.lib.host_cvaEquityOption(cva, 1 << 17)
            mc.append({"workload": "cva", "precision": prec, "n_dates": n_dates, "paths": 1 << 17, "seed": seed,
                       "Expected": float(v.Expected), "Confidence": float(v.Confidence)})
        for n in (10, 64):
            refn = Reference(prec, n)
            a = refn.chol(equicorr(n))
            m = refn.multi([100] * n, sigma_vector(n), a, [0] * n, [1 / n] * n, 100.0, 1.0, 0.048790164)
            paths = 1 << 19 if n == 10 else 1 << 16
            refn.set_seed(seed)
            v = refn.lib.host_basketOpt(m, paths)
            mc.append({"workload": "basket", "precision": prec, "n": n, "paths": paths, "seed": seed,
                       "corr": "equicorr_0.3", "Expected": float(v.Expected), "Confidence": float(v.Confidence)})
    (HERE / "ref_host_mc.json").write_text(json.dumps(mc, indent=1))
    print("wrote", *(p.name for p in sorted(HERE.glob("*.json"))))


if __name__ == "__main__":
    main()
