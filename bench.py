#!/usr/bin/env python
"""bench.py -- the Monte Carlo pricing hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pricing job of the workload (BASELINE.json configs), sharded over the N ranks:
    ONE kernel over this rank's chunk range (no memset: the launch's last CTA leaves its control block zeroed), whose
    last CTA pushes the rank's integer limbs into every peer's mailbox over NVLink (N > 1; split phase: the sum is
    taken out of the mailbox when the result is read; --peer-mode wait: the last CTA also waits and sums;
    --combine nccl: ONE int64 all-reduce after the kernel instead).
Headline workload: European call, 2^32 paths, fp64 (BASELINE.json configs[1]); the fp32 run of the
same config and the other configs are reported under "also" (`--also none` to skip them).

Keys beyond the driver contract:
  roofline     binding pipe of the kernel (fp64 / MUFU), achieved = units/s x thread instructions the kernel EXECUTES
               on that pipe per unit (profiles/kernel_work.json: one `ncu --set full` launch per workload, tied to the
               loaded library by the sha256 of the kernel's SASS -- null with a reason when they differ), peak = the
               MEASURED pipe rate (profiles/r01_pipe_peaks.json); frac_canonical keeps SURVEY.md 8(d)'s yardstick
  cpu_baseline the reference's own MonteCarloHost.c (oracle/_ref, gcc -O2), 1 core, bounded sample
  gpu_baseline the reference's own GPU kernels rebuilt for sm_100a (oracle/_ref), same box, its launch shapes
  e2e          the same job through the blocking C-ABI call a reference user makes
               (host structs in, OptionValue out: one kernel launch, result read from mapped host memory, closing)
  limbs        the 12 combined accumulator words (hex): identical for every GPU count, or bit-identity is broken
`--impl reference` times the reference CPU path fanned out over every host core instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SM_COUNT = 148

# ---- workloads (BASELINE.json configs; synthetic parameters of SURVEY.md 8(d)) -----------------
# canonical per-unit work on the binding pipe; unit = path, or path-step for CVA
WORKLOADS = {
    "vanilla_f64_2p32": dict(kind="vanilla", prec="f64", paths=1 << 32, bound="fp64", work=57.0, units_per_path=1),
    "vanilla_f32_2p32": dict(kind="vanilla", prec="f32", paths=1 << 32, bound="mufu", work=3.0, units_per_path=1),
    "basket10_f64_2p28": dict(kind="basket", n=10, prec="f64", paths=1 << 28, bound="fp64", work=589.0, units_per_path=1),
    "cva50_f64_2p26": dict(kind="cva", dates=50, prec="f64", paths=1 << 26, bound="fp64", work=133.0, units_per_path=50),
    "basket64_f32_2p30": dict(kind="basket", n=64, prec="f32", paths=1 << 30, bound="issue", work=3236.0, units_per_path=1),
}
# BASELINE config 1 (the reference's CPU-runnable case, 2^20 paths): a 3 us kernel -- what it measures is the call
SMALL_WORKLOADS = {
    "vanilla_f64_2p20": dict(kind="vanilla", prec="f64", paths=1 << 20, bound="fp64", work=57.0, units_per_path=1),
}
WORKLOADS.update(SMALL_WORKLOADS)
# the other precision of each config (the reference ships every workload in both): `--also everything`; their
# per-unit work is derived here by the same rule as SURVEY.md 8(d) (not SURVEY figures)
EXTRA_WORKLOADS = {
    "basket10_f32_2p28": dict(kind="basket", n=10, prec="f32", paths=1 << 28, bound="mufu", work=30.0, units_per_path=1,
                              work_note="10 x (lg2, sqrt, sin, cos)/2 + 10 ex2"),
    "cva50_f32_2p26": dict(kind="cva", dates=50, prec="f32", paths=1 << 26, bound="mufu", work=8.0, units_per_path=50,
                           work_note="per path-step: normal 2, spot ex2, pdf ex2, 2 x cnd (rcp + ex2 shared -> 2 rcp), 2 spare"),
    "basket64_f64_2p30": dict(kind="basket", n=64, prec="f64", paths=1 << 30, bound="fp64", work=5476.0, units_per_path=1,
                              work_note="64 x 34 normal + 2080 triangular FMA + 64 x 18 exp + 64 + 4"),
}
WORKLOADS.update(EXTRA_WORKLOADS)
PIPE_PER_CLK_PER_SM = {"fp64": 64.0, "mufu": 16.0, "issue": 128.0}
# What ONE launch of each workload's kernel executed (tools/ncu_summary.py --work, from an `ncu --set full` capture at
# the full path count) and the sha256 of that kernel's SASS.  The library bench.py loads carries a build manifest with
# the same hashes (montecarlocuda_b200/build.py): counters of another build are refused, not quoted.
KERNEL_WORK = ROOT / "profiles" / "kernel_work.json"
PIPE_PEAKS = ROOT / "profiles" / "r01_pipe_peaks.json"   # measured on this pool's B200 (tools/microbench/pipe_peaks.cu)
_MEASURED_PIPE = {"fp64": "dfma", "xu_mufu": "mufu_ex2", "fma": "ffma", "alu": "lop3"}


def _sha256(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest()


_manifest_cache = {}


def loaded_manifest():
    """(manifest, None) of the library this process prices with, or (None, why not)."""
    import montecarlocuda_b200 as m

    lib = Path(m.library_path())
    if lib in _manifest_cache:
        return _manifest_cache[lib]
    try:
        manifest = json.loads(lib.with_name("libmcb200.manifest.json").read_text())
        if manifest.get("library_sha256") != _sha256(lib):
            out = (None, f"{lib.name}: the build manifest next to it describes another binary")
        else:
            out = (manifest, None)
    except (OSError, ValueError) as exc:
        out = (None, f"no build manifest next to {lib.name} ({exc})")
    _manifest_cache[lib] = out
    return out


def kernel_work(name):
    """(entry, None) of profiles/kernel_work.json when it was captured from the SASS this process runs, else (None, reason)."""
    try:
        entry = json.loads(KERNEL_WORK.read_text()).get(name)
    except (OSError, ValueError):
        entry = None
    if not entry:
        return None, "no ncu capture of this workload in profiles/kernel_work.json"
    manifest, why = loaded_manifest()
    if manifest is None:
        return None, why
    have = manifest.get("kernel_sass_sha256", {}).get(entry["kernel"])
    if have != entry["sass_sha256"]:
        return None, (f"stale ncu capture: {entry['kernel']} had SASS {entry['sass_sha256'][:12]} when {entry.get('capture')} was taken, "
                      f"the loaded library has {str(have)[:12]}")
    return entry, None


HEADLINE = "vanilla_f64_2p32"


def cholesky_reference_algorithm(c):
    """Column Cholesky as the reference computes it (MonteCarloHost.c:90-105) -- input preparation
    for the basket workloads (the factor is an INPUT of dev_basketOpt), not part of the timed path."""
    import numpy as np

    n = c.shape[0]
    a = np.zeros_like(c)
    v = np.zeros(n)
    for j in range(n):
        for i in range(j, n):
            v[i] = c[i, j] - a[j, :j] @ a[i, :j]
        if v[j] > 0:
            a[j:, j] = v[j:] / math.sqrt(v[j])
    return a


def make_params(w):
    import numpy as np
    import montecarlocuda_b200 as m

    if w["kind"] == "vanilla":
        return m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    if w["kind"] == "cva":
        return m.CVA(0.03, 0.6, m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0), w["dates"])
    n = w["n"]
    c = np.full((n, n), 0.3)
    np.fill_diagonal(c, 1.0)
    vol = [0.3 if i % 2 == 0 else 0.2 for i in range(n)]
    return m.MultiOptionData([100.0] * n, vol, cholesky_reference_algorithm(c), [0.0] * n, [1.0 / n] * n, 100.0, 1.0, 0.048790164)


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons, power = [], None, set(), []
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                inside = t0 - 0.05 <= t <= t1 + 0.15
                if inside:
                    mhz.append(float(parts[1]))
                    power.append(float(parts[3]))
                mx = float(parts[2])
                if inside:
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz), "power_w_max": max(power) if power else None}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return json.loads(p.read_text())
    except (OSError, ValueError):
        return {}


# ---- the reference CPU path (oracle/_ref) ------------------------------------------------------
def _ref_worker(args):
    kind, precision, n_assets, n_dates, paths, seed = args
    from oracle_lib import Reference
    import numpy as np

    devnull = os.open(os.devnull, os.O_WRONLY)
    ref = Reference(precision, n_assets)
    ref.set_seed(seed)
    t0 = time.perf_counter()
    if kind == "vanilla":
        v = ref.lib.host_vanillaOpt(ref.option(100, 100, 0.05, 0.2, 1.0), paths)
    elif kind == "cva":
        cva = ref.cva(0.03, 0.6, ref.option(100, 100, 0.05, 0.2, 1.0), n_dates)
        v = ref.lib.host_cvaEquityOption(cva, paths)
    else:
        c = np.full((n_assets, n_assets), 0.3)
        np.fill_diagonal(c, 1.0)
        a = ref.chol(c)
        vol = [0.3 if i % 2 == 0 else 0.2 for i in range(n_assets)]
        mo = ref.multi([100] * n_assets, vol, a, [0] * n_assets, [1 / n_assets] * n_assets, 100.0, 1.0, 0.048790164)
        v = ref.lib.host_basketOpt(mo, paths)
    dt = time.perf_counter() - t0
    os.close(devnull)
    return dt, float(v.Expected), float(v.Confidence)


def ref_args(w, paths, seed):
    precision = "dp" if w["prec"] == "f64" else "sp"
    return (w["kind"], precision, w.get("n", 3), w.get("dates", 0), int(paths), seed)


def cpu_baseline_one_core(w, seconds_target=12.0):
    """Reference MonteCarloHost.c, 1 core, bounded sample of the same workload."""
    rate_guess = {"vanilla": 1.1e7, "cva": 5.8e6 / max(w.get("dates", 1), 1), "basket": 1.1e7 / (w.get("n", 3) ** 1.4)}[w["kind"]]
    paths = int(min(max(rate_guess * seconds_target, 1 << 14), (1 << 27)))
    dt, expected, conf = _ref_worker(ref_args(w, paths, 20180206))
    units = paths * w["units_per_path"]
    return {"value": units / dt, "unit": unit_name(w), "cores": 1, "kind": "reference",
            "sample": f"{paths} paths of {describe(w)} through host_{w['kind']} (MonteCarloHost.c, gcc -O2, rand() Box-Muller), {dt:.2f} s",
            "price": expected, "confidence": conf}


def unit_name(w):
    return "path-steps/s" if w["kind"] == "cva" else "paths/s"


def describe(w):
    if w["kind"] == "vanilla":
        return f"European call S0=100 K=100 r=0.05 sigma=0.2 T=1, {w['prec']}"
    if w["kind"] == "cva":
        return f"CVA of a call, {w['dates']} exposure dates, lambda=0.03 LGD=0.6, {w['prec']}"
    return f"basket call, {w['n']} underlyings, equicorrelation 0.3 (Cholesky), {w['prec']}"


def run_reference_arm(args, w, name):
    """--impl reference: the reference CPU estimator on every host core (one process per core, its
    rand() generator is a process-global), same metric/config; under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    rate = {"vanilla": 1.1e7, "cva": 5.8e6 / max(w.get("dates", 1), 1), "basket": 1.1e7 / (w.get("n", 3) ** 1.4)}[w["kind"]]
    per_core = int(min(max(rate * 2.0, 1 << 12), 1 << 26))  # ~2 s of work per core per step
    # dlopen the reference build HERE, before the pool forks: the workers inherit the mapping, and the record of which
    # native libraries this process loaded shows what the arm actually ran
    from oracle_lib import Reference
    Reference(ref_args(w, 1, 0)[1], ref_args(w, 1, 0)[2])
    ctx = mp.get_context("fork")
    times, price = [], None
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [ref_args(w, per_core, 1000 * step + c) for c in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
            price = sum(r[1] for r in res) / len(res)
    total = sum(times)
    units = per_core * cores * w["units_per_path"] * args.steps
    value = units / total
    line = {
        "impl": "reference", "metric": f"{unit_name(w)} ({name})", "value": value, "unit": unit_name(w), "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": w["prec"], "data": "synthetic",
        "config": {"workload": name, "description": describe(w), "paths_per_step": per_core * cores,
                   "note": "bounded sample of the workload: the reference CPU path is ~1e7 paths/s/core"},
        "cpu_baseline": {"value": value, "unit": unit_name(w), "cores": cores, "kind": "reference",
                         "sample": f"{per_core} paths x {cores} processes per step, MonteCarloHost.c gcc -O2 from oracle/_ref"},
        "e2e": {"value": value, "unit": unit_name(w), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "price": price,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- our arm -----------------------------------------------------------------------------------
def time_workload(name, w, pricer, dist, torch, rank, world, steps, warmup, sample_clocks=False, gpu_index=0):
    params = make_params(w)
    prec = w["prec"]
    seed = 0x6D63623230300001

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        plan = pricer.enqueue(w["kind"], params, w["paths"], prec, seed)
    pricer.result(plan)
    barrier()
    sampler = ClockSampler(gpu_index) if sample_clocks else None
    if sampler:
        sampler.start()
    time.sleep(0.25)   # every rank alike: the clock sampler of rank 0 gets its first lines, nobody spins in a barrier meanwhile
    launches0 = pricer.engine.launch_count
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    start.record()
    for _ in range(steps):
        plan = pricer.enqueue(w["kind"], params, w["paths"], prec, seed)
    stop.record()
    barrier()
    t1 = time.perf_counter()
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=pricer.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = pricer.engine.launch_count - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    result = pricer.result(plan)
    limbs = [f"{int(x) & 0xFFFFFFFFFFFFFFFF:x}" for x in pricer.host.tolist()]

    # end to end through the blocking call a user of the reference API makes
    e2e_times = []
    for i in range(steps + 1):
        barrier()
        t = time.perf_counter()
        if world == 1:
            r = getattr(pricer.engine, w["kind"])(params, w["paths"], prec, seed)
        else:
            r = pricer.price(w["kind"], params, w["paths"], prec, seed)
        e = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=pricer.device)
        if world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
        if i > 0:
            e2e_times.append(float(e.item()))
        assert r.Expected == result.Expected, "e2e call and sharded call disagree"
    units = w["paths"] * w["units_per_path"]
    return dict(ms_total=ms, ms_per_step=ms / steps, value=units * steps / (ms * 1e-3), e2e_value=units * len(e2e_times) / sum(e2e_times),
                e2e_ms=1e3 * sum(e2e_times) / len(e2e_times), launches=launches * world, clocks=clocks, result=result,
                params_bytes=param_bytes(w), limbs=limbs)


def param_bytes(w):
    if w["kind"] == "vanilla":
        return 40
    if w["kind"] == "cva":
        return 72
    n = w["n"]
    return 8 * (4 * n + n * n + 3)


def roofline(w, value, clocks, name=None):
    """value: units/s of ONE GPU.  The path is compute-bound (96 bytes out, <= 33 KB of parameters in): the roof is a
    pipe rate, not HBM and not the tensor cores."""
    peaks = measured_peaks()
    f_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    out = {"unit": "Ginstr/s", "traffic": None}
    # ---- executed work on the binding pipe against the MEASURED pipe rate ----
    entry, why = kernel_work(name)
    try:
        pipe_peaks = json.loads(PIPE_PEAKS.read_text())
    except (OSError, ValueError):
        pipe_peaks = {}
    if entry:
        pct = {k: v for k, v in entry["pipe_pct"].items() if v is not None}
        math = {k: pct[k] for k in ("fp64", "xu_mufu", "fma", "alu", "tensor") if k in pct}
        bound = max(math, key=math.get)
        per_unit = entry["thread_inst_per_unit"].get(bound)
        peak = (pipe_peaks.get(_MEASURED_PIPE.get(bound, ""), {}) or {}).get("gops")
        out.update({"bound": {"fp64": "fp64-pipe", "xu_mufu": "mufu-pipe", "fma": "fma-pipe", "alu": "alu-pipe", "tensor": "tensor-pipe"}[bound],
                    "work_per_unit_executed": per_unit, "inst_per_unit_executed": entry["thread_inst_per_unit"].get("total"),
                    "work_source": f"profiles/kernel_work.json <- {entry.get('capture')}: thread instructions ONE launch of {entry['kernel']} executed on the {bound} pipe "
                                   f"(ncu sm__inst_executed_pipe_*.sum x 32 / units), SASS sha256 {entry['sass_sha256'][:16]} = the loaded library's",
                    "pipe_active": dict(sorted(((k, round(v / 100.0, 4)) for k, v in pct.items() if v >= 1.0), key=lambda kv: -kv[1])),
                    "pipe_active_source": f"{entry.get('capture')} (ncu --set full, one launch at the full path count; same SASS as the loaded library)",
                    "traffic": entry.get("dram_bytes"),
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of that launch (instruction and constant fetch: the path has no HBM-resident data)"})
        if per_unit and peak:
            achieved = value * per_unit
            out.update({"achieved": achieved / 1e9, "peak": peak, "frac": achieved / (peak * 1e9),
                        "peak_source": f"measured {_MEASURED_PIPE[bound].upper()} rate of this pool's B200, profiles/r01_pipe_peaks.json "
                                       f"({pipe_peaks[_MEASURED_PIPE[bound]]['per_clk_per_sm_at_max_clock']:.1f} thread-instr/clk/SM x 148 SMs x 1965 MHz); no HBM or tensor roof applies"})
        else:
            out.update({"achieved": None, "peak": peak, "frac": None, "frac_unavailable": f"no measured peak for the {bound} pipe"})
    else:
        out.update({"bound": {"fp64": "fp64-pipe", "mufu": "mufu-pipe", "issue": "issue-slots"}[w["bound"]], "achieved": None, "peak": None, "frac": None,
                    "work_per_unit_executed": None, "frac_unavailable": why})
    # ---- the canonical yardstick of SURVEY.md 8(d), kept beside it (can exceed 1: our kernels need fewer instructions) ----
    pipe = PIPE_PER_CLK_PER_SM[w["bound"]]
    canon_peak = pipe * SM_COUNT * f_max * 1e6
    out["canonical"] = {"bound": {"fp64": "fp64-pipe", "mufu": "mufu-pipe", "issue": "issue-slots"}[w["bound"]], "work_per_unit": w["work"],
                        "achieved": value * w["work"] / 1e9, "peak": canon_peak / 1e9, "frac_canonical": value * w["work"] / canon_peak,
                        "work_source": ("derived by the rule of SURVEY.md 8(d): " + w["work_note"]) if "work_note" in w else
                                       "SURVEY.md 8(d) canonical algorithm (Philox4x32-10 + Box-Muller + libdevice-cost transcendentals)",
                        "peak_source": f"{pipe:g} thread-instr/clk/SM x {SM_COUNT} SMs x clocks.max.sm {f_max:g} MHz"}
    return out


# ---- the reference's own GPU kernels on the same box (oracle/_ref, rebuilt for sm_100a) -------------
class _CaptureStdout:
    """The reference prints its timing lines with printf: catch file descriptor 1."""

    def __enter__(self):
        import tempfile
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.tmp = tempfile.TemporaryFile(mode="w+b")
        os.dup2(self.tmp.fileno(), 1)
        return self

    def __exit__(self, *exc):
        import ctypes
        try:
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.tmp.seek(0)
        self.text = self.tmp.read().decode(errors="replace")
        self.tmp.close()


def _ref_gpu_call(w, sims, blocks, threads):
    """One call of the reference's dev_* entry point: wall time of the call, its own 'Kernel done in ms' / 'RNG done in ms'."""
    import re
    import numpy as np
    from oracle_lib import Reference

    precision = "dp" if w["prec"] == "f64" else "sp"
    ref = Reference(precision, w.get("n", 3))
    opt = ref.option(100, 100, 0.05, 0.2, 1.0)
    if w["kind"] == "vanilla":
        arg, fn = opt, ref.lib.dev_vanillaOpt
    elif w["kind"] == "cva":
        arg, fn = ref.cva(0.03, 0.6, opt, w["dates"]), ref.lib.dev_cvaEquityOption
    else:
        n = w["n"]
        c = np.full((n, n), 0.3)
        np.fill_diagonal(c, 1.0)
        vol = [0.3 if i % 2 == 0 else 0.2 for i in range(n)]
        arg, fn = ref.multi([100] * n, vol, ref.chol(c), [0] * n, [1 / n] * n, 100.0, 1.0, 0.048790164), ref.lib.dev_basketOpt
    best = None
    for _ in range(3):   # the first call pays module load and context set-up
        with _CaptureStdout() as cap:
            t0 = time.perf_counter()
            v = fn(arg, blocks, threads, sims)
            dt = time.perf_counter() - t0
        kernel = [float(x) for x in re.findall(r"Kernel done in ms\s+([0-9.eE+-]+)", cap.text)]
        rng = [float(x) for x in re.findall(r"RNG done in ms\s+([0-9.eE+-]+)", cap.text)]
        rec = {"call_ms": 1e3 * dt, "kernel_ms": kernel[-1] if kernel else None, "rng_setup_ms": rng[-1] if rng else None,
               "price": float(v.Expected), "confidence": float(v.Confidence)}
        if best is None or rec["call_ms"] < best["call_ms"]:
            best = rec
    return best


def gpu_baseline(names):
    """dev_vanillaOpt / dev_basketOpt / dev_cvaEquityOption of the UNMODIFIED reference (MonteCarloKernel.cu) on this box,
    with the launch shapes of its own drivers (vanillaOpt.cu, basketOpt.cu: 512 x 128; cvaOpt.cu: 1024 x 128)."""
    out = {}
    for name in names:
        w = WORKLOADS[name]
        sims = int(min(w["paths"], 1 << 30 if w["kind"] == "vanilla" else 1 << 28 if w.get("n", 0) <= 10 and w["kind"] == "basket" else 1 << 26))
        blocks, threads = (1024, 128) if w["kind"] == "cva" else (512, 128)
        try:
            r = _ref_gpu_call(w, sims, blocks, threads)
        except Exception as exc:   # the reference build did not travel, or its kernel failed: report, do not fake
            out[name] = {"unavailable": f"{type(exc).__name__}: {exc}"}
            continue
        units = blocks * (sims // blocks) * w["units_per_path"]
        out[name] = {"unit": unit_name(w), "value_e2e": units / (r["call_ms"] * 1e-3), "value_kernel": (units / (r["kernel_ms"] * 1e-3)) if r["kernel_ms"] else None,
                     "call_ms": r["call_ms"], "kernel_ms": r["kernel_ms"], "rng_setup_ms": r["rng_setup_ms"], "sims": sims, "launch": f"{blocks} x {threads}",
                     "price": r["price"], "confidence": r["confidence"],
                     "what": "unmodified reference MonteCarloKernel.cu (cuRAND XORWOW + shared-memory tree), nvcc -O3 sm_100a, oracle/_ref; "
                             "value_e2e = its blocking dev_* call (allocation, XORWOW seeding, kernel, copy, host sum), value_kernel = its own 'Kernel done in ms'"}
    return out


def cva_sweep(engine, reps=15):
    """The reference's cvaOpt sweep (cvaOpt.cu:70-109: grids {25, 50, 75, 250, 500} x 131 072 paths): ONE multi-job
    launch against five blocking calls of ours against five of the reference's own GPU path."""
    import statistics
    import montecarlocuda_b200 as m

    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    grids, sims = (25, 50, 75, 250, 500), 131072
    jobs = [("cva", m.CVA(0.03, 0.6, opt, n), sims, "f64") for n in grids]
    for _ in range(3):
        batch = engine.price_batch(jobs)
        serial = [engine.cva(j[1], sims, "f64") for j in jobs]
    l0 = engine.launch_count
    t_batch, t_serial = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        batch = engine.price_batch(jobs)
        t_batch.append(time.perf_counter() - t0)
    launches_batch = (engine.launch_count - l0) // reps
    for _ in range(reps):
        t0 = time.perf_counter()
        serial = [m.dev_cvaEquityOption(j[1], 1024, 128, sims, "f64", device=engine.device) for j in jobs]
        t_serial.append(time.perf_counter() - t0)
    path_steps = sims * sum(grids)
    out = {"jobs": [f"{n} dates x {sims} paths" for n in grids], "path_steps": path_steps,
           "batch_ms": 1e3 * statistics.median(t_batch), "batch_launches": launches_batch,
           "serial_ms": 1e3 * statistics.median(t_serial), "serial_launches": len(grids),
           "identical_bits": [b.Expected for b in batch] == [s_.Expected for s_ in serial], "prices": [b.Expected for b in batch]}
    try:
        from oracle_lib import Reference
        ref = Reference("dp", 3)
        times = []
        for _ in range(3):      # the first sweep also loads its module
            with _CaptureStdout():
                t0 = time.perf_counter()
                rp = [float(ref.lib.dev_cvaEquityOption(ref.cva(0.03, 0.6, ref.option(100, 100, 0.05, 0.2, 1.0), n), 1024, 128, sims).Expected) for n in grids]
                times.append(time.perf_counter() - t0)
        out["reference_gpu_ms"] = 1e3 * min(times)
        out["reference_gpu_prices"] = rp
    except Exception as exc:
        out["reference_gpu_ms"] = None
        out["reference_gpu_unavailable"] = f"{type(exc).__name__}: {exc}"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--also", default="all", help="'all' (the BASELINE configs), 'everything' (+ the other precision of each), 'none' or a comma list")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference's own GPU kernels (oracle/_ref) and the sweep comparison")
    ap.add_argument("--combine", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: cross-GPU sum started inside the pricing kernel over peer memory (peer) or one NCCL all-reduce after it")
    ap.add_argument("--peer-mode", default="push", choices=["push", "wait"],
                    help="peer combine: the last CTA only pushes and the sum is taken when the result is read (push), or it also waits for its peers (wait)")
    ap.add_argument("--no-overlap", action="store_true", help="launch without programmatic dependent launch (back-to-back jobs then do not overlap tail and start)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        return run_reference_arm(args, w, args.workload)

    import torch
    import torch.distributed as dist
    import montecarlocuda_b200 as m
    from montecarlocuda_b200.distributed import ShardedPricer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the pricing path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m.load()
    pricer = ShardedPricer(device=local, combine=args.combine, peer_mode=args.peer_mode, overlap=not args.no_overlap)

    main_run = time_workload(args.workload, w, pricer, dist, torch, rank, world, args.steps, args.warmup, sample_clocks=(rank == 0), gpu_index=local)
    also = {}
    default_also = [k for k in WORKLOADS if k != args.workload and k not in EXTRA_WORKLOADS]
    extra = [] if args.also == "none" else ([k for k in WORKLOADS if k != args.workload] if args.also == "everything" else
                                                default_also if args.also == "all" else args.also.split(","))
    for name in extra:
        ww = WORKLOADS[name]
        r = time_workload(name, ww, pricer, dist, torch, rank, world, max(2, args.steps // 2), 3, sample_clocks=(rank == 0), gpu_index=local)
        also[name] = {"value": r["value"], "unit": unit_name(ww), "ms_per_step": r["ms_per_step"], "e2e": r["e2e_value"], "e2e_ms": r["e2e_ms"],
                      "dtype": ww["prec"], "roofline": roofline(ww, r["value"] / world, r["clocks"], name),
                      "price": r["result"].Expected, "std_error": r["result"].std_error, "limbs": r["limbs"], "clocks": r["clocks"]}

    if rank == 0:
        res = main_run["result"]
        if world == 1:
            collective = "none (1 GPU)"
        elif pricer.combine == "peer" and pricer.peer_mode == "push":
            collective = ("split-phase, started inside the pricing kernel: its last CTA pushes the rank's 96 bytes into every peer mailbox over NVLink and "
                          "the kernel ends; the sum is taken out of the mailbox by one small kernel when the result is read (once per timed region)")
        elif pricer.combine == "peer":
            collective = "fused into the pricing kernel: last CTA pushes 96 bytes to every peer mailbox over NVLink, waits and adds the peers' limbs (no separate collective)"
        else:
            collective = "one int64 SUM all-reduce (NCCL) of 96 bytes per step" + (f" [{pricer.combine_note}]" if pricer.combine_note else "")
        line = {
            "metric": f"{unit_name(w)} ({args.workload})", "value": main_run["value"], "unit": unit_name(w), "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_run["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": w["prec"], "data": "synthetic",
            "config": {"workload": args.workload, "description": describe(w), "paths": w["paths"], "sharding": f"contiguous chunk ranges over {world} rank(s)",
                       "collective": collective,
                       "launch": "one kernel per step, chunks claimed dynamically; " + ("programmatic dependent launch: a step's kernel starts while the previous one drains"
                                                                                            if not args.no_overlap else "plain stream order"),
                       "l2": "not applicable: compute-bound, no resident input (parameters <= 33 KB in the constant bank, output 96 bytes)"},
            "roofline": roofline(w, main_run["value"] / world, main_run["clocks"], args.workload),
            "e2e": {"value": main_run["e2e_value"], "unit": unit_name(w), "ms_per_call": main_run["e2e_ms"],
                    "h2d_bytes_per_step": main_run["params_bytes"] + 296, "d2h_bytes_per_step": 192,
                    "api": "mcb200_vanilla/basket/cva (blocking C-ABI call: host structs in, ONE kernel launch carrying them as its parameter block, "
                           "24 flagged result words written by the kernel into mapped host memory, closing on the host)" if world == 1
                    else "ShardedPricer.price per rank (launch + cross-GPU combine + read-back + closing)"},
            "gpu_launches": main_run["launches"], "clocks": main_run["clocks"],
            "price": res.Expected, "std_error": res.std_error, "confidence": res.Confidence, "limbs": main_run["limbs"],
            "closed_form": 10.450583572185565 if w["kind"] == "vanilla" else None,
            "also": also,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_one_core(w)
                line["cpu_baseline"]["host_cores_available"] = os.cpu_count()
            except Exception as exc:  # the reference build did not travel: report, do not fake
                line["cpu_baseline"] = {"value": None, "unit": unit_name(w), "cores": 1, "kind": "reference", "sample": f"unavailable: {exc}"}
        if world == 1 and not args.no_gpu_baseline:
            names = [args.workload] + [k for k in extra if k not in SMALL_WORKLOADS]
            line["gpu_baseline"] = gpu_baseline(names)
            for name, g in line["gpu_baseline"].items():   # ours next to it, same unit
                ours = main_run if name == args.workload else None
                g["ours_value"] = ours["value"] if ours else also[name]["value"]
                g["ours_e2e"] = ours["e2e_value"] if ours else also[name]["e2e"]
            try:
                line["also"]["cva_sweep"] = cva_sweep(pricer.engine)
            except Exception as exc:
                line["also"]["cva_sweep"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
