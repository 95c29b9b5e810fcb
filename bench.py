#!/usr/bin/env python
"""bench.py -- the Monte Carlo pricing hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pricing job of the workload (BASELINE.json configs), sharded over the N ranks:
    zero accumulator -> ONE kernel over this rank's chunk range, whose last CTA adds the ranks' integer limbs over
    peer memory (N > 1; --combine nccl: ONE int64 all-reduce after the kernel instead).
Headline workload: European call, 2^32 paths, fp64 (BASELINE.json configs[1]); the fp32 run of the
same config and the other configs are reported under "also" (`--also none` to skip them).

Keys beyond the driver contract:
  roofline     binding pipe of the kernel (fp64 / MUFU / issue), achieved = paths/s x canonical
               per-path work of SURVEY.md 8(d), peak = pipe width x 148 SMs x clocks.max.sm
  cpu_baseline the reference's own MonteCarloHost.c (oracle/_ref, gcc -O2), 1 core, bounded sample
  e2e          the same job through the blocking C-ABI call a reference user makes
               (host structs in, OptionValue out: parameter upload, kernel, 96-byte read-back, closing)
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
# `ncu --set full` capture of ONE launch of the workload's kernel at the full path count, condensed by
# tools/ncu_summary.py (committed under profiles/): source of roofline.traffic (dram__bytes_read.sum +
# dram__bytes_write.sum; the path has no HBM-resident data) and of roofline.pipe_active (what the counters say)
NCU_SUMMARY = {
    "vanilla_f64_2p32": "profiles/r01p_vanilla_f64_2p32.txt",
    "vanilla_f32_2p32": "profiles/r01p_vanilla_f32_2p32.txt",
    "basket10_f64_2p28": "profiles/r01q_basket10_f64_2p28.txt",
    "cva50_f64_2p26": "profiles/r01q_cva50_f64_2p26.txt",
    "basket64_f32_2p30": "profiles/r01q_basket64_f32_2p30_tensor.txt",
}
_BYTE_UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
_PIPE_METRICS = {
    "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "xu_mufu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_slots": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "shared_memory": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
}


def ncu_summary(name):
    """{metric: (value, unit)} of the committed ncu summary of a workload's kernel, or {}."""
    path = NCU_SUMMARY.get(name)
    out = {}
    if not path:
        return out
    try:
        for line in (ROOT / path).read_text().splitlines():
            parts = line.split()
            if len(parts) >= 2 and ("__" in parts[0]):
                try:
                    out[parts[0]] = (float(parts[1].replace(",", "")), parts[2] if len(parts) > 2 else "")
                except ValueError:
                    pass
    except OSError:
        pass
    return out


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
    barrier()
    sampler = ClockSampler(gpu_index) if sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
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
                launches=launches * world, clocks=clocks, result=result, params_bytes=param_bytes(w))


def param_bytes(w):
    if w["kind"] == "vanilla":
        return 40
    if w["kind"] == "cva":
        return 72
    n = w["n"]
    return 8 * (4 * n + n * n + 3)


def roofline(w, value, clocks, name=None):
    peaks = measured_peaks()
    f_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    pipe = PIPE_PER_CLK_PER_SM[w["bound"]]
    peak = pipe * SM_COUNT * f_max * 1e6
    achieved = value * w["work"]
    out = {"bound": {"fp64": "fp64-pipe", "mufu": "mufu-pipe", "issue": "issue-slots"}[w["bound"]],
           "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "Ginstr/s", "frac": achieved / peak, "traffic": None,
           "work_per_unit": w["work"], "work_source": "SURVEY.md 8(d) canonical algorithm (Philox4x32-10 + Box-Muller + libdevice-cost transcendentals)",
           "peak_source": f"{pipe:g} thread-instr/clk/SM x {SM_COUNT} SMs x clocks.max.sm {f_max:g} MHz (no HBM or tensor roof applies: the kernel touches 96 bytes)"}
    f_run = (clocks or {}).get("sm_mhz")
    if f_run:
        out["frac_at_sampled_clock"] = achieved / (pipe * SM_COUNT * f_run * 1e6)
    summary = ncu_summary(name)
    if summary:
        rd, wr = summary.get("dram__bytes_read.sum"), summary.get("dram__bytes_write.sum")
        if rd and wr:
            out["traffic"] = rd[0] * _BYTE_UNITS.get(rd[1], 1.0) + wr[0] * _BYTE_UNITS.get(wr[1], 1.0)
            out["traffic_source"] = NCU_SUMMARY[name] + " (bytes per launch, 1 GPU, whole job)"
        # the hardware's own answer next to the canonical-work fraction: busiest pipes of that capture (fractions of
        # their peak while the kernel ran); frac > 1 only says the kernel needs fewer instructions than the canon
        active = {k: round(summary[m][0] / 100.0, 4) for k, m in _PIPE_METRICS.items() if m in summary and summary[m][0] >= 1.0}
        if active:
            out["pipe_active"] = dict(sorted(active.items(), key=lambda kv: -kv[1]))
            out["pipe_active_source"] = NCU_SUMMARY[name] + " (ncu --set full, one launch at the full path count)"
    if "work_note" in w:
        out["work_source"] = "derived by the rule of SURVEY.md 8(d): " + w["work_note"]
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
    ap.add_argument("--combine", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: cross-GPU sum inside the pricing kernel over peer memory (peer) or one NCCL all-reduce after it")
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
    pricer = ShardedPricer(device=local, combine=args.combine)

    main_run = time_workload(args.workload, w, pricer, dist, torch, rank, world, args.steps, args.warmup, sample_clocks=(rank == 0), gpu_index=local)
    also = {}
    extra = [] if args.also == "none" else ([k for k in WORKLOADS if k != args.workload and (args.also == "everything" or k not in EXTRA_WORKLOADS)]
                                                if args.also in ("all", "everything") else args.also.split(","))
    for name in extra:
        ww = WORKLOADS[name]
        r = time_workload(name, ww, pricer, dist, torch, rank, world, max(2, args.steps // 2), 3, sample_clocks=(rank == 0), gpu_index=local)
        also[name] = {"value": r["value"], "unit": unit_name(ww), "ms_per_step": r["ms_per_step"], "e2e": r["e2e_value"],
                      "dtype": ww["prec"], "roofline": roofline(ww, r["value"] / world, r["clocks"], name),
                      "price": r["result"].Expected, "std_error": r["result"].std_error, "clocks": r["clocks"]}

    if rank == 0:
        res = main_run["result"]
        line = {
            "metric": f"{unit_name(w)} ({args.workload})", "value": main_run["value"], "unit": unit_name(w), "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_run["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": w["prec"], "data": "synthetic",
            "config": {"workload": args.workload, "description": describe(w), "paths": w["paths"], "sharding": f"contiguous chunk ranges over {world} rank(s)",
                       "collective": "none (1 GPU)" if world == 1 else
                       ("fused into the pricing kernel: last CTA pushes 96 bytes to every peer mailbox over NVLink and adds the peers' limbs (no separate collective)"
                        if pricer.combine == "peer" else "one int64 SUM all-reduce (NCCL) of 96 bytes per step" + (f" [{pricer.combine_note}]" if pricer.combine_note else "")),
                       "l2": "not applicable: compute-bound, no resident input (parameters <= 33 KB in the constant bank, output 96 bytes)"},
            "roofline": roofline(w, main_run["value"] / world, main_run["clocks"], args.workload),
            "e2e": {"value": main_run["e2e_value"], "unit": unit_name(w), "h2d_bytes_per_step": main_run["params_bytes"] + 208,
                    "d2h_bytes_per_step": 96, "api": "mcb200_vanilla/basket/cva (blocking C-ABI call, host structs in, result out)" if world == 1
                    else "ShardedPricer.price per rank (launch + cross-GPU combine + read-back + closing)"},
            "gpu_launches": main_run["launches"], "clocks": main_run["clocks"],
            "price": res.Expected, "std_error": res.std_error, "confidence": res.Confidence,
            "closed_form": 10.450583572185565 if w["kind"] == "vanilla" else None,
            "also": also,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_one_core(w)
                line["cpu_baseline"]["host_cores_available"] = os.cpu_count()
            except Exception as exc:  # the reference build did not travel: report, do not fake
                line["cpu_baseline"] = {"value": None, "unit": unit_name(w), "cores": 1, "kind": "reference", "sample": f"unavailable: {exc}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
