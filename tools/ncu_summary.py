#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers the roofline argument needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r02_prof.txt
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --work WORKLOAD --manifest gpurun_out/prof_manifest.json \
           --capture profiles/r02_prof.txt [--table profiles/kernel_work.json]

The second form also records, in the committed table bench.py reads (profiles/kernel_work.json), what ONE launch of the
workload's kernel executed: warp instructions per pipe (-> thread instructions per path / path-step), the pipe
percentages, DRAM bytes -- together with the sha256 of that kernel's SASS in the library that was profiled (from the
build manifest that travelled with it).  bench.py refuses the counters when the library it loaded has another hash.
"""
import argparse
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
PIPES = {"fp64": "sm__inst_executed_pipe_fp64", "xu_mufu": "sm__inst_executed_pipe_xu", "fma": "sm__inst_executed_pipe_fma",
         "alu": "sm__inst_executed_pipe_alu", "lsu": "sm__inst_executed_pipe_lsu", "tensor": "sm__inst_executed_pipe_tensor"}
PCT = {"fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "xu_mufu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "issue_slots": "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "shared_memory": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"}
_SCALE = {"": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def number(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--work", help="workload name (bench.py WORKLOADS): record this launch in the kernel-work table")
    ap.add_argument("--manifest", help="build manifest of the library that was profiled")
    ap.add_argument("--capture", help="path of the text summary being committed (recorded in the table)")
    ap.add_argument("--table", default=str(ROOT / "profiles" / "kernel_work.json"))
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    last = None
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:95s} {row[i]:>18s} {units[i]}")
        last = (name, row)
    if not args.work or last is None:
        return
    import bench
    from montecarlocuda_b200.build import short_kernel_name

    name, row = last

    def get(metric, scale=False):
        if metric not in hdr:
            return None
        i = hdr.index(metric)
        v = number(row[i])
        return v * _SCALE.get(units[i], 1.0) if (v is not None and scale) else v

    w = bench.WORKLOADS[args.work]
    n_units = w["paths"] * w["units_per_path"]
    manifest = json.loads(Path(args.manifest).read_text())
    kernel = short_kernel_name(name)
    sass = manifest["kernel_sass_sha256"].get(kernel)
    if sass is None:
        raise SystemExit(f"kernel {kernel!r} is not in the manifest ({sorted(manifest['kernel_sass_sha256'])[:5]} ...)")
    warp = {p: get(m + ".sum") for p, m in PIPES.items()}
    warp["total"] = get("smsp__inst_executed.sum")
    entry = {
        "kernel": kernel, "sass_sha256": sass, "source_sha256": manifest["source_sha256"], "library_sha256": manifest["library_sha256"],
        "capture": args.capture, "units_per_launch": n_units, "duration_ms": get("gpu__time_duration.sum", scale=True),
        "registers_per_thread": get("launch__registers_per_thread"),
        "warp_inst": warp,
        # full warps throughout (256-thread sub-blocks, no divergence in the path loop): thread instructions = 32 x warp instructions
        "thread_inst_per_unit": {p: (32.0 * v / n_units if v is not None else None) for p, v in warp.items()},
        "pipe_pct": {p: get(m) for p, m in PCT.items()},
        "dram_bytes": (get("dram__bytes_read.sum", scale=True) or 0.0) + (get("dram__bytes_write.sum", scale=True) or 0.0),
    }
    table_path = Path(args.table)
    table = json.loads(table_path.read_text()) if table_path.exists() else {}
    table[args.work] = entry
    table_path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
    print(f"# recorded {args.work} -> {table_path}", file=sys.stderr)


if __name__ == "__main__":
    main()
