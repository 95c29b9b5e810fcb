"""In-tree build of the mcb200 shared libraries (nvcc cross-compiles sm_100a without a GPU).

Outputs, all under montecarlocuda_b200/lib/ (git-ignored, shipped to the GPU box by gpurun):
  libmcb200.so                    kernels + engine + extended C ABI (include/mcb200.h)
  libmcb200_{dp,sp}[_nN].so       the reference's dev_* entry points (include/MonteCarlo.h),
                                  one per precision and basket width N in {3, 10, 64}
  pipe_peaks                      FMA / DFMA / MUFU / integer pipe micro-benchmark (tools/microbench/, bench evidence)
  libmcb200.manifest.json         sha256 of libmcb200.so, of the sources it was built from and of the SASS of every
                                  pricing kernel: bench.py ties ncu counters to the library it actually loaded

Run as `python -m montecarlocuda_b200.build` or through `__graft_entry__.build()`.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"
OBJ = PKG / "build"
INCLUDE = ROOT / "include"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"] + ARCH

CUDA_UNITS = ["kernels_vanilla.cu", "kernels_basket.cu", "kernels_cva.cu", "kernels_debug.cu", "engine.cu"]
HEADERS = ["device_common.cuh", "device_math.cuh", "device_math64.cuh", "tables64.inc", "launch.h", "table_lock.h",
           "workload_vanilla.cuh", "basket_tc.cuh"]
DROPIN_WIDTHS = (3, 10, 64, 100)   # the reference's compile-time N: its default, the two BASELINE baskets, one beyond the register templates


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the mcb200 libraries cannot be built (there is no CPU build)")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print("  $", " ".join(str(c) for c in cmd), flush=True)
    res = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"build step failed ({res.returncode}): {' '.join(map(str, cmd))}\n{res.stdout}\n{res.stderr}")
    return res


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every CUDA translation unit for sm_100a and link the shared libraries."""
    nvcc = _nvcc()
    LIB.mkdir(exist_ok=True)
    OBJ.mkdir(exist_ok=True)
    common_deps = [CSRC / h for h in HEADERS] + [INCLUDE / "mcb200.h", Path(__file__)]

    jobs = []
    objects = []
    for unit in CUDA_UNITS:
        src = CSRC / unit
        obj = OBJ / (Path(unit).stem + ".o")
        objects.append(obj)
        if force or _newer(obj, [src] + common_deps):
            jobs.append([nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj])
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            list(pool.map(lambda c: _run(c, verbose), jobs))

    core = LIB / "libmcb200.so"
    if force or _newer(core, objects):
        _run([nvcc, "-shared"] + ARCH + ["-o", core] + objects, verbose)

    # the reference's entry points: thin host-only shims, one per (precision, N)
    shim_src = CSRC / "dropin.cpp"
    shim_deps = [shim_src, INCLUDE / "MonteCarlo.h", INCLUDE / "mcb200.h", Path(__file__)]
    cxx = shutil.which("g++") or "g++"
    for prec, define in (("dp", []), ("sp", ["-DMCB200_SINGLE"])):
        for n in DROPIN_WIDTHS:
            name = f"libmcb200_{prec}.so" if n == 3 else f"libmcb200_{prec}_n{n}.so"
            out = LIB / name
            if force or _newer(out, shim_deps + [core]):
                _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", f"-DN={n}"] + define +
                     ["-I", INCLUDE, shim_src, "-o", out, f"-L{LIB}", "-lmcb200", "-Wl,-rpath,$ORIGIN"], verbose)

    # the reference's CPU helper API (pure C, no CUDA) and the non-interactive CLI drivers
    host_src = CSRC / "hostapi.c"
    cc = shutil.which("gcc") or "gcc"
    for prec, define in (("dp", []), ("sp", ["-DMCB200_SINGLE"])):
        for n in DROPIN_WIDTHS:
            suffix = f"{prec}" if n == 3 else f"{prec}_n{n}"
            out = LIB / f"libmcb200_hostapi_{suffix}.so"
            if force or _newer(out, [host_src, INCLUDE / "MonteCarlo.h", Path(__file__)]):
                _run([cc, "-O2", "-std=gnu11", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", f"-DN={n}"] + define +
                     ["-I", INCLUDE, host_src, "-o", out, "-lm"], verbose)
            for app in ("vanillaOpt", "basketOpt", "cvaOpt"):
                src = CSRC / "apps" / f"{app}.c"
                if not src.exists() or (app != "basketOpt" and n != 3):
                    continue
                exe = LIB / f"mcb200_{app}_{suffix}"
                deps = [src, CSRC / "apps" / "cli_common.h", INCLUDE / "MonteCarlo.h", INCLUDE / "mcb200.h", out, Path(__file__)]
                if force or _newer(exe, deps):
                    dropin = "mcb200_" + suffix
                    _run([cc, "-O2", "-std=gnu11", "-Wall", f"-DN={n}"] + define + ["-I", INCLUDE, src, "-o", exe, f"-L{LIB}",
                         f"-l{dropin}", f"-lmcb200_hostapi_{suffix}", "-lmcb200", "-lm", "-Wl,-rpath,$ORIGIN"], verbose)

    peaks_src = ROOT / "tools" / "microbench" / "pipe_peaks.cu"
    if peaks_src.exists():
        peaks = LIB / "pipe_peaks"
        if force or _newer(peaks, [peaks_src, Path(__file__)]):
            _run([nvcc, "-O3", "-std=c++17", "-lineinfo"] + ARCH + [peaks_src, "-o", peaks], verbose)
    manifest = LIB / "libmcb200.manifest.json"
    if force or _newer(manifest, [core]):
        write_manifest(core, manifest)
    return core


def source_hash() -> str:
    """sha256 over the sources and flags libmcb200.so is built from (stable across rebuilds of the same tree)."""
    h = hashlib.sha256()
    for name in sorted(CUDA_UNITS + HEADERS):
        h.update(name.encode())
        h.update((CSRC / name).read_bytes())
    h.update((INCLUDE / "mcb200.h").read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def file_hash(path: Path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest()


def kernel_sass_hashes(library: Path) -> dict:
    """{demangled kernel name: sha256 of its SASS instruction text} for every pricing kernel in the library."""
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([tool, "-sass", str(library)], capture_output=True, text=True)
    if out.returncode != 0:
        return {}
    filt = shutil.which("cu++filt") or "/usr/local/cuda/bin/cu++filt"
    hashes, name, h = {}, None, None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name is not None:
                hashes[name] = h.hexdigest()
            name, h = m.group(1), hashlib.sha256()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?)\s*;", line)
        if m and name is not None:
            h.update(m.group(1).encode())
            h.update(b"\n")
    if name is not None:
        hashes[name] = h.hexdigest()
    keep = {k: v for k, v in hashes.items() if "accumulate" in k}
    try:
        res = subprocess.run([filt] + list(keep), capture_output=True, text=True)
        names = res.stdout.splitlines() if res.returncode == 0 else []
    except OSError:
        names = []
    if len(names) == len(keep):
        return {short_kernel_name(n): v for n, v in zip(names, keep.values())}
    return keep


def short_kernel_name(demangled: str) -> str:
    """'void mcb::mc_accumulate_kernel<mcb::Vanilla<double, 2, 1, true>>(...)' -> 'mc_accumulate_kernel<Vanilla<double, 2, 1, true>>'
    (ncu prints bool template arguments as 1 / 0 and the toolchain's demangler as true / false: normalised here)."""
    name = demangled.strip()
    if name.startswith("void "):
        name = name[5:]
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    name = name[:cut].replace("mcb::", "")
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)
    name = re.sub(r"\btrue\b", "1", name)
    name = re.sub(r"\bfalse\b", "0", name)
    return re.sub(r"\s+", "", name)


def write_manifest(library: Path, manifest: Path) -> None:
    manifest.write_text(json.dumps({"library": library.name, "library_sha256": file_hash(library), "source_sha256": source_hash(),
                                    "nvcc_flags": NVCC_FLAGS, "kernel_sass_sha256": kernel_sass_hashes(library)}, indent=1) + "\n")


def build_oracle(verbose: bool = False) -> None:
    """Build the CPU oracle (test infrastructure) and, when the reference sources are mounted,
    the unmodified reference into oracle/_ref/.  Building the checker is not using it."""
    _run(["make", "-C", ROOT / "oracle", "oracle"], verbose)
    if Path("/root/reference/double_precision/MonteCarloHost.c").exists():
        _run(["make", "-C", ROOT / "oracle", "ref", "drivers"], verbose)


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    build_oracle(verbose=True)
    print("built:", *sorted(p.name for p in LIB.iterdir()))
