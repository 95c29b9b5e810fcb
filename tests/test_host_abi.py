"""Host-side checks that need no GPU: the C-ABI libraries load and export every symbol the headers
declare, the planning / sharding / closing logic (pure host functions of libmcb200) is right, and a
missing device fails loudly instead of falling back to the CPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_functions(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b((?:mcb200|dev)_[A-Za-z0-9_]+)\s*\(", text)))


def test_core_library_exports_every_declared_symbol(ensure_built):
    m = ensure_built
    lib = C.CDLL(str(m.library_path()))
    names = declared_functions(ROOT / "include" / "mcb200.h")
    assert len(names) >= 28
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mcb200.h but not exported"
    from montecarlocuda_b200 import _lib
    assert set(_lib.exported_symbols()) == set(names)      # the Python binding covers the whole header


@pytest.mark.parametrize("lib_name", ["libmcb200_dp.so", "libmcb200_sp.so", "libmcb200_dp_n10.so", "libmcb200_sp_n64.so", "libmcb200_dp_n100.so"])
def test_dropin_libraries_export_the_reference_symbols(ensure_built, lib_name):
    lib = C.CDLL(str(ROOT / "montecarlocuda_b200" / "lib" / lib_name))
    names = [n for n in declared_functions(ROOT / "include" / "MonteCarlo.h") if n.startswith("dev_")]
    assert names == ["dev_basketOpt", "dev_cvaEquityOption", "dev_vanillaOpt"]   # reference MonteCarloKernel.cu:483,500,517
    for name in names:
        assert hasattr(lib, name)


def test_struct_layouts_match_the_reference(ensure_built):
    # sizes measured on the reference headers: DP 40/192/16/72, SP 20/96/8/36 bytes at N = 3
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include "MonteCarlo.h"
    int main(void) { printf("%zu %zu %zu %zu %zu\n", sizeof(OptionData), sizeof(MultiOptionData), sizeof(OptionValue), sizeof(CVA), sizeof(MonteCarloData)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as tmp:
        (Path(tmp) / "t.c").write_text(src)
        for flags, want in (([], "40 192 16 72 256"), (["-DMCB200_SINGLE"], "20 96 8 36 132"), (["-DN=10"], "40 1144 16 72 1208")):
            subprocess.run(["gcc", "-I", str(ROOT / "include"), *flags, str(Path(tmp) / "t.c"), "-o", str(Path(tmp) / "t")], check=True)
            out = subprocess.run([str(Path(tmp) / "t")], capture_output=True, text=True, check=True).stdout.strip()
            assert out == want, (flags, out)


REFERENCE = Path("/root/reference")


@pytest.mark.skipif(not (REFERENCE / "double_precision" / "MonteCarlo.h").exists(), reason="the reference sources are not mounted here")
@pytest.mark.parametrize("tree,flags", [("double_precision", []), ("single_precision", ["-DMCB200_SINGLE"])])
@pytest.mark.parametrize("n", [3, 10, 64])
def test_struct_layouts_against_the_reference_header_itself(ensure_built, tree, flags, n, tmp_path):
    """Not constants typed in from a survey: the reference's OWN MonteCarlo.h (its `#define N 3` rewritten in a scratch
    copy for the wider builds, as oracle/Makefile does) and ours are compiled by the same gcc, and every struct's size
    and every field's offset must agree."""
    import subprocess
    probe = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "MonteCarlo.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu\n", sizeof(OptionData), sizeof(MultiOptionData), sizeof(OptionValue), sizeof(CVA), sizeof(MonteCarloData));
        printf("%zu %zu %zu %zu %zu\n", offsetof(OptionData, s), offsetof(OptionData, k), offsetof(OptionData, r), offsetof(OptionData, v), offsetof(OptionData, t));
        printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(MultiOptionData, s), offsetof(MultiOptionData, v), offsetof(MultiOptionData, p),
               offsetof(MultiOptionData, d), offsetof(MultiOptionData, w), offsetof(MultiOptionData, k), offsetof(MultiOptionData, t), offsetof(MultiOptionData, r));
        printf("%zu %zu\n", offsetof(OptionValue, Expected), offsetof(OptionValue, Confidence));
        printf("%zu %zu %zu %zu %zu\n", offsetof(CVA, defInt), offsetof(CVA, lgd), offsetof(CVA, ns), offsetof(CVA, option), offsetof(CVA, n));
        return 0;
    }
    """
    (tmp_path / "probe.c").write_text(probe)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    header = (REFERENCE / tree / "MonteCarlo.h").read_text()
    assert "#define N 3" in header
    # the reference header pulls in the CUDA runtime for its CudaCheck macro; the layouts do not depend on it
    header = header.replace("#define N 3", f"#define N {n}")
    header = re.sub(r'#include\s*[<"](cuda[^>"]*|curand[^>"]*|helper[^>"]*)[>"]', "", header)
    (ref_dir / "MonteCarlo.h").write_text(header)
    outs = []
    for include, extra in ((ref_dir, []), (ROOT / "include", flags + [f"-DN={n}"])):
        exe = tmp_path / f"probe_{len(outs)}"
        res = subprocess.run(["gcc", "-x", "c", "-I", str(include), *extra, str(tmp_path / "probe.c"), "-o", str(exe)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1], (tree, n, outs)


def test_no_device_fails_loudly(ensure_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    m = ensure_built
    with pytest.raises(m.Mcb200Error) as exc:
        m.Engine(0)
    assert exc.value.status == 3 and "no CPU fallback" in str(exc.value)


def test_plan_geometry(ensure_built, oracle):
    m = ensure_built
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    p = m.plan("vanilla", opt, 1 << 32, "f32")  # six fp32 normals per Philox block: 715 827 883 draw units
    assert (p.unit_paths, p.rounds, p.chunk_units, p.n_chunks) == (6, 32, 8192, 87382)
    assert (p.scale_exp_sum, p.scale_exp_sumsq) == (73, 66) and p.discount == pytest.approx(np.exp(-0.05))
    p = m.plan("vanilla", opt, 1 << 32, "f64")
    assert (p.unit_paths, p.rounds, p.chunk_units, p.n_chunks) == (4, 64, 16384, 65536)
    cva = m.CVA(0.03, 0.6, opt, 50)
    p = m.plan("cva", cva, 1 << 26, "f64")
    assert (p.unit_paths, p.rounds, p.n_chunks, p.discount) == (1, 4, 65536, 1.0)
    for units in (1, 255, 1 << 20, (1 << 25) - 1, 1 << 25, 1 << 27, 1 << 31, 1 << 40):
        assert m.plan("cva", cva, units, "f64").rounds == oracle.chunk_rounds(units)
    p = m.plan("vanilla", opt, 7, "f32")       # ragged: 7 paths = 2 draw units of 6 = 1 chunk
    assert (p.total_units, p.n_chunks) == (2, 1)
    p = m.plan("vanilla", opt, 5, "f64")       # ... and 5 paths = 2 draw units of 4 in double precision
    assert (p.total_units, p.n_chunks) == (2, 1)
    with pytest.raises(m.Mcb200Error):
        m.plan("vanilla", opt, 0, "f64")


def test_shard_ranges_partition_the_chunks(ensure_built):
    m = ensure_built
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    for n_paths in (1, 1000, (1 << 22) + 12345, 1 << 32):
        p = m.plan("vanilla", opt, n_paths, "f32")
        for world in (1, 2, 3, 4, 7, 8):
            pos = 0
            for rank in range(world):
                first, count = m.shard_range(p, rank, world)
                assert first == pos
                pos += count
            assert pos == p.n_chunks
    with pytest.raises(m.Mcb200Error):
        m.shard_range(p, 8, 8)


def test_finalize_is_the_reference_closing(ensure_built, oracle):
    # mcb200_finalize on oracle-built limbs == the reference's closing formulas (MonteCarloKernel.cu:412-423)
    m = ensure_built
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    n = 100_000
    pay = oracle.vanilla_payoffs(100.0, 100.0, 0.05, 0.2, 1.0, 9, 0, n, "f64")
    p = m.plan("vanilla", opt, n, "f64")
    acc = oracle.accumulate(pay, p)
    r = m.finalize(p, acc)
    assert r.n_paths == n
    assert r.sum == pytest.approx(pay.sum(), rel=1e-14) and r.sumsq == pytest.approx((pay * pay).sum(), rel=1e-14)
    e, c = oracle.closing(r.sum, r.sumsq, n, 0.05, 1.0)
    assert r.Expected == pytest.approx(e, rel=1e-14) and r.Confidence == pytest.approx(c, rel=1e-10)
    assert r.std_error == pytest.approx(c / 1.96 * np.exp(-0.05), rel=1e-10)
    acc_bad = acc.copy()
    acc_bad[10] -= 1                      # a missing path is an error, not a silently smaller n
    with pytest.raises(m.Mcb200Error):
        m.finalize(p, acc_bad)
    acc_bad = acc.copy()
    acc_bad[11] = 1                       # overflow / NaN flag
    with pytest.raises(m.Mcb200Error):
        m.finalize(p, acc_bad)


def test_peer_and_engine_switch_argument_checks(ensure_built):
    """Host-side validation of the additive entry points (no device needed): the peer-group API
    (include/mcb200.h, the fused cross-GPU combine) and the basket engine switch."""
    from montecarlocuda_b200 import _lib
    lib = _lib.load()
    handle = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
    peer = C.c_void_p()
    assert lib.mcb200_peer_create(None, 0, 2, C.byref(peer), handle) == _lib.ERR_INVALID
    assert lib.mcb200_peer_connect(None, handle) == _lib.ERR_INVALID
    assert lib.mcb200_peer_connect_local(None, 2) == _lib.ERR_INVALID
    assert lib.mcb200_peer_attach(None, None) == _lib.ERR_INVALID
    assert lib.mcb200_peer_destroy(None) == _lib.OK
    before = lib.mcb200_get_basket_engine()
    assert before in (0, 1)
    assert lib.mcb200_set_basket_engine(2) == _lib.ERR_INVALID and lib.mcb200_get_basket_engine() == before
    assert lib.mcb200_set_basket_engine(1) == _lib.OK and lib.mcb200_get_basket_engine() == 1
    assert lib.mcb200_set_basket_engine(before) == _lib.OK
