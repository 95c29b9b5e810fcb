"""The N > 1 path on CPU: world_size 2 and 4 `gloo` process groups run the product's sharding
(`plan`, `shard_range`), its ONE collective (`distributed.combine_accumulators`, an int64 SUM
all-reduce) and its closing (`finalize`).  No GPU here, so each rank's accumulator block -- what its
kernel launch would leave in device memory -- is produced by the CPU oracle's restatement of the
chunk reduction.  The property under test is the one the GPU path relies on: any partition of the
chunks, combined by integer addition in any order, gives bit-identical results."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, n_paths, prec, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import montecarlocuda_b200 as m
    from montecarlocuda_b200.distributed import combine_accumulators
    from oracle_lib import Oracle

    oracle = Oracle()
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    p = m.plan("vanilla", opt, n_paths, prec)
    first, count = m.shard_range(p, rank, world)
    chunk_paths = int(p.chunk_units) * int(p.unit_paths)
    lo, hi = first * chunk_paths, min((first + count) * chunk_paths, n_paths)
    values = np.zeros(n_paths, dtype=np.float64 if prec == "f64" else np.float32)
    if hi > lo:  # only this rank's paths are ever computed
        values[lo:hi] = oracle.vanilla_payoffs(100.0, 100.0, 0.05, 0.2, 1.0, 77, lo, hi - lo, prec)
    local = oracle.accumulate(values, p, first, count)
    acc = torch.from_numpy(local.view(np.int64).copy())
    combine_accumulators(acc)                                   # the ONLY collective of a pricing call
    r = m.finalize(p, acc.numpy().view(np.uint64))
    np.save(Path(out_dir) / f"w{world}_r{rank}.npy", np.concatenate([acc.numpy().view(np.uint64).astype(np.float64)[:0],
                                                                     np.array([r.Expected, r.Confidence, r.sum, r.sumsq, r.n_paths])]))
    np.save(Path(out_dir) / f"acc_w{world}_r{rank}.npy", acc.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("prec,n_paths", [("f64", 150_001), ("f32", 99_999)])
def test_gloo_ranks_bit_identical(tmp_path, prec, n_paths):
    results = {}
    port = 29500 + (os.getpid() % 2000)
    for world in (1, 2, 4):
        if world == 1:
            _single(tmp_path, n_paths, prec)
        else:
            mp.spawn(_worker, args=(world, port + world, n_paths, prec, str(tmp_path)), nprocs=world, join=True)
        per_rank = [np.load(tmp_path / f"w{world}_r{r}.npy") for r in range(world)]
        accs = [np.load(tmp_path / f"acc_w{world}_r{r}.npy") for r in range(world)]
        for a in accs[1:]:
            assert np.array_equal(a, accs[0])            # every rank holds the same combined block
        for v in per_rank[1:]:
            assert np.array_equal(v, per_rank[0])
        results[world] = (accs[0], per_rank[0])
    for world in (2, 4):
        assert np.array_equal(results[world][0], results[1][0])   # limbs: bit-identical across world sizes
        assert np.array_equal(results[world][1], results[1][1])   # price, half-width, sums: bit-identical
    assert results[1][1][4] == n_paths


def _single(out_dir, n_paths, prec):
    sys.path.insert(0, str(ROOT / "tests"))
    import montecarlocuda_b200 as m
    from oracle_lib import Oracle

    oracle = Oracle()
    opt = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
    p = m.plan("vanilla", opt, n_paths, prec)
    values = oracle.vanilla_payoffs(100.0, 100.0, 0.05, 0.2, 1.0, 77, 0, n_paths, prec)
    acc = oracle.accumulate(values, p).view(np.int64)
    r = m.finalize(p, acc.view(np.uint64))
    np.save(Path(out_dir) / "w1_r0.npy", np.array([r.Expected, r.Confidence, r.sum, r.sumsq, r.n_paths]))
    np.save(Path(out_dir) / "acc_w1_r0.npy", acc)


def _status_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from montecarlocuda_b200 import _lib
    from montecarlocuda_b200.distributed import agree_on_status

    everyone_fine = agree_on_status(_lib.OK)
    one_timed_out = agree_on_status(_lib.ERR_PEER_TIMEOUT if rank == world - 1 else _lib.OK)   # only the last rank saw it
    np.save(Path(out_dir) / f"status_r{rank}.npy", np.array([everyone_fine, one_timed_out]))
    dist.destroy_process_group()


def test_ranks_agree_on_a_failure_only_one_of_them_saw(tmp_path):
    """ShardedPricer(agree_on_errors=True): a peer timeout that one rank observed becomes the status of the job on every
    rank (one MAX all-reduce), instead of a program whose ranks disagree on whether the job succeeded."""
    from montecarlocuda_b200 import _lib
    world = 3
    mp.spawn(_status_worker, args=(world, 29500 + (os.getpid() % 2000) + 17, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"status_r{r}.npy")
        assert list(got) == [_lib.OK, _lib.ERR_PEER_TIMEOUT]
