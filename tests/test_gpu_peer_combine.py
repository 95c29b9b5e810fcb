"""The cross-GPU combine fused into the pricing kernel (csrc/device_common.cuh peer_combine; include/mcb200.h
mcb200_peer_*), exercised on ONE GPU with virtual ranks: every rank is its own context + stream on cuda:0, the
mailboxes are connected with mcb200_peer_connect_local, and the shards' kernels run concurrently.  The real
multi-process route (CUDA IPC handles exchanged over torch.distributed) is covered by test_gpu_multi.py on a
multi-GPU box.

Only the vanilla workload can run as concurrent virtual ranks of ONE device: basket and CVA read a per-device
__constant__ table whose users are serialised (csrc/table_lock.h), so on a single device rank 1's kernel would
wait for rank 0's, whose last CTA waits for rank 1 -- the bounded wait then raises the error flag, which
test_serialised_ranks_raise_the_error_flag pins.  All kernels share the same tail (finish -> peer_combine)."""
import ctypes as C

import numpy as np
import pytest

import montecarlocuda_b200 as m
from montecarlocuda_b200 import _lib
from test_gpu_parity import VAN

pytestmark = pytest.mark.gpu


class LocalGroup:
    def __init__(self, world, device=0):
        self.lib = _lib.load()
        self.engines = [m.Engine(device) for _ in range(world)]
        self.peers = (C.c_void_p * world)()
        handle = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        for r, e in enumerate(self.engines):
            p = C.c_void_p()
            _lib.check(self.lib.mcb200_peer_create(e.handle, r, world, C.byref(p), handle), e.handle)
            self.peers[r] = p
        _lib.check(self.lib.mcb200_peer_connect_local(self.peers, world))
        for r, e in enumerate(self.engines):
            _lib.check(self.lib.mcb200_peer_attach(e.handle, self.peers[r]), e.handle)

    def close(self):
        for r, e in enumerate(self.engines):
            self.lib.mcb200_peer_destroy(self.peers[r])
            e.close()


def _fused(group, workload, params, n_paths, prec, seed):
    import torch
    from montecarlocuda_b200 import distributed as D
    world = len(group.engines)
    p = m.plan(workload, params, n_paths, prec)
    acc = torch.zeros((world, 12), dtype=torch.int64, device="cuda:0")
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for rank in range(world):   # different streams: a rank's last CTA waits for its peers' kernels
        first, count = m.shard_range(p, rank, world)
        D._launch(group.engines[rank], workload, p, params, seed, first, count, acc[rank], streams[rank].cuda_stream)
    torch.cuda.synchronize()
    return p, acc.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_fused_combine_gives_every_rank_the_job_totals(engine, world):
    group = LocalGroup(world)
    try:
        cases = [("vanilla", VAN, (1 << 22) + 12345, "f32"), ("vanilla", VAN, 1 << 22, "f64"), ("vanilla", VAN, 1000, "f64"),
                 ("vanilla", VAN, (1 << 24) + 1, "f32")]
        for workload, params, n_paths, prec in cases:
            for _ in range(3):   # repeated: sequence numbers, ticket reset and mailbox slots are reused correctly
                p, acc = _fused(group, workload, params, n_paths, prec, 2024)
                for rank in range(1, world):
                    assert np.array_equal(acc[rank], acc[0]), (workload, prec, rank)
                assert acc[0][10] == n_paths and acc[0][11] == 0
                one = getattr(engine, workload)(params, n_paths, prec, 2024)   # unfused, one device
                fin = m.finalize(p, acc[0])
                assert (one.Expected, one.Confidence, one.sum, one.sumsq) == (fin.Expected, fin.Confidence, fin.sum, fin.sumsq)
    finally:
        group.close()


def test_more_ranks_than_chunks(engine):
    # 3 chunks over 8 ranks: the ranks with an empty shard still take part in the combine
    group = LocalGroup(8)
    try:
        p, acc = _fused(group, "vanilla", VAN, 2500, "f64", 5)   # 625 draw units of 4 paths = 3 chunks of 256
        assert p.n_chunks == 3
        for rank in range(8):
            assert np.array_equal(acc[rank], acc[0])
        one = engine.vanilla(VAN, 2500, "f64", 5)
        assert m.finalize(p, acc[0]).sum == one.sum and acc[0][10] == 2500
    finally:
        group.close()


def test_serialised_ranks_raise_the_error_flag(engine):
    # two ranks launched on the SAME stream cannot overlap: the first one's wait for its peer is bounded and ends in
    # the accumulator's error flag (-> MCB200_ERR_OVERFLOW from mcb200_finalize), never in a hung device
    import torch
    from montecarlocuda_b200 import distributed as D
    group = LocalGroup(2)
    try:
        for peer in group.peers:   # the wait is bounded by wall clock (default 10 s): keep the test short
            _lib.check(group.lib.mcb200_peer_set_timeout_ms(peer, 300.0))
        p = m.plan("vanilla", VAN, 1 << 16, "f64")
        acc = torch.zeros((2, 12), dtype=torch.int64, device="cuda:0")
        stream = torch.cuda.Stream()
        for rank in range(2):
            first, count = m.shard_range(p, rank, 2)
            D._launch(group.engines[rank], "vanilla", p, VAN, 3, first, count, acc[rank], stream.cuda_stream)
        torch.cuda.synchronize()
        a = acc.cpu().numpy().view(np.uint64)
        assert a[0][11] >> 32 >= 1                 # rank 0 gave up waiting: the peer-timeout flag, not the overflow count
        with pytest.raises(m.Mcb200Error) as err:
            m.finalize(p, a[0])
        assert err.value.status == _lib.ERR_PEER_TIMEOUT
        assert a[1][11] == 0 and a[1][10] == 1 << 16   # rank 1 found rank 0's words already there
    finally:
        group.close()


def test_detached_context_launches_are_plain_shards(engine):
    group = LocalGroup(2)
    try:
        lib = _lib.load()
        for e in group.engines:
            _lib.check(lib.mcb200_peer_attach(e.handle, None), e.handle)
        p, acc = _fused(group, "vanilla", VAN, 1 << 20, "f64", 9)
        assert not np.array_equal(acc[0], acc[1])           # partial sums of the two halves
        total = acc.sum(axis=0)
        one = engine.vanilla(VAN, 1 << 20, "f64", 9)
        assert m.finalize(p, total).sum == one.sum
    finally:
        group.close()


# ---- split phase (MCB200_PEER_PUSH + mcb200_peer_pull) ------------------------------------------------
def _push_then_pull(group, workload, params, n_paths, prec, seed, stream):
    """Every virtual rank launches its shard on ONE stream (a push never waits, so they may run one after the other),
    then every rank pulls the totals of that launch out of its own mailbox."""
    import torch
    from montecarlocuda_b200 import distributed as D
    world = len(group.engines)
    p = m.plan(workload, params, n_paths, prec)
    acc = torch.full((world, 12), -1, dtype=torch.int64, device="cuda:0")   # poisoned: the pull overwrites all 12 words
    for rank in range(world):
        first, count = m.shard_range(p, rank, world)
        D._launch(group.engines[rank], workload, p, params, seed, first, count, acc[rank], stream.cuda_stream)
    for rank in range(world):
        _lib.check(group.lib.mcb200_peer_pull(group.peers[rank], C.c_void_p(acc[rank].data_ptr()), C.c_void_p(stream.cuda_stream)))
    torch.cuda.synchronize()
    return p, acc.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("world", [2, 5, 8])
def test_split_phase_combine(engine, oracle, world):
    import torch
    from test_gpu_parity import CVA50, make_basket
    group = LocalGroup(world)
    try:
        for peer in group.peers:
            _lib.check(group.lib.mcb200_peer_set_mode(peer, _lib.PEER_PUSH))
        stream = torch.cuda.Stream()
        # basket and CVA can take part now: nobody waits inside a kernel, so sharing the device's table is no deadlock
        cases = [("vanilla", VAN, (1 << 22) + 12345, "f32"), ("vanilla", VAN, 1000, "f64"), ("cva", CVA50, 50_000, "f64"),
                 ("basket", make_basket(oracle, 10), 30_000, "f64")]
        for workload, params, n_paths, prec in cases:
            for _ in range(5):   # more launches than the mailbox ring is deep on small worlds: slots are reused
                p, acc = _push_then_pull(group, workload, params, n_paths, prec, 2024, stream)
                for rank in range(1, world):
                    assert np.array_equal(acc[rank], acc[0]), (workload, prec, rank)
                assert acc[0][10] == n_paths and acc[0][11] == 0
            one = getattr(engine, workload)(params, n_paths, prec, 2024)
            fin = m.finalize(p, acc[0])
            assert (one.Expected, one.Confidence, one.sum, one.sumsq) == (fin.Expected, fin.Confidence, fin.sum, fin.sumsq)
    finally:
        group.close()


def test_split_phase_pull_of_a_missing_peer_times_out(engine):
    """Only rank 0 launches: its pull must end (bounded by the wall-clock timeout) in the peer-timeout status."""
    import torch
    from montecarlocuda_b200 import distributed as D
    group = LocalGroup(2)
    try:
        for peer in group.peers:
            _lib.check(group.lib.mcb200_peer_set_mode(peer, _lib.PEER_PUSH))
            _lib.check(group.lib.mcb200_peer_set_timeout_ms(peer, 200.0))
        p = m.plan("vanilla", VAN, 1 << 16, "f64")
        acc = torch.zeros(12, dtype=torch.int64, device="cuda:0")
        stream = torch.cuda.Stream()
        first, count = m.shard_range(p, 0, 2)
        D._launch(group.engines[0], "vanilla", p, VAN, 3, first, count, acc, stream.cuda_stream)
        _lib.check(group.lib.mcb200_peer_pull(group.peers[0], C.c_void_p(acc.data_ptr()), C.c_void_p(stream.cuda_stream)))
        torch.cuda.synchronize()
        with pytest.raises(m.Mcb200Error) as err:
            m.finalize(p, acc.cpu().numpy().view(np.uint64))
        assert err.value.status == _lib.ERR_PEER_TIMEOUT
    finally:
        group.close()
