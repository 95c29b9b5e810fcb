"""Path sharding over one process per GPU (torch.distributed is plumbing only).

The job's chunks are split into contiguous ranges, rank g of G owning chunks
[n_chunks * g / G, n_chunks * (g + 1) / G).  Each rank launches ONE kernel on its GPU, which leaves
a 96-byte accumulator of exact integer limbs in device memory.  The limbs of all ranks are added
  * "peer" (default on GPUs): inside that same kernel -- its last CTA pushes the limbs into every peer's
    mailbox over NVLink (CUDA IPC memory; mcb200_peer_*, include/mcb200.h; csrc/device_common.cuh peer_exchange).
    Split phase by default: the kernel ends after the push, so ranks do not lock-step on the slowest one and
    back-to-back jobs overlap; the totals of the LAST enqueued job are summed out of the mailbox by one small
    kernel when result() asks for them.  `peer_mode="wait"` keeps the single-phase variant (the last CTA also
    waits for its peers and leaves the job's totals in the accumulator: the kernel is the collective);
  * "nccl": by ONE int64 SUM all-reduce enqueued on the same stream as the kernel (also the gloo path of the
    CPU tests).
Integer addition is associative, so the combined limbs -- and therefore price and standard error -- are
bit-identical for any G, either transport and any order of arrival.  The reference has no multi-GPU path
(SURVEY.md 2.1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import DEFAULT_SEED, CVA, Engine, MultiOptionData, OptionData, OptionValue, _prec, finalize, plan, shard_range


def combine_accumulators(acc, group=None):
    """All-reduce (SUM, int64) an accumulator block in place: a torch tensor of ACC_WORDS int64 on
    the device (NCCL) or on the host (gloo).  The ONLY collective of a pricing call."""
    import torch
    import torch.distributed as dist

    if acc.dtype != torch.int64 or acc.numel() != _lib.ACC_WORDS:
        raise ValueError("accumulator must be int64[ACC_WORDS]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def _launch(engine: Engine, workload: str, p: _lib.PlanT, params, seed: int, first: int, count: int, acc, stream: int):
    lib = engine._lib
    c = params._c()
    fn = {"vanilla": lib.mcb200_vanilla_launch, "basket": lib.mcb200_basket_launch, "cva": lib.mcb200_cva_launch}[workload]
    _lib.check(fn(engine.handle, C.byref(p), C.byref(c), seed, first, count, C.c_void_p(acc.data_ptr()),
                  C.c_void_p(stream)), engine.handle)


class PeerGroup:
    """This rank's membership of a peer-memory combine group, attached to an Engine.  `exchange` turns this
    rank's 64-byte mailbox handle into the list of all ranks' handles (any transport: here torch.distributed).
    Every rank calls exchange() exactly once, whatever happens locally (a rank whose mailbox could not be created
    contributes an empty handle), so a local failure cannot leave the others waiting in the collective."""

    def __init__(self, engine: Engine, rank: int, world: int, exchange):
        self.engine, self.rank, self.world = engine, rank, world
        lib = engine._lib
        self._peer = C.c_void_p()
        handle = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        status = lib.mcb200_peer_create(engine.handle, rank, world, C.byref(self._peer), handle)
        handles = exchange(bytes(handle) if status == _lib.OK else b"")
        _lib.check(status, engine.handle)
        if len(handles) != world or any(len(h) != _lib.PEER_HANDLE_BYTES for h in handles):
            self.close()
            raise _lib.Mcb200Error(_lib.ERR_CUDA, "a peer rank could not create its mailbox")
        blob = (C.c_ubyte * (_lib.PEER_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
        status = lib.mcb200_peer_connect(self._peer, blob)
        if status != _lib.OK:
            msg = lib.mcb200_last_error(engine.handle).decode()
            self.close()
            raise _lib.Mcb200Error(status, msg or "mcb200_peer_connect failed")
        self.attach()

    def attach(self):
        _lib.check(self.engine._lib.mcb200_peer_attach(self.engine.handle, self._peer), self.engine.handle)

    def detach(self):
        _lib.check(self.engine._lib.mcb200_peer_attach(self.engine.handle, None), self.engine.handle)

    def set_mode(self, mode: str):
        code = {"wait": _lib.PEER_WAIT, "push": _lib.PEER_PUSH}[mode]
        _lib.check(self.engine._lib.mcb200_peer_set_mode(self._peer, code), self.engine.handle)

    def pull(self, acc, stream: int):
        """Enqueue the second phase of the most recent push-mode launch: job totals -> acc (12 int64 on the device)."""
        _lib.check(self.engine._lib.mcb200_peer_pull(self._peer, C.c_void_p(acc.data_ptr()), C.c_void_p(stream)), self.engine.handle)

    def close(self):
        if self._peer:
            self.engine._lib.mcb200_peer_destroy(self._peer)
            self._peer = C.c_void_p()


def _exchange_over_process_group(group):
    import torch.distributed as dist

    def exchange(handle: bytes):
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, handle, group=group)
        return out

    return exchange


def agree_on_status(status: int, group=None, device=None) -> int:
    """The worst status any rank of the group saw (one MAX all-reduce of a 4-byte word): every rank gets the same
    answer, so a failure only SOME ranks observed -- e.g. a peer that launched after the others' mailbox timeout --
    becomes a failure of the job on all of them instead of a diverging program."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(status)
    word = torch.tensor([int(status)], dtype=torch.int32, device=device if device is not None else "cpu")
    dist.all_reduce(word, op=dist.ReduceOp.MAX, group=group)
    return int(word.item())


class ShardedPricer:
    """One rank's view of a sharded pricing job: persistent engine + device accumulator.
    combine = "peer": the cross-GPU sum runs inside the pricing kernel over peer memory (peer_mode "push": split phase,
    "wait": single phase); "nccl": one all-reduce after it; "auto": peer when the group has more than one rank,
    falling back to nccl (with the reason kept in `combine_note`) only if the peer mailboxes cannot be mapped.
    overlap: launch with programmatic dependent launch, so that back-to-back jobs overlap tail and start.
    agree_on_errors: result() ends with one tiny MAX all-reduce of the ranks' status words, so that an error only some
    ranks saw (a peer timeout) is raised on every rank; off by default -- it puts a collective back behind every job."""

    RING = 32

    def __init__(self, engine: Engine | None = None, device: int | None = None, group=None, combine: str = "auto",
                 peer_mode: str = "push", overlap: bool = True, agree_on_errors: bool = False):
        import torch

        if device is None:
            device = torch.cuda.current_device()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.engine = engine or Engine(device)
        self.engine.set_overlap(overlap)
        self.group = group
        self.agree_on_errors = agree_on_errors
        # a ring of accumulator blocks zeroed in one memset every RING steps (no memset kernel per step)
        self.ring = torch.zeros((self.RING, _lib.ACC_WORDS), dtype=torch.int64, device=self.device)
        self.slot = 0
        self.acc = self.ring[0]
        self.host = torch.zeros(_lib.ACC_WORDS, dtype=torch.int64).pin_memory()
        self._host_words = self.host.numpy().view(np.uint64)
        self._job_key = None
        self._job = None
        self._pulled = True
        self.peers = None
        self.combine_note = ""
        rank, world = self._world()
        if combine not in ("auto", "peer", "nccl"):
            raise ValueError("combine must be 'auto', 'peer' or 'nccl'")
        if peer_mode not in ("push", "wait"):
            raise ValueError("peer_mode must be 'push' or 'wait'")
        self.peer_mode = peer_mode
        self.combine = "nccl"
        if world > 1 and combine in ("auto", "peer"):
            import torch.distributed as dist
            failure = None
            try:
                with torch.cuda.device(self.device):
                    self.peers = PeerGroup(self.engine, rank, world, _exchange_over_process_group(self.group))
                    self.peers.set_mode(peer_mode)
                self.combine = "peer"
            except Exception as exc:  # mapping peer memory can be refused by the platform (IPC disabled, no P2P)
                failure = exc
            # every rank must take the same route: agree before anybody raises or falls back
            flag = torch.tensor([1 if self.combine == "peer" else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                if self.peers is not None:
                    self.peers.close()
                    self.peers = None
                self.combine = "nccl"
                self.combine_note = (f"peer mailboxes unavailable ({failure}); using the NCCL all-reduce" if failure is not None
                                     else "a peer rank could not map the mailboxes; using the NCCL all-reduce")
                if combine == "peer":
                    raise RuntimeError(self.combine_note.replace("; using the NCCL all-reduce", ""))

    def _world(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def enqueue(self, workload: str, params, n_paths: int, precision=_lib.F64, seed: int = DEFAULT_SEED):
        """Asynchronously: run this rank's shard and start the combine.  Returns the plan."""
        torch = self.torch
        # keyed by VALUE: the parameter dataclasses are mutable, `opt.k = k; pricer.price(...)` must price the new strike
        key = (workload, params._snapshot(), n_paths, precision, seed)
        if key != self._job_key:        # plan, shard range and C structs of a repeated job are built once
            rank, world = self._world()
            p = plan(workload, params, n_paths, precision)
            first, count = shard_range(p, rank, world)
            lib = self.engine._lib
            fn = {"vanilla": lib.mcb200_vanilla_launch, "basket": lib.mcb200_basket_launch, "cva": lib.mcb200_cva_launch}[workload]
            self._job_key, self._job = key, (p, first, count, fn, params._c())   # the struct owns its buffers
        p, first, count, fn, c_params = self._job
        # (no torch.cuda.device() around this: the library selects its context's device itself, and the stream is asked
        # for by device -- the context manager alone was 10 us of a 15 us launch)
        push = self.combine == "peer" and self.peer_mode == "push"
        if not push:                     # the launch ADDS to the block: a zeroed one (push mode never touches it)
            self.slot = (self.slot + 1) % self.RING
            if self.slot == 0:
                with torch.cuda.device(self.device):
                    self.ring.zero_()
            self.acc = self.ring[self.slot]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(fn(self.engine.handle, C.byref(p), C.byref(c_params), seed, first, count,
                      C.c_void_p(self.acc.data_ptr()), C.c_void_p(stream)), self.engine.handle)
        if self.combine != "peer":       # "peer": the kernel itself pushed (and, in wait mode, summed) the limbs
            with torch.cuda.device(self.device):
                combine_accumulators(self.acc, self.group)
        self._pulled = not push
        self._host_is_current = False
        return p

    def result(self, p: _lib.PlanT) -> OptionValue:
        """The combined accumulator of the LAST enqueued job: [second phase of the combine ->] device -> host copy ->
        closing formulas."""
        if not self.agree_on_errors:
            return self._result(p)
        status, mine = _lib.OK, None
        try:
            mine = self._result(p)
        except _lib.Mcb200Error as exc:
            status = exc.status
        worst = agree_on_status(status, self.group, self.device)
        if worst != _lib.OK:
            raise _lib.Mcb200Error(worst, "the sharded job failed on this rank" if status != _lib.OK
                                   else "the sharded job failed on another rank of the group")
        return mine

    def _result(self, p: _lib.PlanT) -> OptionValue:
        stream = self.torch.cuda.current_stream(self.device)
        if not self._pulled:
            # the pull kernel writes the 12 words straight into the pinned host block (unified addressing: the
            # device reaches pinned memory under the same pointer): no copy operation behind it
            self.peers.pull(self.host, stream.cuda_stream)
            self._pulled = True
            self._host_is_current = True
        elif not getattr(self, "_host_is_current", False):
            with self.torch.cuda.device(self.device):
                self.host.copy_(self.acc, non_blocking=True)
        stream.synchronize()
        return finalize(p, self._host_words)

    def price(self, workload: str, params, n_paths: int, precision=_lib.F64, seed: int = DEFAULT_SEED) -> OptionValue:
        return self.result(self.enqueue(workload, params, n_paths, precision, seed))


def price_sharded(workload: str, params, n_paths: int, precision=_lib.F64, seed: int = DEFAULT_SEED,
                  pricer: ShardedPricer | None = None) -> OptionValue:
    """Price one job over all ranks of the default process group; every rank returns the same bits."""
    pricer = pricer or ShardedPricer()
    return pricer.price(workload, params, n_paths, _prec(precision), seed)


__all__ = ["PeerGroup", "ShardedPricer", "agree_on_status", "combine_accumulators", "price_sharded", "OptionData", "MultiOptionData", "CVA"]
