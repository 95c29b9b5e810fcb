"""Path sharding over one process per GPU (torch.distributed is plumbing only).

The job's chunks are split into contiguous ranges, rank g of G owning chunks
[n_chunks * g / G, n_chunks * (g + 1) / G).  Each rank launches ONE kernel on its GPU, which leaves
a 96-byte accumulator of exact integer limbs in device memory; ONE int64 SUM all-reduce (NCCL over
NVLink, enqueued on the same stream as the kernel) combines them.  Integer addition is associative,
so the combined limbs -- and therefore price and standard error -- are bit-identical for any G and
any reduction order inside the collective.  The reference has no multi-GPU path.
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


class ShardedPricer:
    """One rank's view of a sharded pricing job: persistent engine + device accumulator."""

    RING = 32

    def __init__(self, engine: Engine | None = None, device: int | None = None, group=None):
        import torch

        if device is None:
            device = torch.cuda.current_device()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.engine = engine or Engine(device)
        self.group = group
        # a ring of accumulator blocks zeroed in one memset every RING steps (no memset kernel per step)
        self.ring = torch.zeros((self.RING, _lib.ACC_WORDS), dtype=torch.int64, device=self.device)
        self.slot = 0
        self.acc = self.ring[0]
        self.host = torch.zeros(_lib.ACC_WORDS, dtype=torch.int64).pin_memory()
        self._job_key = None
        self._job = None

    def _world(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def enqueue(self, workload: str, params, n_paths: int, precision=_lib.F64, seed: int = DEFAULT_SEED):
        """Asynchronously: zero the accumulator, run this rank's shard, all-reduce.  Returns the plan."""
        torch = self.torch
        key = (workload, id(params), n_paths, precision, seed)
        if key != self._job_key:        # plan, shard range and C structs of a repeated job are built once
            rank, world = self._world()
            p = plan(workload, params, n_paths, precision)
            first, count = shard_range(p, rank, world)
            lib = self.engine._lib
            fn = {"vanilla": lib.mcb200_vanilla_launch, "basket": lib.mcb200_basket_launch, "cva": lib.mcb200_cva_launch}[workload]
            self._job_key, self._job = key, (p, first, count, fn, params._c(), params)
        p, first, count, fn, c_params, _ = self._job
        with torch.cuda.device(self.device):
            self.slot = (self.slot + 1) % self.RING
            if self.slot == 0:
                self.ring.zero_()
            self.acc = self.ring[self.slot]
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(fn(self.engine.handle, C.byref(p), C.byref(c_params), seed, first, count,
                          C.c_void_p(self.acc.data_ptr()), C.c_void_p(stream)), self.engine.handle)
            combine_accumulators(self.acc, self.group)
        return p

    def result(self, p: _lib.PlanT) -> OptionValue:
        """Device -> host copy of the combined accumulator and the closing formulas."""
        self.host.copy_(self.acc, non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return finalize(p, self.host.numpy().view(np.uint64))

    def price(self, workload: str, params, n_paths: int, precision=_lib.F64, seed: int = DEFAULT_SEED) -> OptionValue:
        return self.result(self.enqueue(workload, params, n_paths, precision, seed))


def price_sharded(workload: str, params, n_paths: int, precision=_lib.F64, seed: int = DEFAULT_SEED,
                  pricer: ShardedPricer | None = None) -> OptionValue:
    """Price one job over all ranks of the default process group; every rank returns the same bits."""
    pricer = pricer or ShardedPricer()
    return pricer.price(workload, params, n_paths, _prec(precision), seed)


__all__ = ["ShardedPricer", "combine_accumulators", "price_sharded", "OptionData", "MultiOptionData", "CVA"]
