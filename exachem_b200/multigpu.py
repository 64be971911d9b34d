"""One process per GPU: rank bookkeeping and the single collective of the (T) path.

The task list shards with no data-path exchange: each rank runs the kernel tasks the
library's static cost-balanced split assigns to it (`ccsdt_options.rank/nranks`, `ccsdt_partition`) and
only the two scalars E[T], E(T) are combined -- ONE all-reduce of 16 bytes (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Replaces the two `ec.pg().reduce` calls of exachem/cc/ccsd_t/ccsd_t.cpp:262-263
and the GA atomic counter of ccsd_t_fused_driver.hpp:169-172,456.

Reduction order (for the 1e-9 Eh statement): inside a rank, per-box partials in box order, then tasks in
canonical task order; across ranks, the all-reduce's order (NCCL ring/tree: deterministic for a fixed
rank count).
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class RankInfo:
    rank: int = 0
    nranks: int = 1
    local_rank: int = 0

    @staticmethod
    def from_env() -> "RankInfo":
        return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                        int(os.environ.get("LOCAL_RANK", "0")))


def combine_energies(e1: float, e2: float, device=None, group=None):
    """The one collective: sum (E[T], E(T)) rank partials over all ranks; returns the totals on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(e1), float(e2)
    t = torch.tensor([e1, e2], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    a, b = t.tolist()
    return float(a), float(b)


def caller_side_partials(e1_total: float, e2_total: float, rank: int):
    """ExaChem's caller reduces the values `execute` returns (ccsd_t.cpp:262-263).  A driver that already
    all-reduced must hand the total back on rank 0 only, or the caller double counts."""
    return (e1_total, e2_total) if rank == 0 else (0.0, 0.0)


class SharedTaskCounter:
    """An int64 in POSIX shared memory that all ranks of one node map: the dynamic task hand-out
    (`ccsdt_set_task_counter`).  Rank 0 creates it; `name` must be agreed on beforehand (bench.py derives
    it from MASTER_PORT).  `reset()` is called by rank 0 between two barriers before every run."""

    def __init__(self, name: str, create: bool):
        from multiprocessing import shared_memory
        import ctypes
        if create:
            try:
                stale = shared_memory.SharedMemory(name=name)
                stale.close()
                stale.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=64)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
            try:  # only the creator owns the segment: keep the attaching ranks out of the resource tracker
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        self.owner = create
        self._c = ctypes.c_int64.from_buffer(self.shm.buf)
        self.address = ctypes.addressof(self._c)
        if create:
            self._c.value = 0

    def reset(self):
        self._c.value = 0

    @property
    def value(self) -> int:
        return int(self._c.value)

    def close(self):
        self.address = 0
        del self._c
        self.shm.close()
        if self.owner:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
