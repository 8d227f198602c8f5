"""Multi-GPU arg-min exchange: one process per GPU, contiguous shards of the candidate rows.

The candidate batch shards trivially (the reference already chunks it over worker processes,
reactive_planner.py:200-202); only the final selection couples the shards.  NCCL has no arg-min
reduction and (float64 cost, int64 row) does not fit one 64-bit key, so the "all-reduce with
op = argmin" is ONE all-gather of the 16-byte winner record per rank followed by the same
deterministic reduction on every rank (lowest cost, ties -> lowest global row): semantically an
all-reduce, one collective call, latency bound over NVLink/NVSwitch.

Two transports.  On one node (the default) the ranks share ONE page of pinned host memory (:class:`SharedPageExchange`):
the last CTA of every plan stores its rank's record into its slot of that page next to the result record -- a posted
PCIe write of one 64-byte line -- and every rank's host reads all slots after it has waited for its own plan.  No
collective kernel, no second device round trip: the payload is 16 bytes destined for the CPUs, so system memory is the
short way and NVLink would be a detour through a second GPU.  Across nodes (or on request) the exchange is the NCCL
all-gather of :class:`ArgminExchange`.

With the ``nccl`` backend the payload is the winner record in HBM itself (the device address the
library exposes through ``frx_winner_device_pointer``), so there is no host round trip before the
collective; with ``gloo`` (CPU tests) the record comes from the host result.
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np


def shard_rows(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [first, first + count) of rank `rank`."""
    per = -(-n_rows // world_size)
    first = min(rank * per, n_rows)
    count = max(0, min(per, n_rows - first))
    return first, count


def reduce_winners(costs: np.ndarray, rows: np.ndarray) -> Tuple[float, int]:
    """Deterministic arg-min over per-rank winners; row -1 marks "no candidate"."""
    best_c, best_r = float("inf"), -1
    for c, r in zip(costs.tolist(), rows.tolist()):
        if r < 0:
            continue
        if best_r < 0 or c < best_c or (c == best_c and r < best_r):
            best_c, best_r = c, int(r)
    return best_c, best_r


class _DevView:
    """Expose a raw device address to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ArgminExchange:
    """Reusable buffers for the per-plan exchange."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        self._gather = None
        self._local = None
        self._views = {}

    def enqueue(self, handler) -> None:
        """nccl only: queue the all-gather of `handler`'s winner record (and the read-back of the gathered
        records) behind the plan that is in flight on the current stream -- no host synchronisation."""
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._gather is None:
            self._gather = torch.empty(16 * self.world, dtype=torch.uint8, device=dev)
            self._host = torch.empty(16 * self.world, dtype=torch.uint8).pin_memory()
        ptr = handler.winner_device_pointer()
        v = self._views.get(ptr)
        if v is None:
            v = self._views[ptr] = torch.as_tensor(_DevView(ptr, 16), device=dev)
        dist.all_gather_into_tensor(self._gather, v, group=self.group)
        self._host.copy_(self._gather, non_blocking=True)

    def finish(self) -> Tuple[float, int, int]:
        """After the stream has been synchronised (frx_plan_wait does): reduce the gathered records."""
        raw = self._host.numpy().tobytes()
        recs = [struct.unpack_from("<dq", raw, 16 * r) for r in range(self.world)]
        costs = np.array([c for c, _ in recs])
        rows = np.array([r for _, r in recs], dtype=np.int64)
        c, r = reduce_winners(costs, rows)
        owner = int(np.nonzero(rows == r)[0][0]) if r >= 0 else -1
        return c, r, owner

    def exchange(self, min_cost: float, global_row: int, handler=None) -> Tuple[float, int, int]:
        """-> (global min cost, global row, owner rank).  `handler`: the _capi.Handler whose winner record
        should be sent straight from HBM (nccl only)."""
        if self.world == 1:
            return float(min_cost), int(global_row), 0
        torch, dist = self.torch, self.dist
        if self.backend == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            if self._gather is None:
                self._gather = torch.empty(16 * self.world, dtype=torch.uint8, device=dev)
                self._host = torch.empty(16 * self.world, dtype=torch.uint8).pin_memory()
            if handler is not None:
                ptr = handler.winner_device_pointer()
                v = self._views.get(ptr)
                if v is None:
                    v = torch.as_tensor(_DevView(ptr, 16), device=dev)
                    self._views[ptr] = v
                local = v
            else:
                local = torch.frombuffer(bytearray(struct.pack("<dq", float(min_cost), int(global_row))),
                                         dtype=torch.uint8).to(dev)
            dist.all_gather_into_tensor(self._gather, local, group=self.group)
            self._host.copy_(self._gather, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            raw = self._host.numpy().tobytes()
        else:
            local = torch.frombuffer(bytearray(struct.pack("<dq", float(min_cost), int(global_row))), dtype=torch.uint8)
            if self._gather is None:
                self._gather = torch.empty(16 * self.world, dtype=torch.uint8)
            dist.all_gather_into_tensor(self._gather, local, group=self.group)
            raw = self._gather.numpy().tobytes()
        recs = [struct.unpack_from("<dq", raw, 16 * r) for r in range(self.world)]
        costs = np.array([c for c, _ in recs])
        rows = np.array([r for _, r in recs], dtype=np.int64)
        c, r = reduce_winners(costs, rows)
        owner = int(np.nonzero(rows == r)[0][0]) if r >= 0 else -1
        return c, r, owner


def single_node() -> bool:
    """True when every rank of the default group runs on this node (torchrun exports LOCAL_WORLD_SIZE)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("LOCAL_WORLD_SIZE", str(world))) == world


class SharedPageExchange:
    """Arg-min exchange through one page of POSIX shared memory registered with every rank's CUDA context
    (``frx_set_exchange`` / ``frx_exchange_wait``, include/frx.h).  torch.distributed is only used once, to agree on the
    name of the page; the per-plan exchange involves no collective."""

    def __init__(self, handler, group=None):
        import ctypes
        import mmap
        import os
        import uuid
        import torch.distributed as dist
        from . import _capi
        self.handler = handler
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        name = [f"/dev/shm/frx_xchg_{uuid.uuid4().hex}" if self.rank == 0 else None]
        if self.world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
        self.path = name[0]
        if self.rank == 0:
            fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, _capi.EXCHANGE_PAGE_BYTES)                    # zero-filled by the kernel
        if self.world > 1:
            dist.barrier(group=group)
        if self.rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self._mm = mmap.mmap(fd, _capi.EXCHANGE_PAGE_BYTES)
        os.close(fd)
        self._keep = ctypes.c_char.from_buffer(self._mm)
        handler.set_exchange(ctypes.addressof(self._keep), self.rank, self.world)
        if self.world > 1:
            dist.barrier(group=group)                                      # every rank attached before the first plan
        if self.rank == 0:
            os.unlink(self.path)                                           # the mappings keep the page alive
        self.bytes_per_rank_and_plan = 64

    def finish(self):
        """After the local plan has been waited for: (global min cost, global row, owner rank)."""
        return self.handler.exchange_wait()

    def close(self):
        if self._mm is not None:
            self.handler.set_exchange(None)
            del self._keep
            self._mm.close()
            self._mm = None
