"""Frequency sharding across ranks: frequencies are independent systems over the same H, Q, A.

One process per GPU.  Frequencies are dealt round-robin (``batch`` consecutive ones at a time; ``FEM3D`` deals
single frequencies) because the iteration count rises with frequency and contiguous blocks would leave the last
rank with the expensive end of the sweep.  Inside a rank the solver's slots refill continuously.  There is no collective on the solve path; the result rows are gathered once at the end.
"""
from __future__ import annotations

import os

import numpy as np


def owned_frequencies(n_freq, batch, rank, world):
    """Indices (ascending) of the frequencies rank ``rank`` of ``world`` solves."""
    if world <= 1:
        return np.arange(n_freq, dtype=np.int64)
    n_batches = (n_freq + batch - 1) // batch
    mine = np.arange(rank, n_batches, world, dtype=np.int64)
    idx = (mine[:, None] * batch + np.arange(batch, dtype=np.int64)[None, :]).ravel()
    return idx[idx < n_freq]


def deal_by_cost(cost, rank, world):
    """Indices (ascending) of the items rank ``rank`` takes when they are dealt one at a time in order of decreasing ``cost``:
    every rank gets the same share of the expensive ones."""
    cost = np.asarray(cost, dtype=np.float64)
    if world <= 1:
        return np.arange(len(cost), dtype=np.int64)
    order = np.argsort(-cost, kind="stable")
    return np.sort(order[rank::world]).astype(np.int64)


def dist_info():
    """(rank, world, local_rank) from torch.distributed if initialised, else from the torchrun environment
    only when a process group exists; a plain single process is (0, 1, 0)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))
    except ImportError:
        pass
    return 0, 1, 0


def same_node(group=None):
    """True when every rank of the group runs on this node (what a shared-memory work queue needs)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    return world > 1 and int(os.environ.get("LOCAL_WORLD_SIZE", "0")) == dist.get_world_size()


class SharedCounters:
    """``n`` int64 counters in POSIX shared memory (a file under /dev/shm mapped by every rank of the node): the common work queue
    of a sharded sweep.  ``System.sweep(..., shared_next=counters.address(k))`` claims systems with atomic fetch-and-adds on
    counter k, so a rank that runs out of work takes the next system instead of idling - still no collective on the solve path.
    Collective to create; the file is unlinked as soon as every rank has mapped it."""

    def __init__(self, n=256, group=None):
        import mmap
        import uuid

        import torch.distributed as dist
        rank = dist.get_rank(group)
        box = [None]
        if rank == 0:
            box[0] = f"/dev/shm/femder_b200_{os.getpid()}_{uuid.uuid4().hex}"
            with open(box[0], "wb") as f:
                f.write(b"\0" * (8 * n))
        _broadcast(box, group)
        self._f = open(box[0], "r+b")
        self._map = mmap.mmap(self._f.fileno(), 8 * n)
        self.values = np.frombuffer(self._map, dtype=np.int64)
        import ctypes
        self._base = ctypes.addressof(ctypes.c_char.from_buffer(self._map))
        self.n = n
        self._next = 0
        self._group = group
        dist.barrier(group)
        if rank == 0:
            os.unlink(box[0])

    def address(self, k):
        return self._base + 8 * int(k)

    def fresh(self):
        """Index of a zeroed counter, the same on every rank (collective: all ranks call it the same number of times)."""
        import torch.distributed as dist
        k = self._next % self.n
        self._next += 1
        if self._next > self.n:      # wrapped round: somebody may still be reading the old value
            dist.barrier(self._group)
            if dist.get_rank(self._group) == 0:
                self.values[k] = 0
            dist.barrier(self._group)
        return k


_COUNTERS = None


def counters():
    """The process group's counters (created collectively on first use)."""
    global _COUNTERS
    if _COUNTERS is None:
        _COUNTERS = SharedCounters()
    return _COUNTERS


def _broadcast(box, group=None):
    import torch
    import torch.distributed as dist
    kw = {}
    if dist.get_backend(group) == "nccl":
        kw["device"] = torch.device("cuda", torch.cuda.current_device())
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group, **kw)


def share_modal_setup(system, group=None):
    """Make the ranks of the process group one communicator of ``system``'s library (NCCL, created by the library itself from an id
    rank 0 makes and torch.distributed hands round): ``System.modal_compute`` then deals the set-up's columns to the ranks.
    The one exchange on the path - every rank needs the same room modes; the sweep that follows has no collective."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world <= 1 or system.comm_world == world:
        return
    box = [system.comm_unique_id() if rank == 0 else None]
    kw = {}
    if dist.get_backend(group) == "nccl":
        kw["device"] = torch.device("cuda", system.device)
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group, **kw)
    system.comm_init(box[0], rank, world)


def gather_rows(local_rows, owned, n_freq, group=None):
    """All ranks contribute ``local_rows`` (len(owned), ...) and get back the full (n_freq, ...) array.
    Host-side gather of python objects: the payload is receiver pressures (small) or, on request, pN rows."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(owned), None if local_rows is None else np.asarray(local_rows)), group=group)
    first = next((p[1] for p in parts if p[1] is not None and len(p[0])), None)
    if first is None:
        return None
    out = np.zeros((n_freq,) + first.shape[1:], dtype=first.dtype)
    seen = np.zeros(n_freq, dtype=bool)
    for idx, rows in parts:
        if rows is None or len(idx) == 0:
            continue
        out[idx] = rows
        seen[idx] = True
    if not seen.all():
        raise RuntimeError("frequency sharding left rows unsolved")
    return out
