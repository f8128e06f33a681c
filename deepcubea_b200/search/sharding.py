"""Multi-GPU plumbing: the search shards by independent problem instance (the reference solves its start states
strictly one after the other, search_methods/astar.py:508-520, each with private OPEN/CLOSED), so instance i goes to
rank i mod N and nothing crosses GPUs on the data path.  torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests) is used only for the completion barrier, the (time MAX, nodes SUM) reduction and the result gather.
"""
from __future__ import annotations

import os
from typing import Any, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: instance i -> rank i % world_size."""
    return list(range(rank, n_items, world_size))


class InstanceQueue:
    """Dynamic whole-instance queue across ranks: every rank draws the next unsolved instance with an atomic fetch-add on the job's
    c10d store (host side, ~0.1 ms per draw, nothing on the data path).  Cube3 searches span 10^3 ... > 10^8 nodes (load
    imbalance between instances is THE scaling loss), so a rank that draws easy instances simply takes more of them."""

    def __init__(self, n_items: int, key: str = "dcb_instance_queue", group=None):
        self.n, self.local, self.key = n_items, 0, key
        self.store = None
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            self.store = dist.distributed_c10d._get_default_store()

    def next(self) -> Optional[int]:
        if self.store is None:
            i = self.local
            self.local += 1
        else:
            i = int(self.store.add(self.key, 1)) - 1
        return i if i < self.n else None

    def __iter__(self):
        while True:
            i = self.next()
            if i is None:
                return
            yield i


def merge_sharded(per_rank: Sequence[Sequence[Tuple[int, Any]]], n_items: int) -> List[Any]:
    """Inverse of shard_indices: [(global index, result)] lists from every rank -> results in input order."""
    out: List[Any] = [None] * n_items
    seen = 0
    for chunk in per_rank:
        for idx, res in chunk:
            if out[idx] is not None:
                raise ValueError("instance %d reported by two ranks" % idx)
            out[idx] = res
            seen += 1
    if seen != n_items:
        raise ValueError("expected %d results, got %d" % (n_items, seen))
    return out


def gather_results(local: Sequence[Tuple[int, Any]], n_items: int, group=None) -> Optional[List[Any]]:
    """Collect every rank's (index, result) pairs on rank 0 (KBs: moves, node counts, times); None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return merge_sharded([list(local)], n_items)
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    bucket: Optional[List[Any]] = [None] * ws if rank == 0 else None
    dist.gather_object(list(local), bucket, dst=0, group=group)
    return merge_sharded(bucket, n_items) if rank == 0 else None


def reduce_throughput(nodes: float, seconds: float, device: torch.device, group=None) -> Tuple[float, float]:
    """Whole-job numbers: nodes summed over ranks, time = max over ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(nodes), float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    n = torch.tensor([nodes], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return float(n.item()), float(t.item())


def completion_barrier(group=None) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)
