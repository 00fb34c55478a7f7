"""Data-parallel plumbing: one process per GPU, graphs of the global batch sharded across ranks,
gradients averaged with ONE NCCL all-reduce per backward over a flat fp32 bucket (SURVEY section 8e,
"default" semantics = what DistributedDataParallel around the reference would compute: GraphNorm /
type-table statistics stay rank-local).  The payload is tiny (G 1.10 MB, D 63 KB) => latency-bound:
a single bucket per model, launched on the compute stream right after the last backward kernel.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist


class GradSync:
    """``sync(model)`` averages ``p.grad`` of every parameter over the ranks with a single all-reduce."""

    def __init__(self, world_size: int):
        self.world = world_size
        self._buckets: Dict[int, torch.Tensor] = {}

    def __call__(self, model: torch.nn.Module) -> None:
        if self.world <= 1:
            return
        params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.grad is not None]
        if not params:
            return
        total = sum(p.numel() for p in params)
        flat = self._buckets.get(id(model))
        if flat is None or flat.numel() != total:
            flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
            self._buckets[id(model)] = flat
        views, off = [], 0
        for p in params:
            views.append(flat[off: off + p.numel()].view_as(p))
            off += p.numel()
        torch._foreach_copy_(views, [p.grad for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / self.world)
        torch._foreach_copy_([p.grad for p in params], views)


def shard_ids(ids: List[int], rank: int, world: int) -> List[int]:
    """Contiguous shard of the global batch's graph list for this rank (balanced to +-1 graph)."""
    n = len(ids)
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return ids[lo:hi]
