"""Data-parallel plumbing: one process per GPU, graphs of the global batch sharded across ranks, gradients averaged
with ONE all-reduce per backward over a flat fp32 bucket ("default" semantics = what
DistributedDataParallel around the reference would compute: GraphNorm / type-table statistics and the batch-mean
losses stay rank-local).  The payload is tiny (G 1.10 MB, D 63 KB) => latency-bound: a single bucket per model, issued
on the compute stream right after the last backward kernel.  In BG_GRADS=bucket mode the bucket IS where the kernels
accumulated the gradients (every ``p.grad`` is a view of it), so there is no flatten / unflatten copy at all.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


class GradSync:
    """``sync(model)`` averages the parameter gradients of ``model`` over the ranks with a single all-reduce."""

    def __init__(self, world_size: int):
        self.world = world_size
        self._buckets: Dict[int, torch.Tensor] = {}

    def __call__(self, model: torch.nn.Module) -> None:
        if self.world <= 1:
            return
        native = getattr(model, "_native", None)
        bucket = getattr(native, "bucket", None)
        params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.grad is not None]
        if not params:
            return
        if bucket is not None and params[0].grad.data_ptr() == native.views[0].data_ptr():
            dist.all_reduce(bucket, op=dist.ReduceOp.AVG)  # grads already live in one flat bucket; 1/world folded into the reduction
            return
        total = sum(p.numel() for p in params)
        flat = self._buckets.get(id(model))
        if flat is None or flat.numel() != total or flat.device != params[0].device:
            flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
            self._buckets[id(model)] = flat
        views, off = [], 0
        for p in params:
            views.append(flat[off: off + p.numel()].view_as(p))
            off += p.numel()
        torch._foreach_copy_(views, [p.grad for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM)
        if dist.get_backend() != "nccl":  # gloo has no AVG
            flat.mul_(1.0 / self.world)
        torch._foreach_copy_([p.grad for p in params], views)


class PeerSync:
    """Gradient exchange FUSED with the optimiser step over NVLink peer memory (``bg_p2p_allreduce_adam``, csrc/bg_p2p.cu):
    ``attach(model, optimizer)`` moves the model's flat gradient bucket into symmetric memory
    (torch.distributed._symmetric_memory) and tells the flat Adam (``optim.Adam``) to run the one-launch
    "read every rank's bucket, average, Adam" kernel instead of ncclAllReduce + scale + Adam.  Calling the object like a
    ``GradSync`` (``sync(model)``) is then a no-op for attached models - the exchange happens inside ``optimizer.step()`` -
    and falls back to the NCCL all-reduce for anything else."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self._attached: Dict[int, object] = {}
        self._nccl = GradSync(self.world)
        self._keep = []  # symmetric-memory handles must outlive the tensors' use

    def attach(self, model, optimizer) -> None:
        import torch.distributed._symmetric_memory as symm
        from . import lib
        from .models import _param_list
        from .optim import Adam
        if not isinstance(optimizer, Adam):
            raise TypeError("PeerSync.attach: the fused exchange needs building_gan_b200.optim.Adam (flat buffers)")
        if self.world > lib.MAX_PEERS:
            raise ValueError(f"PeerSync: at most {lib.MAX_PEERS} ranks (one NVSwitch domain), got {self.world}")
        st, params = model._native, _param_list(model)
        dev = params[0].device
        name = self.group.group_name
        bucket = symm.empty(st.layout.total, dtype=torch.float32, device=dev)
        bucket.zero_()
        flags = symm.empty(2 * lib.MAX_PEERS, dtype=torch.int32, device=dev)
        flags.zero_()
        hb, hf = symm.rendezvous(bucket, name), symm.rendezvous(flags, name)
        torch.cuda.synchronize(dev)
        hb.barrier()  # every rank's zeroing is complete before anybody signals
        peers = lib.BgPeers()
        peers.rank, peers.world = self.rank, self.world
        for r in range(self.world):
            peers.grad[r], peers.flags[r] = hb.buffer_ptrs[r], hf.buffer_ptrs[r]
        st.install_bucket(bucket, params)
        optimizer._peer = (peers, torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
        self._attached[id(model)] = optimizer
        self._keep += [bucket, flags, hb, hf]

    def __call__(self, model) -> None:
        if id(model) in self._attached:
            return  # exchanged inside optimizer.step()
        self._nccl(model)


def shard_ids(ids: Sequence[int], rank: int, world: int) -> List[int]:
    """Contiguous shard of the global batch's graph list for this rank (sizes differ by at most one graph)."""
    n = len(ids)
    return list(ids[(n * rank) // world: (n * (rank + 1)) // world])


def shard_by_nodes(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Contiguous partition of a batch's graphs into ``world`` shards balanced by voxel count (greedy prefix split):
    returns the graph indices of every shard.  Every shard holds at least one graph (a rank with an empty shard could not
    build a batch and would miss the gradient all-reduce); fewer graphs than ranks is an error."""
    if world < 1:
        raise ValueError(f"shard_by_nodes: world = {world}")
    if len(sizes) < world:
        raise ValueError(f"shard_by_nodes: {len(sizes)} graphs cannot be split over {world} ranks (every rank needs >= 1 graph)")
    total, out, acc, cur, r = sum(sizes), [], 0, [], 0
    for i, s in enumerate(sizes):
        cur.append(i)
        acc += s
        remaining_graphs = len(sizes) - i - 1
        shards_after = world - r - 1  # shards still to be opened after the current one
        if r < world - 1 and (acc >= total * (r + 1) / world or remaining_graphs == shards_after):
            out.append(cur)
            cur, r = [], r + 1
    out.append(cur)
    assert len(out) == world and all(out), (sizes, world, out)
    return out
