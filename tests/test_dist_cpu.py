"""Multi-process host logic of the data-parallel path on CPU (gloo, world_size 2): graph sharding and the flat-bucket
gradient all-reduce.  The kernels themselves need a GPU; what is checked here is the plumbing around them: N ranks on
shards + averaged gradients == one process that runs every shard and averages (the DDP-equivalent semantics)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from building_gan_b200 import dist as bdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))


def _shard_loss(model, ids):
    g = torch.Generator().manual_seed(100 + sum(ids))
    x = torch.randn(4 * len(ids), 6, generator=g)
    return model(x).mean()


def _worker(rank, world, port, ids, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _make_model()
    mine = bdist.shard_ids(ids, rank, world)
    _shard_loss(model, mine).backward()
    bdist.GradSync(world)(model)
    if rank == 0:
        torch.save([p.grad.clone() for p in model.parameters()], out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_sync_equals_single_process_average(tmp_path):
    ids, world, out = list(range(7)), 2, str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(world, _free_port(), ids, out), nprocs=world, join=True)
    got = torch.load(out)
    ref_model = _make_model()
    acc = [torch.zeros_like(p) for p in ref_model.parameters()]
    for r in range(world):
        ref_model.zero_grad()
        _shard_loss(ref_model, bdist.shard_ids(ids, r, world)).backward()
        for a, p in zip(acc, ref_model.parameters()):
            a += p.grad / world
    for g, a in zip(got, acc):
        assert torch.allclose(g, a, rtol=1e-6, atol=1e-7)


def test_shard_ids_cover_and_balance():
    for n in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            shards = [bdist.shard_ids(list(range(n)), r, world) for r in range(world)]
            assert sum(shards, []) == list(range(n))
            assert max(map(len, shards)) - min(map(len, shards)) <= 1


def test_shard_by_nodes_is_contiguous_and_balanced():
    sizes = [400, 120, 900, 300, 310, 1500, 200, 220, 640, 90]
    for world in (2, 4):
        shards = bdist.shard_by_nodes(sizes, world)
        assert sum(shards, []) == list(range(len(sizes))) and len(shards) == world
        loads = [sum(sizes[i] for i in s) for s in shards]
        assert max(loads) <= sum(sizes) / world + max(sizes)


@pytest.mark.parametrize("sizes,world", [([1, 1, 100], 3), ([100, 1, 1], 3), ([1, 1, 100], 2), ([5] * 8, 4), ([400, 10, 10, 10, 10, 400], 4),
                                         ([180, 378, 450, 504] * 8, 8), ([7], 1), ([3, 900, 2, 2], 4)])
def test_shard_by_nodes_never_leaves_a_rank_without_graphs(sizes, world):
    shards = bdist.shard_by_nodes(sizes, world)
    assert len(shards) == world and all(len(s) >= 1 for s in shards)
    assert [i for s in shards for i in s] == list(range(len(sizes)))  # contiguous, order kept, every graph exactly once


def test_shard_by_nodes_balances_by_voxel_count():
    sizes = [180, 378, 450, 504] * 64  # a batch-256 worth of buildings
    shards = bdist.shard_by_nodes(sizes, 8)
    load = [sum(sizes[i] for i in s) for s in shards]
    assert max(load) - min(load) <= max(sizes)


def test_shard_by_nodes_refuses_fewer_graphs_than_ranks():
    with pytest.raises(ValueError, match="cannot be split"):
        bdist.shard_by_nodes([10, 20], 4)
