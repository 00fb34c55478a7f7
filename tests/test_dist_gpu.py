"""Multi-rank data-parallel path ON THE GPUS (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`;
skipped on a one-GPU box).  Two ranks on different shards of one global batch, real kernels:

* the fused peer-memory exchange (bg_p2p_allreduce_adam: one-shot all-reduce over NVLink + Adam in one launch) equals
  ncclAllReduce(AVG) + bg_adam_flat BIT FOR BIT at world size 2 (a + b is commutative, the halving exact), and equals the
  single-process reference "run every shard, average the gradients, step once" to fp32 rounding;
* a graphed training step (graphs.GraphedStep) with the fused exchange keeps the ranks' parameters bit-identical.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shards(world):
    from building_gan_b200 import graph, synth
    from building_gan_b200.dist import shard_by_nodes
    from workloads import synth as wsynth
    ids = list(range(61, 61 + 4 * world))
    parts = shard_by_nodes([wsynth.voxel_count(i) for i in ids], world)
    return [graph.collate_fn([synth.building_pair_fast(ids[k]) for k in part]) for part in parts]


def _models(dev, seed=3):
    from building_gan_b200 import Configuration
    from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
    cfg = Configuration()
    torch.manual_seed(seed)
    return cfg, VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)


def _critic_grads(cfg, G, D, lb, vb, seed):
    """One critic loss + backward on a shard (eval-mode dropout: deterministic); leaves the gradients in D's bucket."""
    from building_gan_b200 import step as bstep
    G.eval(), D.eval()
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(1, vb.num_nodes, cfg.Z_DIM, generator=gen).to(vb.x.device)
    noise = -torch.empty(vb.num_nodes, cfg.NUM_CLASSES).exponential_(generator=gen).log().to(vb.x.device)
    e = torch.rand(vb.num_nodes, 1, generator=gen).to(vb.x.device)
    with torch.no_grad():
        _, hard, soft = G(lb, vb, z, noise)
    D.zero_grad()
    loss = bstep.discriminator_loss(D, lb, vb, hard.unsqueeze(0), soft.unsqueeze(0), cfg, rng="cpu", e=e)
    loss.backward()
    return loss.detach()


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from building_gan_b200.dist import GradSync, PeerSync
    from building_gan_b200.optim import Adam
    result = {}
    shards = _shards(world)
    lb, vb = shards[rank]
    lb, vb = lb.to(dev), vb.to(dev)
    # ---- A: NCCL all-reduce (AVG) + one-launch Adam
    cfg, G, D = _models(dev)
    opt = Adam(D.parameters(), lr=1e-3, betas=cfg.BETAS)
    _critic_grads(cfg, G, D, lb, vb, 100 + rank)
    GradSync(world)(D)
    avg_nccl = D._native.bucket.clone()
    opt.step()
    p_nccl = D._native.pflat.clone()
    # ---- B: fused peer-memory exchange + Adam, same weights / shards / draws
    cfg, G2, D2 = _models(dev)
    opt2 = Adam(D2.parameters(), lr=1e-3, betas=cfg.BETAS)
    sync = PeerSync()
    sync.attach(D2, opt2)
    _critic_grads(cfg, G2, D2, lb, vb, 100 + rank)
    sync(D2)  # no-op for an attached model
    opt2.step()
    p_fused = D2._native.pflat.clone()
    result["fused_equals_nccl"] = bool(torch.equal(p_nccl, p_fused))
    result["max_diff_fused_nccl"] = float((p_nccl - p_fused).abs().max())
    # ---- C: single-process reference on THIS rank: every shard, gradients averaged, one step
    cfg, G3, D3 = _models(dev)
    opt3 = Adam(D3.parameters(), lr=1e-3, betas=cfg.BETAS)
    acc = None
    for r, (l, v) in enumerate(shards):
        _critic_grads(cfg, G3, D3, l.to(dev), v.to(dev), 100 + r)
        b = D3._native.bucket.clone()
        acc = b if acc is None else acc + b
    D3._native.bucket.copy_(acc / world)
    result["avg_grad_rel_err_vs_single_process"] = float((avg_nccl - D3._native.bucket).abs().max() / D3._native.bucket.abs().max())
    opt3.step()
    result["param_max_diff_vs_single_process"] = float((D3._native.pflat - p_fused).abs().max())
    # ---- D: a few graphed training steps with the fused exchange for both models: ranks stay bit-identical
    from building_gan_b200.graphs import GraphedStep
    cfg, G4, D4 = _models(dev, seed=9)
    G4.train(), D4.train()
    og, od = Adam(G4.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D4.parameters(), lr=2e-4, betas=cfg.BETAS)
    sync4 = PeerSync()
    sync4.attach(D4, od)
    sync4.attach(G4, og)
    gs = GraphedStep(G4, D4, og, od, cfg, grad_sync=sync4)
    torch.manual_seed(1234 + rank)
    for _ in range(3):
        d_losses, g_loss, _ = gs(lb, vb, sync_losses="step")
    torch.cuda.synchronize()
    sig = torch.stack([G4._native.pflat.double().sum(), G4._native.pflat.double().abs().sum(), D4._native.pflat.double().sum(),
                       D4._native.pflat.double().abs().sum()])
    gathered = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(gathered, sig)
    result["ranks_identical_after_graphed_steps"] = all(torch.equal(gathered[0], g) for g in gathered)
    result["losses_finite"] = bool(all(abs(x) < 1e6 for x in d_losses + [g_loss]))
    if rank == 0:
        torch.save(result, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_fused_exchange_and_graphed_step(tmp_path):
    out = str(tmp_path / "result.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    print(r)
    assert r["fused_equals_nccl"], r
    assert r["avg_grad_rel_err_vs_single_process"] <= 1e-6, r
    assert r["param_max_diff_vs_single_process"] <= 2e-6, r   # Adam's first step moves every parameter by ~lr = 1e-3
    assert r["ranks_identical_after_graphed_steps"] and r["losses_finite"], r
