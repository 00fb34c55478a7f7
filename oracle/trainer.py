"""ORACLE (test infrastructure): restatement of the loss / optimisation-step glue that drives the
hot path, ``building_gan/src/trainer.py:291-385`` (gradient penalty, critic loss, generator loss)
and ``:459-495`` (one step = N_CRITIC critic updates + one generator update).

PINNED by tests/test_oracle_golden.py against vectors recorded from the unmodified reference
``TrainerHelper`` methods (oracle/make_golden.py).  RNG draws happen in the reference's order and
on the reference's devices (z and the GP mixing factor come from the CPU generator,
trainer.py:298,470,484).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


def gradient_penalty(discriminator, local_graph, voxel_graph, label_soft: Tensor, cfg, e: Optional[Tensor] = None) -> Tensor:
    """trainer.py:291-316.  ``e`` (optional): the mixing factor, injected by parity checks that run this restatement in
    fp64 (the reference always draws it: ``torch.rand(N, 1)`` on the CPU generator, float32)."""
    if e is None:
        e = torch.rand(voxel_graph.types_onehot.shape[0], 1)
    e = e.to(label_soft.device)
    mixed = (e * voxel_graph.types_onehot + (1 - e) * label_soft.squeeze(0)).requires_grad_(True)
    score = discriminator(local_graph, voxel_graph, mixed.unsqueeze(0))
    (grad,) = torch.autograd.grad(score, mixed, torch.ones_like(score), create_graph=True, only_inputs=True)
    return ((grad.norm(dim=1) - 1) ** 2).mean() * cfg.LAMBDA_GP


def discriminator_loss(discriminator, local_graph, voxel_graph, label_hard: Tensor, label_soft: Tensor, cfg,
                       e: Optional[Tensor] = None) -> Tensor:
    """trainer.py:318-332 (label_* carry the leading unsqueeze(0) the trainer adds)."""
    d_real = discriminator(local_graph, voxel_graph, voxel_graph.types_onehot.unsqueeze(0))
    d_fake = discriminator(local_graph, voxel_graph, label_hard)
    if cfg.USE_WGANGP:
        loss = d_fake.mean() - d_real.mean()
        loss = loss + gradient_penalty(discriminator, local_graph, voxel_graph, label_soft, cfg, e)
        return loss
    return F.binary_cross_entropy(d_fake, torch.zeros_like(d_fake)) + F.binary_cross_entropy(
        d_real, torch.ones_like(d_real)
    )


def far_loss(voxel_graph, label_hard: Tensor, cfg) -> Tensor:
    """trainer.py:357-381: per-graph generated floor-area ratio vs the recorded one; built through
    ``torch.tensor(list)`` in the reference, i.e. a detached CPU constant."""
    generated = label_hard.squeeze(0).argmax(dim=1)
    want, got, lo = [], [], 0
    for gi in range(voxel_graph.num_graphs):
        g = voxel_graph[gi]
        hi = lo + g.num_nodes
        dims = g.x[:, 3:6] * cfg.NORMALIZATION_FACTOR_DIMENSION
        used = dims[generated[lo:hi] != cfg.VOID]
        got.append((used[:, 1] * used[:, 2]).sum() / g.site_area[0])
        want.append(g.x[0][9])
        lo = hi
    return F.mse_loss(torch.tensor(got), torch.tensor(want)) * cfg.LAMBDA_FAR


def generator_loss(discriminator, local_graph, voxel_graph, logits: Tensor, label_hard: Tensor, cfg) -> Tensor:
    """trainer.py:334-385."""
    d_fake = discriminator(local_graph, voxel_graph, label_hard)
    if cfg.USE_WGANGP:
        adv = -d_fake.mean()
    else:
        adv = F.binary_cross_entropy(d_fake, torch.ones_like(d_fake))
    adv = adv * cfg.LAMBDA_ADV
    ce = F.cross_entropy(logits, voxel_graph.type) * cfg.LAMBDA_LABEL
    n = voxel_graph.num_nodes
    ratio_g = label_hard.squeeze(0).sum(dim=0) / n
    ratio = voxel_graph.types_onehot.sum(dim=0) / n
    r_main = F.mse_loss(ratio_g[:-2], ratio[:-2]) * cfg.LAMBDA_RATIO
    r_void = F.mse_loss(ratio_g[-2:], ratio[-2:]) * cfg.LAMBDA_RATIO_VOID
    return adv + r_main + ce + r_void + far_loss(voxel_graph, label_hard, cfg)


def train_step(generator, discriminator, opt_g, opt_d, local_graph, voxel_graph, cfg) -> Tuple[List[float], float, Tensor]:
    """trainer.py:467-495 for one (already device-resident) batch.  Returns the N_CRITIC critic
    losses, the generator loss and the final label_hard[1,N,7]."""
    dev = voxel_graph.x.device
    d_losses: List[float] = []
    for _ in range(cfg.N_CRITIC):
        with torch.no_grad():
            z = torch.randn(1, voxel_graph.num_nodes, cfg.Z_DIM).to(dev)
            _, hard, soft = generator(local_graph, voxel_graph, z)
            hard, soft = hard.unsqueeze(0), soft.unsqueeze(0)
        opt_d.zero_grad()
        d_loss = discriminator_loss(discriminator, local_graph, voxel_graph, hard, soft, cfg)
        d_loss.backward()
        d_losses.append(d_loss.item())
        opt_d.step()
    z = torch.randn(1, voxel_graph.num_nodes, cfg.Z_DIM).to(dev)
    logits, hard, soft = generator(local_graph, voxel_graph, z)
    hard = hard.unsqueeze(0)
    opt_g.zero_grad()
    g_loss = generator_loss(discriminator, local_graph, voxel_graph, logits, hard, cfg)
    g_loss.backward()
    g_val = g_loss.item()
    opt_g.step()
    return d_losses, g_val, hard.detach()
