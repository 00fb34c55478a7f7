"""The four ``TrainerHelper`` methods on the hot path (reference building_gan/src/trainer.py:291-443) and the per-batch
body of ``Trainer._train_each_epoch`` (trainer.py:459-503), for the drop-in models.

The reference's own ``trainer.py`` drives the drop-in models and ``Batch`` unchanged (tests/test_trainer_dropin.py runs the
UNMODIFIED reference ``TrainerHelper`` on this package's ``Batch``); that file cannot be imported on a machine without the
reference checkout, so two flavours of the same interface live here:

``ReferenceTrainerHelper``  the reference's semantics step for step - z and the gradient-penalty mixing factor drawn on the
    CPU generator and copied (trainer.py:298-299,470,484), ``.item()`` after every backward (:479,493), the per-building Python
    FAR loop over ``voxel_graph[gi]`` (:362-380) and the sklearn metrics with their device->host copies (:387-443).  This is
    what "trainer.py unchanged" costs on the kernels: ``bench.py`` times it as the ``dropin`` block.

``TrainerHelper``  a mix-in with the SAME method names and return values for a maintainer who may touch the trainer class
    (``class Trainer(building_gan_b200.trainer_helper.TrainerHelper, reference.Trainer)``): the FAR loop is one masked
    segment sum (``bg_segment_pool``), the metrics one ``[B,7,7]`` confusion-matrix kernel (``bg_segment_confusion``) with a
    single device->host copy for the floats the reference logs, the critic loss glue two fused launches.  Same numbers
    (tests/test_trainer_ops_gpu.py: equal to sklearn / the Python loop), no per-building syncs.

Both expect what the reference's ``TrainerHelper`` expects on ``self``: ``generator``, ``discriminator``,
``configuration`` (and, for ``train_batch``, ``optimizer_generator`` / ``optimizer_discriminator``).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import step as _step


class ReferenceTrainerHelper:
    """Restatement of trainer.py:291-443 (no kernels of its own: everything below the model calls is torch / sklearn,
    exactly what the reference executes around the models)."""

    generator = None
    discriminator = None
    configuration = None

    # trainer.py:291-316
    def _compute_gradient_penalty(self, local_graph, voxel_graph, label_soft: Tensor) -> Tensor:
        onehot = voxel_graph.types_onehot
        mix = torch.rand(onehot.shape[0], 1).to(label_soft.device)  # CPU generator, then copied (appendix C #6)
        x_hat = (mix * onehot + (1 - mix) * label_soft.squeeze(0)).requires_grad_(True)
        score = self.discriminator(local_graph, voxel_graph, x_hat.unsqueeze(0))
        (grad,) = torch.autograd.grad(outputs=score, inputs=x_hat, grad_outputs=torch.ones_like(score), create_graph=True,
                                      only_inputs=True)
        return ((grad.norm(dim=1) - 1) ** 2).mean() * self.configuration.LAMBDA_GP

    # trainer.py:318-332
    def _compute_discriminator_loss(self, local_graph, voxel_graph, label_hard: Tensor, label_soft: Tensor) -> Tensor:
        d_real = self.discriminator(local_graph, voxel_graph, voxel_graph.types_onehot.unsqueeze(0))
        d_fake = self.discriminator(local_graph, voxel_graph, label_hard)
        if self.configuration.USE_WGANGP:
            loss = d_fake.mean() - d_real.mean()
            loss += self._compute_gradient_penalty(local_graph, voxel_graph, label_soft)
            return loss
        return (F.binary_cross_entropy(d_fake, torch.zeros_like(d_fake)) + F.binary_cross_entropy(d_real, torch.ones_like(d_real)))

    # trainer.py:334-385
    def _compute_generator_loss(self, local_graph, voxel_graph, logits: Tensor, label_hard: Tensor) -> Tensor:
        cfg = self.configuration
        d_fake = self.discriminator(local_graph, voxel_graph, label_hard)
        adv = -d_fake.mean() if cfg.USE_WGANGP else F.binary_cross_entropy(d_fake, torch.ones_like(d_fake))
        adv *= cfg.LAMBDA_ADV
        ce = F.cross_entropy(logits, voxel_graph.type)
        ce *= cfg.LAMBDA_LABEL
        share_g = label_hard.squeeze(0).sum(dim=0) / voxel_graph.num_nodes
        share = voxel_graph.types_onehot.sum(dim=0) / voxel_graph.num_nodes
        r_main = F.mse_loss(share_g[:-2], share[:-2])
        r_main *= cfg.LAMBDA_RATIO
        r_void = F.mse_loss(share_g[-2:], share[-2:])
        r_void *= cfg.LAMBDA_RATIO_VOID
        return adv + r_main + ce + r_void + self._far_term(voxel_graph, label_hard)

    def _far_term(self, voxel_graph, label_hard: Tensor) -> Tensor:
        """trainer.py:357-383: one pass of Python per building over ``voxel_graph[gi]``; ``torch.tensor(list of 0-dim
        tensors)`` reads every element back to the host, so the term is a detached CPU constant (appendix C #5)."""
        cfg = self.configuration
        generated = label_hard.squeeze(0).argmax(dim=1)
        want, got, lo = [], [], 0
        for gi in range(voxel_graph.num_graphs):
            one = voxel_graph[gi]
            hi = lo + one.num_nodes
            dims = one.x[:, 3:6] * cfg.NORMALIZATION_FACTOR_DIMENSION
            used = dims[generated[lo:hi] != cfg.VOID]
            got.append((used[:, 1] * used[:, 2]).sum() / one.site_area[0])
            want.append(one.x[0][9])
            lo = hi
        far = F.mse_loss(torch.tensor(got), torch.tensor(want))
        far *= cfg.LAMBDA_FAR
        return far

    # trainer.py:387-443
    def _compute_metrics(self, voxel_graph, label_hard: Tensor):
        from sklearn import metrics
        avg = self.configuration.METRICS_AVERAGE
        target = voxel_graph.type.cpu()
        pred = label_hard.squeeze(0).argmax(dim=1).cpu()
        f1 = metrics.f1_score(target, pred, average=avg, zero_division=0)
        precision = metrics.precision_score(target, pred, average=avg, zero_division=0)
        recall = metrics.recall_score(target, pred, average=avg, zero_division=0)
        accuracy = metrics.accuracy_score(target, pred)
        per_building, lo = [], 0
        for gi in range(voxel_graph.num_graphs):
            one = voxel_graph[gi]
            hi = lo + one.num_nodes
            per_building.append(metrics.f1_score(one.type.cpu(), label_hard.squeeze(0)[lo:hi].argmax(dim=1).cpu(), average=avg,
                                                 zero_division=0))
            lo = hi
        assert lo == voxel_graph.num_nodes
        return f1, per_building, precision, recall, accuracy

    # trainer.py:459-503, the body of the loop over the train dataloader for ONE batch
    def train_batch(self, local_graph, voxel_graph, with_metrics: bool = True):
        """Returns (critic losses [N_CRITIC floats], generator loss float, metrics tuple or None)."""
        cfg = self.configuration
        local_graph, voxel_graph = local_graph.to(cfg.DEVICE), voxel_graph.to(cfg.DEVICE)
        assert [set(d) for d in local_graph.data_number] == [set(d) for d in voxel_graph.data_number]
        d_losses: List[float] = []
        for _ in range(cfg.N_CRITIC):
            with torch.no_grad():
                z = torch.randn(1, voxel_graph.num_nodes, cfg.Z_DIM).to(cfg.DEVICE)
                _, hard, soft = self.generator(local_graph, voxel_graph, z)
                hard, soft = hard.unsqueeze(0), soft.unsqueeze(0)
            self.optimizer_discriminator.zero_grad()
            d_loss = self._compute_discriminator_loss(local_graph, voxel_graph, hard, soft)
            d_loss.backward()
            d_losses.append(d_loss.item())
            self.optimizer_discriminator.step()
        z = torch.randn(1, voxel_graph.num_nodes, cfg.Z_DIM).to(cfg.DEVICE)
        logits, hard, soft = self.generator(local_graph, voxel_graph, z)
        hard = hard.unsqueeze(0)
        self.optimizer_generator.zero_grad()
        g_loss = self._compute_generator_loss(local_graph, voxel_graph, logits, hard)
        g_loss.backward()
        g_val = g_loss.item()
        self.optimizer_generator.step()
        return d_losses, g_val, (self._compute_metrics(voxel_graph, hard) if with_metrics else None)


class TrainerHelper(ReferenceTrainerHelper):
    """Mix-in: the same four methods with the host loops replaced by device segment ops."""

    def _compute_gradient_penalty(self, local_graph, voxel_graph, label_soft: Tensor) -> Tensor:
        return _step.gradient_penalty(self.discriminator, local_graph, voxel_graph, label_soft, self.configuration, rng="cpu")

    def _compute_discriminator_loss(self, local_graph, voxel_graph, label_hard: Tensor, label_soft: Tensor) -> Tensor:
        return _step.discriminator_loss(self.discriminator, local_graph, voxel_graph, label_hard, label_soft, self.configuration,
                                        rng="cpu")

    def _far_term(self, voxel_graph, label_hard: Tensor) -> Tensor:
        return _step.far_loss(voxel_graph, label_hard, self.configuration)

    def _compute_metrics(self, voxel_graph, label_hard: Tensor) -> Tuple[float, List[float], float, float, float]:
        f1, f1_each, prec, rec, acc = _step.compute_metrics(voxel_graph, label_hard, self.configuration)
        flat = torch.cat([torch.stack([f1, prec, rec, acc]), f1_each]).tolist()  # ONE device->host copy
        return flat[0], flat[4:], flat[1], flat[2], flat[3]
