"""One optimisation step of the reference trainer on the B200 path.

Mirrors ``TrainerHelper`` (reference building_gan/src/trainer.py): ``gradient_penalty`` (:291-316),
``discriminator_loss`` (:318-332), ``generator_loss`` (:334-385) and the body of
``_train_each_epoch`` (:467-495, N_CRITIC critic updates + one generator update).  The reference's
own ``trainer.py`` can drive the drop-in models unchanged; this module exists because (a)
/root/reference is not importable at run time and (b) the reference's per-graph Python FAR loop
(``voxel_graph[gi]`` B times with D2H syncs, trainer.py:362-380) is replaced by a sync-free
per-graph segment sum that returns the same number.

``rng="cpu"`` draws z and the GP mixing factor on the CPU generator exactly like the reference
(trainer.py:298,470,484: bit-identical stream, one H2D per draw); ``rng="device"`` draws them on the
GPU (same distributions, no host round trip) - the fast default of the benchmark.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import lib


def _rand(shape, device, rng: str, normal: bool) -> Tensor:
    if rng == "cpu":
        t = torch.randn(*shape) if normal else torch.rand(*shape)
        return t.to(device, non_blocking=True)
    return torch.randn(*shape, device=device) if normal else torch.rand(*shape, device=device)


def gradient_penalty(discriminator, local_graph, voxel_graph, label_soft: Tensor, cfg, rng: str = "cpu",
                     e: Optional[Tensor] = None) -> Tensor:
    n = voxel_graph.types_onehot.shape[0]
    if e is None:
        e = _rand((n, 1), label_soft.device, rng, normal=False)
    mixed = (e * voxel_graph.types_onehot + (1 - e) * label_soft.squeeze(0)).requires_grad_(True)
    score = discriminator(local_graph, voxel_graph, mixed.unsqueeze(0))
    (grad,) = torch.autograd.grad(score, mixed, torch.ones_like(score), create_graph=True, only_inputs=True)
    return ((grad.norm(dim=1) - 1) ** 2).mean() * cfg.LAMBDA_GP


def stream_priorities():
    """(gradient-penalty chain, D(real) / D(fake) lanes, sampling stream); BG_PRIO="a,b,c" overrides (experiments).  The library's
    weight-gradient side streams run at priority 0."""
    env = os.environ.get("BG_PRIO")
    return tuple(int(v) for v in env.split(",")) if env else (-2, -1, -1)


class Lanes:
    """Side streams of the overlapped training step (``train_step(..., overlap=True)``), one set per device.

    At the reference's batch size every kernel of the step is latency-bound (N ~ 15 k voxels: tens of CTAs on 148 SMs), and
    the step contains independent work the reference runs back to back: the three critic passes D(real), D(fake), D(mixed)
    of one update (trainer.py:319-323) and the no-grad generator passes of all N_CRITIC updates (trainer.py:469-473: G does
    not change inside the critic loop).  Each gets its own stream: ``real`` / ``fake`` for the two plain critic passes (forward
    and, through autograd's stream affinity, backward), ``gen`` for the sampling passes; the gradient-penalty pass and
    everything else stay on the caller's stream.  Same arithmetic, same kernels; only the parameter-gradient contributions
    of the three critic passes are summed in a different order (lane buckets, models._NativeState)."""
    _per_device = {}

    def __init__(self, device):
        pr = stream_priorities()
        self.real, self.fake = (torch.cuda.Stream(device=device, priority=pr[1]) for _ in range(2))
        self.gen = torch.cuda.Stream(device=device, priority=pr[2])

    @classmethod
    def get(cls, device) -> "Lanes":
        key = torch.device(device).index
        if key not in cls._per_device:
            cls._per_device[key] = cls(device)
        return cls._per_device[key]


@contextlib.contextmanager
def lanes_pdl_scope():
    """Programmatic dependent launch and event-forked streams do not mix (csrc/bg_misc.cu: 26.3 vs 15.5 ms per step), so
    the overlapped step switches PDL off for ITS launches only and restores the previous setting on exit (an explicit
    ``BG_PDL`` in the environment is left alone)."""
    if os.environ.get("BG_PDL") is not None:
        yield
        return
    prev = lib.set_pdl(False)
    try:
        yield
    finally:
        lib.set_pdl(prev)


class _CriticLossFn(torch.autograd.Function):
    """loss = mean(D(fake)) - mean(D(real)) + lambda * mean((||grad_i||_2 - 1)^2)  (trainer.py:314,323) as one forward and one
    backward launch (bg_critic_loss_fwd / _bwd) instead of ~25 torch launches on the critical path of every critic update."""

    @staticmethod
    def forward(ctx, d_fake, d_real, grad, lam):
        out4, coef = lib.critic_loss_fwd(d_fake.contiguous(), d_real.contiguous(), grad.contiguous(), lam)
        ctx.save_for_backward(coef, grad)
        return out4[0]

    @staticmethod
    def backward(ctx, g):
        coef, grad = ctx.saved_tensors
        g_fake, g_real, g_grad = lib.critic_loss_bwd(g.contiguous(), coef, grad.contiguous(), *ctx.needs_input_grad[:3])
        return g_fake, g_real, g_grad, None


def _penalty_gradient(discriminator, local_graph, voxel_graph, label_soft: Tensor, rng: str, e: Optional[Tensor]) -> Tensor:
    """trainer.py:298-312: the critic's input gradient at the interpolate (differentiable: create_graph=True)."""
    n = voxel_graph.types_onehot.shape[0]
    if e is None:
        e = _rand((n, 1), label_soft.device, rng, normal=False)
    mixed = lib.gp_mix(e.contiguous(), voxel_graph.types_onehot.contiguous(), label_soft.squeeze(0).contiguous()).requires_grad_(True)
    score = discriminator(local_graph, voxel_graph, mixed.unsqueeze(0))
    (grad,) = torch.autograd.grad(score, mixed, torch.ones_like(score), create_graph=True, only_inputs=True)
    return grad


def discriminator_loss(discriminator, local_graph, voxel_graph, label_hard: Tensor, label_soft: Tensor, cfg,
                       rng: str = "cpu", lanes: Optional[Lanes] = None, e: Optional[Tensor] = None) -> Tensor:
    """trainer.py:318-332.  WGAN-GP on the kernel path: the interpolate and the loss arithmetic are fused launches
    (bg_gp_mix, bg_critic_loss_*); ``gradient_penalty`` above is the same term spelled with torch ops, as the reference does."""
    real = voxel_graph.types_onehot.unsqueeze(0)
    fused = cfg.USE_WGANGP and label_soft.is_cuda and hasattr(discriminator, "_native")
    if lanes is None or not fused or not hasattr(discriminator, "lane"):
        d_real = discriminator(local_graph, voxel_graph, real)
        d_fake = discriminator(local_graph, voxel_graph, label_hard)
        if not cfg.USE_WGANGP:  # trainer.py:326-330 (the discriminator then ends in a sigmoid, models.py:222-223)
            return F.binary_cross_entropy(d_fake, torch.zeros_like(d_fake)) + F.binary_cross_entropy(d_real, torch.ones_like(d_real))
        if not fused:
            return d_fake.mean() - d_real.mean() + gradient_penalty(discriminator, local_graph, voxel_graph, label_soft, cfg, rng, e)
        grad = _penalty_gradient(discriminator, local_graph, voxel_graph, label_soft, rng, e)
        return _CriticLossFn.apply(d_fake, d_real, grad, cfg.LAMBDA_GP)
    # same three passes in the same host order, on three streams: fork after everything enqueued so far, join before the loss
    main = torch.cuda.current_stream()
    fork = main.record_event()
    with torch.cuda.stream(lanes.real), discriminator.lane(1):
        lanes.real.wait_event(fork)
        d_real = discriminator(local_graph, voxel_graph, real)
    with torch.cuda.stream(lanes.fake), discriminator.lane(2):
        lanes.fake.wait_event(fork)
        d_fake = discriminator(local_graph, voxel_graph, label_hard)
    grad = _penalty_gradient(discriminator, local_graph, voxel_graph, label_soft, rng, e)
    main.wait_stream(lanes.real)
    main.wait_stream(lanes.fake)
    return _CriticLossFn.apply(d_fake, d_real, grad, cfg.LAMBDA_GP)


def far_loss(voxel_graph, label_hard: Tensor, cfg) -> Tensor:
    """trainer.py:357-381 without the per-graph Python loop: generated floor area per building by a segment
    sum over ``Batch.ptr`` (bg_segment_pool), then MSE against the recorded FAR.  Detached, like the
    reference's ``torch.tensor(list)`` construction."""
    with torch.no_grad():
        lab = label_hard.squeeze(0)
        x = voxel_graph.x
        solid = (lab.argmax(dim=1) != cfg.VOID).to(torch.float32)
        dims = x[:, 3:6] * cfg.NORMALIZATION_FACTOR_DIMENSION
        area = (dims[:, 1] * dims[:, 2] * solid).unsqueeze(1).contiguous()
        ptr = voxel_graph.ptr
        ptr32 = ptr.to(torch.int32) if ptr.dtype != torch.int32 else ptr
        gfa = lib.segment_pool(area, ptr32.contiguous(), "sum").squeeze(1)
        first = ptr[:-1]
        got = gfa / voxel_graph.site_area[first]
        want = x[first, 9]
        return F.mse_loss(got, want) * cfg.LAMBDA_FAR


class SideLoss:
    """The terms of the generator loss that do not involve the critic (trainer.py:343-384: label cross-entropy, the two ratio
    terms, the detached FAR term) evaluated AHEAD of the critic pass together with their gradients w.r.t. (logits, label_hard).
    They depend on the generator's outputs only, so the overlapped step computes them on the sampling stream right after the
    generator's forward - beside the last critic updates - instead of as ~75 small torch launches (0.3 ms) between D(fake)'s
    forward and backward on the step's critical path.  The gradients are written out in closed form (no autograd pass):
    d ce / d logits = (softmax - onehot) * LAMBDA_LABEL / N;  the ratio terms are means of squared column-mean differences, so
    their gradient w.r.t. label_hard is one [K] vector broadcast over the rows: 2 * diff_k * LAMBDA / (cols * N)."""
    __slots__ = ("r_main", "ce", "r_void", "far", "g_logits", "g_hard_main", "g_hard_void")

    def __init__(self, voxel_graph, logits: Tensor, label_hard: Tensor, cfg):
        with torch.no_grad():
            n, k = voxel_graph.num_nodes, logits.shape[1]
            logp = F.log_softmax(logits, dim=1)
            self.ce = F.nll_loss(logp, voxel_graph.type) * cfg.LAMBDA_LABEL  # == F.cross_entropy(logits, type) * LAMBDA_LABEL
            self.g_logits = torch.sub(logp.exp(), voxel_graph.types_onehot) * (cfg.LAMBDA_LABEL / n)
            ratio_g = label_hard.squeeze(0).sum(dim=0) / n
            ratio = voxel_graph.types_onehot.sum(dim=0) / n
            self.r_main = F.mse_loss(ratio_g[:-2], ratio[:-2]) * cfg.LAMBDA_RATIO
            self.r_void = F.mse_loss(ratio_g[-2:], ratio[-2:]) * cfg.LAMBDA_RATIO_VOID
            diff = (ratio_g - ratio) / n
            w = _ratio_weights(k, cfg.LAMBDA_RATIO, cfg.LAMBDA_RATIO_VOID, diff.device)
            self.g_hard_main, self.g_hard_void = diff * w[0], diff * w[1]  # [K] each: the same row for every voxel
            self.far = far_loss(voxel_graph, label_hard, cfg)


_RATIO_W = {}


def _ratio_weights(k: int, lam_main: float, lam_void: float, device):
    """[2, K] constants of the ratio terms' gradients: 2 * LAMBDA / (number of columns of the term) on the term's columns."""
    key = (k, float(lam_main), float(lam_void), str(device))
    if key not in _RATIO_W:
        w = torch.zeros(2, k, dtype=torch.float32)
        w[0, :-2] = 2.0 * lam_main / (k - 2)
        w[1, -2:] = 2.0 * lam_void / 2
        _RATIO_W[key] = w.to(device)
    return _RATIO_W[key]


class _SideLossFn(torch.autograd.Function):
    """(logits, label_hard) -> (r_main, ce, r_void, far) with the precomputed values / gradients of ``SideLoss``."""

    @staticmethod
    def forward(ctx, logits, label_hard, side: SideLoss):
        # only the gradient tensors are kept on the node: holding `side` (which owns the OUTPUT tensors) would close a reference
        # cycle tensor -> grad_fn -> ctx -> side -> tensor through the C++ node, invisible to Python's collector - and with it the
        # generator's whole forward workspace would leak every step
        ctx.save_for_backward(side.g_logits, side.g_hard_main, side.g_hard_void)
        ctx.hard_shape = label_hard.shape
        # fresh views: the outputs get this node as their grad_fn, `side` keeps its own tensors untouched (and reusable)
        far = side.far.view_as(side.far)
        ctx.mark_non_differentiable(far)
        return side.r_main.view_as(side.r_main), side.ce.view_as(side.ce), side.r_void.view_as(side.r_void), far

    @staticmethod
    def backward(ctx, g_main, g_ce, g_void, _g_far):
        g_logits, g_hard_main, g_hard_void = ctx.saved_tensors
        row = torch.addcmul(g_hard_main * g_main, g_hard_void, g_void)  # [K]
        return g_logits * g_ce, row.expand(ctx.hard_shape), None


def _label_terms(voxel_graph, logits: Tensor, label_hard: Tensor, cfg):
    ce = F.cross_entropy(logits, voxel_graph.type) * cfg.LAMBDA_LABEL
    n = voxel_graph.num_nodes
    ratio_g = label_hard.squeeze(0).sum(dim=0) / n
    ratio = voxel_graph.types_onehot.sum(dim=0) / n
    r_main = F.mse_loss(ratio_g[:-2], ratio[:-2]) * cfg.LAMBDA_RATIO
    r_void = F.mse_loss(ratio_g[-2:], ratio[-2:]) * cfg.LAMBDA_RATIO_VOID
    return ce, r_main, r_void


def generator_loss(discriminator, local_graph, voxel_graph, logits: Tensor, label_hard: Tensor, cfg,
                   side: Optional[SideLoss] = None) -> Tensor:
    """trainer.py:334-385.  ``side``: the critic-independent terms precomputed by ``SideLoss`` (same values, same sum order;
    the gradient contributions reach (logits, label_hard) through one custom node instead of ~40 autograd nodes)."""
    d_fake = discriminator(local_graph, voxel_graph, label_hard)
    adv = -d_fake.mean() if cfg.USE_WGANGP else F.binary_cross_entropy(d_fake, torch.ones_like(d_fake))  # trainer.py:337-341
    adv = adv * cfg.LAMBDA_ADV
    if side is not None:
        r_main, ce, r_void, far = _SideLossFn.apply(logits, label_hard, side)
        return adv + r_main + ce + r_void + far
    ce = F.cross_entropy(logits, voxel_graph.type) * cfg.LAMBDA_LABEL
    n = voxel_graph.num_nodes
    ratio_g = label_hard.squeeze(0).sum(dim=0) / n
    ratio = voxel_graph.types_onehot.sum(dim=0) / n
    r_main = F.mse_loss(ratio_g[:-2], ratio[:-2]) * cfg.LAMBDA_RATIO
    r_void = F.mse_loss(ratio_g[-2:], ratio[-2:]) * cfg.LAMBDA_RATIO_VOID
    return adv + r_main + ce + r_void + far_loss(voxel_graph, label_hard, cfg)


def _prf(cm: Tensor):
    """sklearn precision / recall / f1 with average="macro", zero_division=0 from confusion counts cm[..., K, K]
    (rows = target, columns = prediction): per class over the classes present in the targets or the predictions."""
    cm = cm.to(torch.float64)
    tp = cm.diagonal(dim1=-2, dim2=-1)
    pred, true = cm.sum(-2), cm.sum(-1)
    present = (pred + true) > 0
    n = present.sum(-1).clamp_min(1)
    safe = lambda a, b: torch.where(b > 0, a / b.clamp_min(1), torch.zeros_like(a))
    prec, rec, f1 = safe(tp, pred), safe(tp, true), safe(2 * tp, pred + true)
    avg = lambda v: (v * present).sum(-1) / n
    return avg(prec), avg(rec), avg(f1)


def compute_metrics(voxel_graph, label_hard: Tensor, cfg):
    """trainer.py:387-443 without sklearn and without per-building D2H copies: ONE confusion-matrix kernel
    (bg_segment_confusion, [B,7,7] counts) and a handful of tiny tensor ops.  Returns device tensors
    (f1, f1 per building [B], precision, recall, accuracy) equal to sklearn's macro scores with zero_division=0;
    call ``.item()`` / ``.tolist()`` where the reference logs them."""
    if cfg.METRICS_AVERAGE != "macro":
        raise NotImplementedError(f"METRICS_AVERAGE={cfg.METRICS_AVERAGE!r}: only 'macro' (config.py:81) is on the B200 path")
    ptr = voxel_graph.ptr
    ptr32 = (ptr.to(torch.int32) if ptr.dtype != torch.int32 else ptr).contiguous()
    cm = lib.segment_confusion(label_hard.squeeze(0).detach().to(torch.float32).contiguous(), voxel_graph.type.contiguous(), ptr32)
    _, _, f1_each = _prf(cm)
    tot = cm.sum(0)
    prec, rec, f1 = _prf(tot)
    acc = tot.diagonal().sum().to(torch.float64) / tot.sum().clamp_min(1)
    return f1, f1_each, prec, rec, acc


def train_step(generator, discriminator, opt_g, opt_d, local_graph, voxel_graph, cfg, rng: str = "cpu",
               grad_sync=None, sync_losses="each", overlap=False):
    """trainer.py:467-495 for one device-resident batch (see ``_train_step``); the multi-stream variant runs with
    programmatic dependent launch scoped off."""
    with (lanes_pdl_scope() if overlap is True else contextlib.nullcontext()):
        return _train_step(generator, discriminator, opt_g, opt_d, local_graph, voxel_graph, cfg, rng, grad_sync, sync_losses,
                           overlap)


def _train_step(generator, discriminator, opt_g, opt_d, local_graph, voxel_graph, cfg, rng, grad_sync, sync_losses, overlap):
    """trainer.py:467-495 for one device-resident batch.  ``grad_sync(model)`` (optional) is called after each
    backward, before the optimiser step - the data-parallel gradient all-reduce hooks in here.
    Returns (critic losses, generator loss, label_hard[1,N,7]).  ``sync_losses``: "each" = ``.item()`` right after every
    backward exactly like the reference (6 host syncs per step, trainer.py:479,493); "step" = the same 6 floats read
    back with ONE device->host copy at the end of the step (the host keeps running ahead of the GPU inside the step);
    False = 0-dim device tensors, no sync.
    ``overlap``: False = every pass on the caller's stream in the reference's call order.  True = the independent passes of
    the step run concurrently (``Lanes``): the N_CRITIC sampling passes of the generator are issued up front on their own
    stream (z and the GP mixing factors are still drawn in the reference's order z1, e1, z2, e2, ...), and D(real) / D(fake)
    of each critic update run beside the gradient-penalty pass.  "order" = the call order of True on a single stream (what
    the tests compare True against)."""
    dev = voxel_graph.x.device
    n = voxel_graph.num_nodes
    each = sync_losses in (True, "each")
    d_losses = []
    lanes = Lanes.get(dev) if overlap is True else None
    samples, es = None, None
    if overlap:
        from . import models
        models._batch_ctx(local_graph, voxel_graph, cfg.NUM_CLASSES)  # per-batch constants: built once, on this stream
        main = torch.cuda.current_stream()
        zs, es = [], []
        for _ in range(cfg.N_CRITIC):
            zs.append(_rand((1, n, cfg.Z_DIM), dev, rng, normal=True))
            es.append(_rand((n, 1), dev, rng, normal=False))
        samples = []
        start = main.record_event()
        gen_stream = lanes.gen if lanes is not None else main
        with torch.no_grad(), torch.cuda.stream(gen_stream):
            if lanes is not None:
                gen_stream.wait_event(start)
            for z in zs:
                _, hard, soft = generator(local_graph, voxel_graph, z)
                if lanes is not None:  # produced on the sampling stream, consumed on the other three
                    for t in (hard, soft):
                        for s_ in (main, lanes.real, lanes.fake):
                            t.record_stream(s_)
                samples.append((hard.unsqueeze(0), soft.unsqueeze(0), gen_stream.record_event() if lanes is not None else None))
    for it in range(cfg.N_CRITIC):
        if samples is None:
            with torch.no_grad():
                z = _rand((1, n, cfg.Z_DIM), dev, rng, normal=True)
                _, hard, soft = generator(local_graph, voxel_graph, z)
                hard, soft = hard.unsqueeze(0), soft.unsqueeze(0)
            e = None
        else:
            hard, soft, ready = samples[it]
            e = es[it]
            if ready is not None:
                torch.cuda.current_stream().wait_event(ready)
        opt_d.zero_grad()
        d_loss = discriminator_loss(discriminator, local_graph, voxel_graph, hard, soft, cfg, rng, lanes, e)
        d_loss.backward()
        if lanes is not None:
            main = torch.cuda.current_stream()
            main.wait_stream(lanes.real)
            main.wait_stream(lanes.fake)
            discriminator.merge_lanes()
        if grad_sync is not None:
            grad_sync(discriminator)
        d_losses.append(d_loss.item() if each else d_loss.detach())
        opt_d.step()
    z = _rand((1, n, cfg.Z_DIM), dev, rng, normal=True)
    logits, hard, soft = generator(local_graph, voxel_graph, z)
    hard = hard.unsqueeze(0)
    opt_g.zero_grad()
    g_loss = generator_loss(discriminator, local_graph, voxel_graph, logits, hard, cfg)
    g_loss.backward()
    if grad_sync is not None:
        grad_sync(generator)
    g_val = g_loss.item() if each else g_loss.detach()
    opt_g.step()
    if sync_losses == "step":
        vals = torch.stack(d_losses + [g_val]).tolist()  # one D2H copy + one sync for the whole step
        d_losses, g_val = vals[:-1], vals[-1]
    return d_losses, g_val, hard.detach()


@torch.no_grad()
def sample(generator, local_graph, voxel_graph, cfg, rng: str = "device") -> Tensor:
    """Generator-only sampling pass (trainer.py:769-770 + :73): voxel -> program labels [N] (int64)."""
    z = _rand((1, voxel_graph.num_nodes, cfg.Z_DIM), voxel_graph.x.device, rng, normal=True)
    _, hard, _ = generator(local_graph, voxel_graph, z)
    return hard.argmax(dim=1)
