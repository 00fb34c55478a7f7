"""SURVEY section 8(f) N1 + the USE_WGANGP=False branch: the trainer-side pieces around the models.

* ``step.compute_metrics`` (one confusion-matrix kernel) against sklearn called exactly like trainer.py:387-443;
* the vanilla-GAN critic (sigmoid tail, BCE losses; trainer.py:326-330,340-341; models.py:222-223) against the oracle."""
import numpy as np
import pytest
import torch
from sklearn import metrics as skm

from oracle import trainer as otrainer
from test_models_gpu import _grads_close, _setup
from util import assert_close

from building_gan_b200 import lib, step

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_segment_confusion_counts_are_exact():
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 300, 2, 1025]
    ptr = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    n, k = int(ptr[-1]), 7
    score = torch.randn(n, k, generator=g)
    score[5] = 0.0  # a full tie: argmax takes the first index
    target = torch.randint(0, k, (n,), generator=g)
    cm = lib.segment_confusion(score.to(DEV), target.to(DEV), ptr.to(DEV)).cpu()
    pred = score.argmax(1)
    for s in range(len(sizes)):
        ref = torch.zeros(k, k, dtype=torch.int32)
        for t, p in zip(target[ptr[s]:ptr[s + 1]].tolist(), pred[ptr[s]:ptr[s + 1]].tolist()):
            ref[t, p] += 1
        assert torch.equal(cm[s], ref)


@pytest.mark.parametrize("skew", [False, True])
def test_metrics_match_sklearn_like_the_reference_calls_it(skew):
    cfg, G, D, oG, oD, lb, vb, olb, ovb = _setup(ids=(31, 32, 33))
    n = vb.num_nodes
    g = torch.Generator().manual_seed(9)
    hard = torch.zeros(n, 7)
    idx = torch.randint(0, 3 if skew else 7, (n,), generator=g)  # skew: classes missing from the predictions
    hard[torch.arange(n), idx] = 1.0
    f1, f1_each, prec, rec, acc = step.compute_metrics(vb, hard.unsqueeze(0).to(DEV), cfg)
    y, yp = ovb.type.numpy(), idx.numpy()
    assert abs(float(f1) - skm.f1_score(y, yp, average="macro", zero_division=0)) < 1e-12
    assert abs(float(prec) - skm.precision_score(y, yp, average="macro", zero_division=0)) < 1e-12
    assert abs(float(rec) - skm.recall_score(y, yp, average="macro", zero_division=0)) < 1e-12
    assert abs(float(acc) - skm.accuracy_score(y, yp)) < 1e-12
    ptr = ovb.ptr.tolist()
    for gi in range(ovb.num_graphs):
        want = skm.f1_score(y[ptr[gi]:ptr[gi + 1]], yp[ptr[gi]:ptr[gi + 1]], average="macro", zero_division=0)
        assert abs(float(f1_each[gi]) - want) < 1e-12


def test_vanilla_gan_branch_use_wgangp_false():
    """USE_WGANGP=False: discriminator ends in a sigmoid, critic / generator losses are BCE."""
    from building_gan_b200 import Configuration
    from building_gan_b200.models import VoxelGNNDiscriminator
    from oracle import models as omodels
    cfg, G, D, oG, oD, lb, vb, olb, ovb = _setup()
    cfg.USE_WGANGP = False
    D2 = VoxelGNNDiscriminator(cfg, 17, 12)
    assert isinstance(D2.decoder[-1], torch.nn.Sigmoid)
    D2.load_state_dict(D.state_dict())
    oD2 = omodels.OracleDiscriminator(cfg, 17, 12)
    oD2.load_state_dict({k: v.cpu() for k, v in D.state_dict().items()})
    oD2 = oD2.double().eval()
    D2 = D2.to(DEV).eval()
    G.eval(), oG.eval()
    n = vb.num_nodes
    z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5))
    noise = -torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()
    with torch.no_grad():
        _, hard, soft = G(lb, vb, z.to(DEV), noise.to(DEV))
    ohard, osoft = hard.cpu().double(), soft.cpu().double()
    kd = step.discriminator_loss(D2, lb, vb, hard.unsqueeze(0), soft.unsqueeze(0), cfg)
    od = otrainer.discriminator_loss(oD2, olb, ovb, ohard.unsqueeze(0), osoft.unsqueeze(0), cfg)
    assert_close(kd.reshape(1), od.reshape(1), 1e-5, "BCE critic loss")
    kd.backward(), od.backward()
    # two forwards (real, fake) share the weights, so the oracle's ReLU patterns cannot be pinned to the kernel path's here:
    # sites within rounding distance of 0 may flip (see test_models_gpu._PatternAct), hence 1e-3 instead of 1e-4
    _grads_close(D2, oD2, 1e-3, "BCE critic gradients")
    # generator side: BCE(D(fake), 1) with the straight-through labels
    D2.zero_grad(), oD2.zero_grad()
    logits, h2, s2 = G(lb, vb, z.to(DEV), noise.to(DEV))
    ol, oh, os_ = oG(olb, ovb, z.double(), noise.double())
    kg = step.generator_loss(D2, lb, vb, logits, h2.unsqueeze(0), cfg)
    og = otrainer.generator_loss(oD2, olb, ovb, ol, oh.unsqueeze(0), cfg)
    assert_close(kg.reshape(1), og.reshape(1), 2e-5, "BCE generator loss")


@pytest.mark.gpu
def test_fused_critic_loss_matches_torch():
    """bg_gp_mix / bg_critic_loss_fwd / _bwd against the reference's torch spelling (trainer.py:298-301, 314, 323)."""
    from building_gan_b200 import lib, step as bstep

    torch.manual_seed(0)
    n, k, lam = 5003, 7, 10.0
    dev = "cuda"
    e = torch.rand(n, 1, device=dev)
    onehot = torch.nn.functional.one_hot(torch.randint(0, k, (n,), device=dev), k)  # int64, as the reference holds it
    soft = torch.softmax(torch.randn(n, k, device=dev), dim=1)
    mixed = lib.gp_mix(e, onehot, soft)
    assert torch.equal(mixed, e * onehot + (1 - e) * soft)
    d_fake = torch.randn(n, 1, device=dev, requires_grad=True)
    d_real = torch.randn(n, 1, device=dev, requires_grad=True)
    grad = (torch.randn(n, k, device=dev) * 0.3)
    grad[7] = 0.0  # a zero row: torch's norm backward gives a zero subgradient there
    grad.requires_grad_(True)
    ref = d_fake.double().mean() - d_real.double().mean() + ((grad.double().norm(dim=1) - 1) ** 2).mean() * lam
    gref = torch.autograd.grad(ref, (d_fake, d_real, grad))
    got = bstep._CriticLossFn.apply(d_fake, d_real, grad, lam)
    ggot = torch.autograd.grad(got, (d_fake, d_real, grad))
    assert abs(got.item() - ref.item()) <= 1e-6 * abs(ref.item())
    for a, b in zip(ggot, gref):
        assert_close(a, b, 1e-6)
    # deterministic
    got2 = bstep._CriticLossFn.apply(d_fake, d_real, grad, lam)
    assert got2.item() == got.item()


def test_precomputed_side_loss_matches_the_spelled_generator_loss():
    """step.SideLoss (critic-independent generator-loss terms + their gradients, evaluated ahead of the critic pass) gives the
    same loss value bit for bit and the same parameter gradients as the reference spelling (trainer.py:334-385)."""
    cfg, G, D, oG, oD, lb, vb, olb, ovb = _setup()
    G.eval(), D.eval()
    n = vb.num_nodes
    z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5)).to(DEV)
    noise = (-torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()).to(DEV)
    res = []
    for use_side in (False, True):
        G.zero_grad(), D.zero_grad()
        logits, hard, _ = G(lb, vb, z, noise)
        side = step.SideLoss(vb, logits, hard.unsqueeze(0), cfg) if use_side else None
        loss = step.generator_loss(D, lb, vb, logits, hard.unsqueeze(0), cfg, side=side)
        loss.backward()
        res.append((loss.detach().clone(), [p.grad.detach().clone() for p in G.parameters()]))
    assert torch.equal(res[0][0], res[1][0])
    # error norm: max |a - b| / max(|a|_max, 1e-2 x the model's largest gradient) - tensors whose true gradient is exactly zero
    # (every att_dst: the edge softmax is invariant to a per-destination shift) hold rounding noise only
    gmax = max(float(a.abs().max()) for a in res[0][1])
    for i, (a, b) in enumerate(zip(res[0][1], res[1][1])):
        err = float((a - b).abs().max()) / max(float(a.abs().max()), 1e-2 * gmax)
        assert err <= 1e-4, f"generator gradient {i} with the precomputed side terms: {err:.3e}"


def test_gradient_penalty_backward_runs_no_zero_gradient_sweep():
    """d_loss.backward() of the WGAN-GP critic loss must not visit the gradient-penalty pass's forward node: nothing flows into
    its score (only into its input gradient), and a materialised all-zero g_score would cost one full first-order backward per
    critic update on the step's critical path.  Counted by the library's launch counter; gradients checked against a run whose
    forward node IS visited (a zero-weighted score term added to the loss)."""
    cfg, G, D, oG, oD, lb, vb, olb, ovb = _setup()
    D.eval()
    n = vb.num_nodes
    g = torch.Generator().manual_seed(11)
    soft = torch.softmax(torch.randn(n, 7, generator=g), dim=1).to(DEV)
    e = torch.rand(n, 1, generator=g).to(DEV)

    def run(visit_forward_node: bool):
        D.zero_grad()
        mixed = lib.gp_mix(e, vb.types_onehot.contiguous(), soft).requires_grad_(True)
        score = D(lb, vb, mixed.unsqueeze(0))
        (grad,) = torch.autograd.grad(score, mixed, torch.ones_like(score), create_graph=True, only_inputs=True)
        loss = ((grad.norm(dim=1) - 1) ** 2).mean() * cfg.LAMBDA_GP
        if visit_forward_node:
            loss = loss + 0.0 * score.sum()
        l0 = lib.LAUNCHES
        loss.backward()
        return lib.LAUNCHES - l0, [p.grad.detach().clone() for p in D.parameters()]

    n_skip, g_skip = run(False)
    n_visit, g_visit = run(True)
    assert n_skip < n_visit, (n_skip, n_visit)
    for a, b in zip(g_skip, g_visit):
        assert_close(a, b, 1e-6, "gradient-penalty parameter gradients")
