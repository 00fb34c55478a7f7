"""step.SideLoss / step._SideLossFn (host logic, no GPU): the closed-form gradients of the critic-independent generator-loss
terms (label cross-entropy, the two ratio terms; reference trainer.py:343-356) against torch autograd of the reference spelling,
and the custom node that feeds them into the generator's backward.  The FAR term goes through a device kernel (bg_segment_pool)
and is replaced by a constant here; its value is checked on the GPU (tests/test_trainer_ops_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

from building_gan_b200 import step


class _Voxels:
    def __init__(self, n, k, seed):
        g = torch.Generator().manual_seed(seed)
        self.num_nodes = n
        self.type = torch.randint(0, k, (n,), generator=g)
        self.types_onehot = F.one_hot(self.type, k)  # int64, as the reference holds it


class _Cfg:
    LAMBDA_LABEL, LAMBDA_RATIO, LAMBDA_RATIO_VOID = 3.0, 5.0, 7.0


@pytest.mark.parametrize("n,k", [(50, 7), (1, 7), (333, 5)])
def test_side_loss_closed_form_gradients_match_autograd(monkeypatch, n, k):
    monkeypatch.setattr(step, "far_loss", lambda *a: torch.tensor(0.25))
    vg = _Voxels(n, k, seed=n)
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(n, k, generator=g, requires_grad=True)
    hard = torch.rand(1, n, k, generator=g, requires_grad=True)
    side = step.SideLoss(vg, logits.detach(), hard.detach(), _Cfg)
    ce, r_main, r_void = step._label_terms(vg, logits, hard, _Cfg)
    assert torch.equal(side.ce, ce.detach()) and torch.equal(side.r_main, r_main.detach()) and torch.equal(side.r_void, r_void.detach())
    (g_logits,) = torch.autograd.grad(ce, logits)
    (g_main,) = torch.autograd.grad(r_main, hard, retain_graph=True)
    (g_void,) = torch.autograd.grad(r_void, hard)
    scale = lambda t: max(float(t.abs().max()), 1e-12)
    assert float((side.g_logits - g_logits).abs().max()) <= 1e-6 * scale(g_logits)
    assert float((side.g_hard_main.expand_as(g_main) - g_main).abs().max()) <= 1e-6 * scale(g_main)
    assert float((side.g_hard_void.expand_as(g_void) - g_void).abs().max()) <= 1e-6 * scale(g_void)
    # the node: arbitrary weights on the three differentiable terms, nothing through the FAR term
    a, b, c, d = step._SideLossFn.apply(logits, hard, side)
    assert not d.requires_grad and float(d) == 0.25
    (2.0 * a + 3.0 * b + 0.5 * c + d).backward()
    assert float((logits.grad - 3.0 * g_logits).abs().max()) <= 1e-5 * scale(g_logits)
    assert float((hard.grad - (2.0 * g_main + 0.5 * g_void)).abs().max()) <= 1e-5 * scale(g_main)
    # the SideLoss object stays usable (its tensors were not turned into graph outputs)
    assert side.r_main.grad_fn is None and side.ce.grad_fn is None
    again = step._SideLossFn.apply(logits.detach().requires_grad_(True), hard.detach().requires_grad_(True), side)
    assert torch.equal(again[0], side.r_main)


def test_side_loss_node_holds_no_reference_to_its_outputs():
    """The first version kept the SideLoss object (which owns the output tensors) on the autograd node: a reference cycle through the
    C++ node that Python's collector cannot see - the whole generator graph behind `logits` leaked every step."""
    import gc
    import weakref

    vg = _Voxels(20, 7, seed=1)
    old = step.far_loss
    step.far_loss = lambda *a: torch.tensor(0.0)
    try:
        base, saved = torch.randn(20, 7, requires_grad=True), torch.randn(20, 7)
        logits = base * saved  # stands for the generator graph: its node keeps `saved` alive (as the real one keeps the workspace)
        hard = torch.rand(1, 20, 7, requires_grad=True)
        side = step.SideLoss(vg, logits.detach(), hard.detach(), _Cfg)
        outs = step._SideLossFn.apply(logits, hard, side)
        probe = weakref.ref(saved)
        del logits, outs, side, saved
        gc.collect()
        assert probe() is None, "the autograd graph behind `logits` is still alive after every user-visible reference is gone"
    finally:
        step.far_loss = old
