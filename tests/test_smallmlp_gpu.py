"""bg_small_mlp_fwd / _bwd (csrc/bg_smallmlp.cu): a chain of Linear (+LayerNorm) (+activation) layers on <= 8 rows as one launch
per direction - the generator's matched_features_encoder on the 7 type-table rows (reference models.py:36-47,131-133) - against
fp64 torch autograd of the same chain, and against the per-layer kernels it replaces inside the generator pass."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from util import assert_close

from building_gan_b200 import lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _chain(widths, ln, act, seed):
    g = torch.Generator().manual_seed(seed)
    layers = []
    for cin, cout in zip(widths[:-1], widths[1:]):
        l = {"W": torch.randn(cout, cin, generator=g, dtype=torch.float64) * (1.0 / cin ** 0.5),
             "bias": torch.randn(cout, generator=g, dtype=torch.float64) * 0.3, "act": act}
        if ln:
            l["ln"] = (1 + 0.2 * torch.randn(cout, generator=g, dtype=torch.float64), 0.2 * torch.randn(cout, generator=g, dtype=torch.float64))
        layers.append(l)
    return layers


def _ref_forward(x, layers):
    outs = []
    for l in layers:
        y = x @ l["W"].t() + l["bias"]
        if l.get("ln") is not None:
            y = F.layer_norm(y, (y.shape[1],), l["ln"][0], l["ln"][1], 1e-5)
        x = F.leaky_relu(y, 0.2) if l["act"] == 2 else torch.relu(y) if l["act"] == 1 else y
        outs.append(x)
    return outs


def _to_dev(layers):
    f = lambda t: t.float().to(DEV).contiguous()
    return [{"W": f(l["W"]), "bias": f(l["bias"]), "act": l["act"], **({"ln": (f(l["ln"][0]), f(l["ln"][1]))} if l.get("ln") is not None else {})}
            for l in layers]


@pytest.mark.parametrize("rows", [1, 7, 8])
@pytest.mark.parametrize("widths,ln,act", [((17, 128, 128, 128, 128, 128), True, 2), ((5, 33, 64, 1), False, 1), ((128, 96, 128), True, 0),
                                           ((12, 128), True, 2)])
def test_small_mlp_forward_backward_match_fp64_autograd(rows, widths, ln, act):
    layers = _chain(widths, ln, act, seed=len(widths) * 10 + rows)
    g = torch.Generator().manual_seed(99 + rows)
    x = torch.randn(rows, widths[0], generator=g, dtype=torch.float64)
    gout = torch.randn(rows, widths[-1], generator=g, dtype=torch.float64)
    leaves = []
    for l in layers:
        for k in ("W", "bias"):
            l[k].requires_grad_(True)
            leaves.append(l[k])
        if l.get("ln") is not None:
            l["ln"][0].requires_grad_(True), l["ln"][1].requires_grad_(True)
            leaves += list(l["ln"])
    xr = x.clone().requires_grad_(True)
    outs = _ref_forward(xr, layers)
    grads = torch.autograd.grad(outs[-1], [xr] + leaves, gout)
    dl = _to_dev([{k: (v.detach() if torch.is_tensor(v) else tuple(t.detach() for t in v) if isinstance(v, tuple) else v) for k, v in l.items()}
                  for l in layers])
    xd = x.float().to(DEV).contiguous()
    saved = lib.small_mlp_fwd(xd, dl)
    for i, (sv, o) in enumerate(zip(saved, outs)):
        assert_close(sv["out"], o.detach(), 1e-5, f"layer {i} output")
    gin, gp = lib.small_mlp_bwd(xd, dl, saved, gout.float().to(DEV).contiguous())
    assert_close(gin, grads[0], 2e-5, "input gradient")
    it = iter(grads[1:])
    for i, (l, gd) in enumerate(zip(layers, gp)):
        assert_close(gd["dW"], next(it), 2e-5, f"layer {i} dW")
        assert_close(gd["dbias"], next(it), 2e-5, f"layer {i} dbias")
        if l.get("ln") is not None:
            assert_close(gd["dgamma"], next(it), 2e-5, f"layer {i} dgamma")
            assert_close(gd["dbeta"], next(it), 2e-5, f"layer {i} dbeta")
    # accumulate mode adds into existing buffers; two runs are bit-identical
    base = [{k: torch.full_like(v, 0.5) for k, v in gd.items()} for gd in gp]
    gin2, gp2 = lib.small_mlp_bwd(xd, dl, saved, gout.float().to(DEV).contiguous(), grads=base)
    assert torch.equal(gin2, gin)
    for a, b in zip(gp2, gp):
        for k in b:
            assert_close(a[k] - 0.5, b[k], 1e-6, f"accumulate {k}")


def test_small_mlp_rejects_unsupported_shapes():
    x = torch.zeros(9, 4, device=DEV)
    with pytest.raises(RuntimeError, match="rows"):
        lib.small_mlp_fwd(x, [{"W": torch.zeros(4, 4, device=DEV)}])
    with pytest.raises(RuntimeError, match="widths"):
        lib.small_mlp_fwd(torch.zeros(2, 200, device=DEV), [{"W": torch.zeros(4, 200, device=DEV)}])


_AB = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
from building_gan_b200 import Configuration, graph, synth
from building_gan_b200.models import VoxelGNNGenerator
cfg = Configuration()
lb, vb = graph.collate_fn([synth.building_pair(i) for i in (4001, 4002)])
lb, vb = lb.to("cuda"), vb.to("cuda")
torch.manual_seed(5)
G = VoxelGNNGenerator(cfg, 17, 12).to("cuda").eval()
z = torch.randn(1, vb.num_nodes, cfg.Z_DIM, device="cuda")
noise = -torch.empty(vb.num_nodes, 7, device="cuda").exponential_().log()
logits, hard, soft = G(lb, vb, z, noise)
(logits.square().sum() + (hard * torch.arange(7, device="cuda")).sum()).backward()
torch.save({"logits": logits.detach().cpu(), "grads": [p.grad.cpu() for p in G.parameters()]}, sys.argv[2])
"""


def test_generator_pass_with_the_fused_encoder_chain_matches_the_per_layer_launches(tmp_path):
    """The whole generator forward + backward with the fused matched_features_encoder (default) against BG_SMALL_MLP=0 (the
    per-layer kernels), each in its own process (the switch is read once per process)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for flag in ("1", "0"):
        out = tmp_path / f"g{flag}.pt"
        env = dict(os.environ, BG_SMALL_MLP=flag)
        subprocess.run([sys.executable, "-c", _AB, root, str(out)], check=True, env=env, timeout=600)
        res.append(torch.load(out))
    assert_close(res[0]["logits"], res[1]["logits"], 1e-4, "logits")
    # the encoder's own 20 parameter tensors come out of the fused backward; the rest of the generator sees the encoder only through
    # its (1e-7-different) forward values, amplified by the 1- / 2-channel bottleneck blocks like any fp32 reordering
    # (tests/test_models_gpu.py states 2e-3 for those against fp64)
    # Error norm: max |a - b| over max(|b|_max, 1e-2 x the largest gradient of the model) - some tensors (every att_dst: the
    # edge softmax is invariant to a per-destination shift) have an exactly-zero true gradient and hold rounding noise only.
    gmax = max(float(b.abs().max()) for b in res[1]["grads"])
    for i, (a, b) in enumerate(zip(res[0]["grads"], res[1]["grads"])):
        err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-2 * gmax)
        assert err <= (3e-3 if i < 20 else 1e-2), f"parameter gradient {i}: {err:.3e}"
