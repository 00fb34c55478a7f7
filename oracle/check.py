"""ORACLE-side parity checker (test infrastructure, see oracle/__init__.py): runs the drop-in models of the product and
the oracle models (fp64, CPU) on the SAME buildings, weights and random draws and reports how far apart they are.

Used as the checker by ``tests/`` (benchmark-size parity), ``__graft_entry__.smoke()`` and the ``parity`` block of
``bench.py``'s line - never as the thing measured.  Error norm everywhere: max |a - ref| / max |ref| per tensor (the
"rel 1e-5 of the tensor's magnitude" of BASELINE.json's north star); ``*_elementwise`` adds the stricter per-element
relative error over the elements that are not tiny (|ref| > 1e-3 max|ref|).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from . import models as omodels
from . import pyg as opyg
from . import trainer as otrainer


def rel(a: Tensor, ref: Tensor) -> float:
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def rel_elementwise(a: Tensor, ref: Tensor, floor: float = 1e-3) -> float:
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    big = ref.abs() > floor * ref.abs().max()
    if not bool(big.any()):
        return 0.0
    return float(((a - ref).abs()[big] / ref.abs()[big]).max())


def to_oracle_batch(batch, dtype=torch.float64) -> "opyg.Batch":
    """Product ``Batch`` (any device) -> oracle ``Batch`` on the CPU, floating-point fields in ``dtype``."""
    graphs = []
    for i in range(batch.num_graphs):
        g = batch[i]
        fields = {}
        for k, v in g._fields.items():
            if isinstance(v, Tensor):
                v = v.detach().cpu()
                if v.is_floating_point():
                    v = v.to(dtype)
            fields[k] = v
        graphs.append(opyg.Data(**fields))
    return opyg.Batch.from_data_list(graphs)


def oracle_twins(G, D, cfg, dtype=torch.float64):
    oG, oD = omodels.OracleGenerator(cfg, G.local_graph_dim, G.voxel_graph_dim), omodels.OracleDiscriminator(cfg, D.local_graph_dim, D.voxel_graph_dim)
    oG.load_state_dict({k: v.detach().cpu() for k, v in G.state_dict().items()})
    oD.load_state_dict({k: v.detach().cpu() for k, v in D.state_dict().items()})
    return oG.to(dtype), oD.to(dtype)


def _worst_grad(model, omodel, omodel32=None) -> Dict[str, object]:
    """Parameter-gradient error of a model against the fp64 oracle.
    ``grad_rel_l2``  || g - g_ref ||_2 / || g_ref ||_2 over ALL parameters concatenated (the error of the update direction).
    ``worst_grad_rel``  worst tensor: max|err| / max(max|ref grad|, 1e-2 gmax), gmax = the model's largest reference gradient:
        the absolute error of any fp32 evaluation scales with the magnitudes upstream, so tensors whose own gradient is two
        orders below gmax (exact-cancellation gradients: a conv bias in front of a GraphNorm, weight columns that only ever
        multiply zero inputs) are judged on that absolute scale.
    With ``omodel32`` (the same oracle in fp32 = the reference's own arithmetic) also the fp32 oracle's own figures."""
    gmax = max(float(p.grad.abs().max()) for p in omodel.parameters() if p.grad is not None)
    o32 = dict(omodel32.named_parameters()) if omodel32 is not None else {}
    worst, name, w32 = 0.0, "", 0.0
    num = den = num32 = 0.0
    for (k, p), (_, op) in zip(model.named_parameters(), omodel.named_parameters()):
        if op.grad is None:
            continue
        if p.grad is None:
            return {"worst_grad_rel": float("inf"), "worst_grad_param": k, "grad_rel_l2": float("inf")}
        diff = p.grad.detach().double().cpu() - op.grad
        num += float(diff.pow(2).sum())
        den += float(op.grad.pow(2).sum())
        scale = max(float(op.grad.abs().max()), 1e-2 * gmax)
        e = float(diff.abs().max()) / scale
        if e > worst:
            worst, name = e, k
        if k in o32 and o32[k].grad is not None:
            d32 = o32[k].grad.double() - op.grad
            num32 += float(d32.pow(2).sum())
            w32 = max(w32, float(d32.abs().max()) / scale)
    out = {"worst_grad_rel": worst, "worst_grad_param": name, "grad_rel_l2": (num / max(den, 1e-300)) ** 0.5}
    if o32:
        out["fp32_oracle_worst_grad_rel"], out["fp32_oracle_grad_rel_l2"] = w32, (num32 / max(den, 1e-300)) ** 0.5
    return out


def parity_report(G, D, local_graph, voxel_graph, cfg, step_module, seed: int = 2024, gradients: bool = True,
                  envelope: bool = True) -> Dict[str, object]:
    """Eval-mode (no dropout) generator forward, critic loss with its gradient penalty (second-order backward) and generator
    loss, product vs fp64 oracle with shared z / Gumbel noise / mixing factor.  ``step_module`` = the product's ``step``
    (passed in: oracle/ never imports the product).  ``envelope``: also runs the oracle in fp32 - the reference's own
    arithmetic - and reports ITS distance from fp64 (the rounding envelope of a correct fp32 implementation)."""
    dev = voxel_graph.x.device
    was = G.training, D.training
    G.eval(), D.eval()
    oG, oD = oracle_twins(G, D, cfg)
    oG.eval(), oD.eval()
    olb, ovb = to_oracle_batch(local_graph), to_oracle_batch(voxel_graph)
    n, k = voxel_graph.num_nodes, cfg.NUM_CLASSES
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(1, n, cfg.Z_DIM, generator=gen)
    noise = -torch.empty(n, k).exponential_(generator=gen).log()
    out: Dict[str, object] = {"N": int(n), "graphs": int(voxel_graph.num_graphs), "oracle": "fp64 CPU restatement (oracle/models.py)"}
    with torch.no_grad():
        ol, oh, os_ = oG(olb, ovb, z.double(), noise.double())
        kl, kh, ks = G(local_graph, voxel_graph, z.to(dev), noise.to(dev))
    out["logits_rel"], out["logits_rel_elementwise"] = rel(kl, ol), rel_elementwise(kl, ol)
    out["label_soft_rel"] = rel(ks, os_)
    top2 = os_.topk(2, dim=1).values
    tie = (top2[:, 0] - top2[:, 1]) <= 1e-4
    differ = kh.argmax(1).cpu() != oh.argmax(1)
    out["labels_differ_outside_ties"] = int((differ & ~tie).sum())
    out["labels_differ_at_ties"], out["near_ties"] = int((differ & tie).sum()), int(tie.sum())
    oG32 = oD32 = None
    if envelope:
        oG32, oD32 = oracle_twins(G, D, cfg, torch.float32)
        oG32.eval(), oD32.eval()
        olb32, ovb32 = to_oracle_batch(local_graph, torch.float32), to_oracle_batch(voxel_graph, torch.float32)
        with torch.no_grad():
            l32, h32, s32 = oG32(olb32, ovb32, z, noise)
        out["fp32_oracle_logits_rel"] = rel(l32, ol)
    # critic loss: one float32 draw of the mixing factor from the CPU generator (trainer.py:298), handed to both sides
    hard_in, soft_in = oh.detach().unsqueeze(0), os_.detach().unsqueeze(0)
    D.zero_grad(set_to_none=True), oD.zero_grad(set_to_none=True)
    e = torch.rand(n, 1, generator=gen)
    o_d = otrainer.discriminator_loss(oD, olb, ovb, hard_in, soft_in, cfg, e=e.double())
    k_d = step_module.discriminator_loss(D, local_graph, voxel_graph, hard_in.float().to(dev), soft_in.float().to(dev), cfg,
                                         rng="cpu", e=e.to(dev))
    out["d_loss_rel"] = rel(k_d.reshape(1), o_d.reshape(1))
    if gradients:
        o_d.backward(), k_d.backward()
        if envelope:  # the same loss in the reference's own fp32 arithmetic
            otrainer.discriminator_loss(oD32, olb32, ovb32, hard_in.float(), soft_in.float(), cfg, e=e).backward()
        g = _worst_grad(D, oD, oD32)
        out["d_worst_grad_rel"], out["d_worst_grad_param"], out["d_grad_rel_l2"] = g["worst_grad_rel"], g["worst_grad_param"], g["grad_rel_l2"]
        if envelope:
            out["fp32_oracle_d_worst_grad_rel"], out["fp32_oracle_d_grad_rel_l2"] = g["fp32_oracle_worst_grad_rel"], g["fp32_oracle_grad_rel_l2"]
    # generator loss through the (updated-gradient-free) critic
    G.zero_grad(set_to_none=True), oG.zero_grad(set_to_none=True)
    ol2, oh2, _ = oG(olb, ovb, z.double(), noise.double())
    kl2, kh2, _ = G(local_graph, voxel_graph, z.to(dev), noise.to(dev))
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)  # int64 one-hot counts / N must come out in the oracle's precision
    try:
        o_g = otrainer.generator_loss(oD, olb, ovb, ol2, oh2.unsqueeze(0), cfg)
    finally:
        torch.set_default_dtype(prev)
    k_g = step_module.generator_loss(D, local_graph, voxel_graph, kl2, kh2.unsqueeze(0), cfg)
    out["g_loss_rel"] = rel(k_g.reshape(1), o_g.reshape(1))
    if gradients:
        o_g.backward(), k_g.backward()
        if envelope:
            oD32.zero_grad(set_to_none=True)
            l32b, h32b, _ = oG32(olb32, ovb32, z, noise)
            otrainer.generator_loss(oD32, olb32, ovb32, l32b, h32b.unsqueeze(0), cfg).backward()
        g = _worst_grad(G, oG, oG32)
        out["g_worst_grad_rel"], out["g_worst_grad_param"], out["g_grad_rel_l2"] = g["worst_grad_rel"], g["worst_grad_param"], g["grad_rel_l2"]
        if envelope:
            out["fp32_oracle_g_worst_grad_rel"], out["fp32_oracle_g_grad_rel_l2"] = g["fp32_oracle_worst_grad_rel"], g["fp32_oracle_grad_rel_l2"]
        out["worst_grad_rel"] = max(out["d_worst_grad_rel"], out["g_worst_grad_rel"])
    G.zero_grad(set_to_none=True), D.zero_grad(set_to_none=True)
    G.train(was[0]), D.train(was[1])
    out["grad_norm"] = ("*_grad_rel_l2: ||g - ref||_2 / ||ref||_2 over all parameters; *_worst_grad_rel: worst tensor, max|err| / "
                        "max(max|ref grad|, 1e-2 x the model's largest gradient); fp32_oracle_*: the reference's own fp32 arithmetic "
                        "measured the same way")
    return {k_: (round(v, 10) if isinstance(v, float) else v) for k_, v in out.items()}
