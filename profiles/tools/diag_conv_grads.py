"""Diagnostic: per-parameter gradient error of the generator built with a conv type, against the fp64 oracle, next to the
fp32 oracle's own error (the rounding envelope).  python profiles/tools/diag_conv_grads.py [KIND] [SEED]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_models_gpu as T
from building_gan_b200 import lib

DEV = "cuda"


def run(kind, seed, tc, sync=True):
    lib.set_dense_tc(tc)
    cfg, G, D, oG, oD, lb, vb, olb, ovb = T._setup(conv=kind, seed=seed)
    n = vb.num_nodes
    z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5))
    noise = -torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()
    G.eval(), oG.eval()
    ologits, ohard, osoft = oG(olb, ovb, z.double(), noise.double())
    G.debug_keep_saved = True
    logits, hard, soft = G(lb, vb, z.to(DEV), noise.to(DEV))
    w1, w2, w3 = (torch.randn(n, 7, generator=torch.Generator().manual_seed(s), dtype=torch.float64) for s in (8, 9, 10))
    if sync:
        T._sync_patterns(oG, G.debug_saved, ovb.type)
    plogits, phard, psoft = oG(olb, ovb, z.double(), noise.double())
    ((plogits * w1).sum() + (phard * w2).sum() + (psoft * w3).sum()).backward()
    oG32b, lb32b, vb32b = T._fp32_twin(oG, olb, ovb)
    ql, qh, qs = oG32b(lb32b, vb32b, z, noise)
    ((ql * w1.float()).sum() + (qh * w2.float()).sum() + (qs * w3.float()).sum()).backward()
    ((logits * w1.float().to(DEV)).sum() + (hard * w2.float().to(DEV)).sum() + (soft * w3.float().to(DEV)).sum()).backward()
    o32 = dict(oG32b.named_parameters())
    gmax = max(float(op.grad.abs().max()) for op in oG.parameters() if op.grad is not None)
    rows = []
    for (k, p), (ok, op) in zip(G.named_parameters(), oG.named_parameters()):
        if op.grad is None:
            continue
        scale = float(op.grad.abs().max())
        if scale < 1e-4 * gmax:
            continue
        err = float((p.grad.double().cpu() - op.grad).abs().max())
        e32 = float((o32[k].grad.double() - op.grad).abs().max())
        rows.append((err / scale, e32 / scale, scale, k))
    rows.sort(reverse=True)
    print(kind, "seed", seed, "tc", tc, "logits rel", f"{T.rel_err(logits, ologits):.2e}", "gmax", f"{gmax:.2e}")
    for a, b, sc, k in rows[:8]:
        print(f"   {k:40s} ours {a:.1e}  fp32-oracle {b:.1e}  scale {sc:.1e}")
    # fraction of exactly-zero activations per block (ReLU after the narrow blocks)
    print("   zero fraction of x1 per block:", [round(float((c["x1"] == 0).float().mean()), 2) for c in G.debug_saved["conv"]])


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "GATV2CONV"
    for seed in (0, 1):
        for tc in (False, True):
            run(kind, seed, tc)
