import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import test_models_gpu as T
import test_conv_types_gpu as C
from building_gan_b200 import lib
DEV="cuda"
def run(kind, seed, tc):
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
    T._sync_patterns(oG, G.debug_saved, ovb.type)
    plogits, phard, psoft = oG(olb, ovb, z.double(), noise.double())
    ((plogits * w1).sum() + (phard * w2).sum() + (psoft * w3).sum()).backward()
    oG32b, lb32b, vb32b = T._fp32_twin(oG, olb, ovb)
    ql, qh, qs = oG32b(lb32b, vb32b, z, noise)
    ((ql * w1.float()).sum() + (qh * w2.float()).sum() + (qs * w3.float()).sum()).backward()
    ((logits * w1.float().to(DEV)).sum() + (hard * w2.float().to(DEV)).sum() + (soft * w3.float().to(DEV)).sum()).backward()
    o32 = dict(oG32b.named_parameters())
    rows = []
    for (k, p), (ok, op) in zip(G.named_parameters(), oG.named_parameters()):
        if op.grad is None: continue
        err = float((p.grad.double().cpu() - op.grad).abs().max()); scale = float(op.grad.abs().max())
        e32 = float((o32[k].grad.double() - op.grad).abs().max())
        rows.append((err / max(scale, 1e-30), e32 / max(scale, 1e-30), k))
    rows.sort(reverse=True)
    print(kind, "seed", seed, "tc", tc, "logits rel", T.rel_err(logits, ologits), "top:", [(f"{a:.1e}", f"{b:.1e}", k) for a, b, k in rows[:4]])
for kind in ("GATV2CONV", "GATCONV"):
    for seed in (0, 1, 2):
        for tc in (True, False):
            run(kind, seed, tc)
