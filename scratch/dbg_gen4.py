import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_models_gpu as T
from building_gan_b200 import executor as ex, lib
cfg, G, D, oG, oD, lb, vb, olb, ovb = T._setup()
n = vb.num_nodes
print("N", n)
z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5))
noise = -torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()
G.eval()
w1 = torch.randn(n, 7, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
orig_gn = lib.graphnorm_bwd
cnt = {"k": 13}
def spy_gn(gx1, o, x1, w, alpha, stats, scale, dpar=None, accumulate=False):
    out = orig_gn(gx1, o, x1, w, alpha, stats, scale, dpar, accumulate)
    go, dp, bst = out
    k = cnt["k"]; cnt["k"] -= 1
    C = o.shape[1]; N = o.shape[0]
    d = lambda t: t.double()
    mu, r = d(stats[:C]), d(stats[C:2*C])
    gy = torch.where(x1 > 0, d(gx1) * scale, torch.zeros_like(d(gx1)))
    oh = d(o) - d(alpha) * mu
    G0, G1 = gy.mean(0), (gy * oh).mean(0)
    ohat = d(w) * r * gy - d(w) * r**3 * G1 * oh
    go_ref = ohat - d(alpha) * ohat.mean(0)
    rel = lambda a, b: ((d(a) - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    # also re-run the kernel standalone on clones
    go2, dp2, bst2 = orig_gn(gx1.clone(), o.clone(), x1.clone(), w.clone(), alpha.clone(), stats.clone(), scale)
    print(k, "C", C, "go", f"{rel(go, go_ref):.1e}", "G0", f"{rel(bst[:C], G0):.1e}", "G1", f"{rel(bst[C:], G1):.1e}",
          "standalone go", f"{rel(go2, go_ref):.1e}", "o aligned", o.data_ptr() % 16, gx1.data_ptr() % 16, x1.data_ptr() % 16,
          "contig", gx1.is_contiguous(), o.is_contiguous(), x1.is_contiguous(), "w ptr%16", w.data_ptr() % 16, alpha.data_ptr() % 16, stats.data_ptr()%16)
    return out
lib.graphnorm_bwd = spy_gn
logits, hard, soft = G(lb, vb, z.to("cuda"), noise.to("cuda"), keeps=[None]*14)
(logits * w1.float().cuda()).sum().backward()
