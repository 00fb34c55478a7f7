import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_models_gpu as T
from building_gan_b200 import executor as ex, lib
cfg, G, D, oG, oD, lb, vb, olb, ovb = T._setup()
n = vb.num_nodes
z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5))
noise = -torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()
G.eval(), oG.eval()
T._inject_masks(oG, [None]*14)
w1 = torch.randn(n, 7, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
og = {}
fw = {}
for k in range(14):
    def mk(k):
        def hook_conv(mod, inp, out):
            fw[("o", k)] = out.detach().clone()
            out.register_hook(lambda g: og.__setitem__(("go", k), g.clone()))
        def hook_drop(mod, inp, out):
            fw[("x1", k)] = out.detach().clone()
            out.register_hook(lambda g: og.__setitem__(("gx1", k), g.clone()))
        return hook_conv, hook_drop
    hc, hd = mk(k)
    getattr(oG.encoder, f"module_{4*k}").register_forward_hook(hc)
    getattr(oG.encoder, f"module_{4*k+3}").register_forward_hook(hd)
ologits, ohard, osoft = oG(olb, ovb, z.double(), noise.double())
(ologits * w1).sum().backward()
rec = {}
orig_gn = lib.graphnorm_bwd; orig_gat = lib.gat_bwd
cnt = {"k": 13}
def spy_gn(gx1, o, x1, *a, **kw):
    out = orig_gn(gx1, o, x1, *a, **kw)
    k = cnt["k"]; rec[("gx1", k)] = gx1.clone(); rec[("go", k)] = out[0].clone(); rec[("o", k)] = o.clone(); rec[("x1", k)] = x1.clone()
    cnt["k"] -= 1
    return out
lib.graphnorm_bwd = spy_gn
logits, hard, soft = G(lb, vb, z.to("cuda"), noise.to("cuda"), keeps=[None]*14)
(logits * w1.float().cuda()).sum().backward()
def rel(a, b): return (a.double().cpu()-b).abs().max().item()/max(b.abs().max().item(), 1e-30)
for k in range(13, -1, -1):
    print(k, "C", rec[("o",k)].shape[1], "o", f"{rel(rec[('o',k)], fw[('o',k)]):.1e}", "x1", f"{rel(rec[('x1',k)], fw[('x1',k)]):.1e}",
          "gx1", f"{rel(rec[('gx1',k)], og[('gx1',k)]):.1e}", "go", f"{rel(rec[('go',k)], og[('go',k)]):.1e}",
          "frac x1>0", f"{(fw[('x1',k)]>0).double().mean().item():.3f}", "min|y|", f"{fw[('x1',k)][fw[('x1',k)]>0].min().item():.1e}")
