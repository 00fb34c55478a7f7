import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_models_gpu as T
cfg, G, D, oG, oD, lb, vb, olb, ovb = T._setup()
n = vb.num_nodes
z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5))
noise = -torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()
G.eval(), oG.eval()
T._inject_masks(oG, [None]*14)
w1 = torch.randn(n, 7, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
ologits, ohard, osoft = oG(olb, ovb, z.double(), noise.double())
(ologits * w1).sum().backward()
runs = []
for rep in range(3):
    G.zero_grad()
    logits, hard, soft = G(lb, vb, z.to("cuda"), noise.to("cuda"), keeps=[None]*14)
    (logits * w1.float().cuda()).sum().backward()
    runs.append({k: p.grad.clone() for k, p in G.named_parameters()})
for k, p in oG.named_parameters():
    a = runs[0][k].double().cpu(); b = p.grad
    same = all(torch.equal(runs[0][k], r[k]) for r in runs[1:])
    e = (a-b).abs().max().item()/max(b.abs().max().item(),1e-30)
    if e > 1e-4 or not same:
        print(f"{k:45s} rel {e:.2e} scale {b.abs().max().item():.2e} deterministic={same}")
print("done")
