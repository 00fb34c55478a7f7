import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_models_gpu as T
from building_gan_b200 import lib
for train in (False, True):
    cfg, G, D, oG, oD, lb, vb, olb, ovb = T._setup()
    n = vb.num_nodes
    z = torch.randn(1, n, cfg.Z_DIM, generator=torch.Generator().manual_seed(5))
    noise = -torch.empty(n, 7).exponential_(generator=torch.Generator().manual_seed(6)).log()
    keeps = T._keeps(n, T.G_WIDTHS, 7) if train else [None] * 14
    G.train(train), oG.train(train)
    T._inject_masks(oG, keeps)
    grabbed = {}
    def hook(mod, inp, out):
        out.register_hook(lambda g: grabbed.__setitem__("g_enc", g))
    hdl = oG.matched_features_encoder.register_forward_hook(hook)
    ologits, ohard, osoft = oG(olb, ovb, z.double(), noise.double())
    w1 = torch.randn(n, 7, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
    (ologits * w1).sum().backward()
    ge_ref = torch.zeros(7, 128, dtype=torch.float64).index_add_(0, ovb.type, grabbed["g_enc"])
    kk = [None if k is None else k.to(torch.uint8).to("cuda") for k in keeps]
    # monkeypatch type_scatter_sum to record
    rec = []
    orig = lib.type_scatter_sum
    def spy(g, t, k, width=None):
        out = orig(g, t, k, width); rec.append((g.clone(), out.clone())); return out
    lib.type_scatter_sum = spy
    logits, hard, soft = G(lb, vb, z.to("cuda"), noise.to("cuda"), keeps=kk)
    (logits * w1.float().cuda()).sum().backward()
    lib.type_scatter_sum = orig
    tot = rec[0][1] + rec[1][1]
    print("train", train, "ge err", (tot.double().cpu() - ge_ref).abs().max().item(), "scale", ge_ref.abs().max().item())
    for g, out in rec:
        want = torch.zeros(7, 128, dtype=torch.float64).index_add_(0, ovb.type, g.double().cpu())
        print("   scatter self-check err", (out.double().cpu() - want).abs().max().item(), want.abs().max().item())
    print("   types present", torch.bincount(ovb.type, minlength=7).tolist(), "local", torch.bincount(olb.type, minlength=7).tolist())
    for k in ["matched_features_encoder.12.weight", "matched_features_encoder.0.weight", "mlp_encoder.0.weight"]:
        a = dict(G.named_parameters())[k].grad.double().cpu(); b = dict(oG.named_parameters())[k].grad
        print("   ", k, (a-b).abs().max().item(), b.abs().max().item())
