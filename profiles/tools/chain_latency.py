"""Per-kernel latency ON A DEPENDENCY CHAIN at the training step's size (batch 32, N ~ 15 k, L2-resident): each op is
captured R times back to back into one CUDA graph (every launch depends on the previous one through its buffers / stream
order) and the replay is timed: us per launch = what the op costs on the step's critical path.
    python profiles/tools/chain_latency.py [R]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench
from building_gan_b200 import lib

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
(lb, vb), = bench._make_batches(0, 1, 32, pin=False)
lb, vb = lb.to(dev), vb.to(dev)
csr = vb.bg_csr
n, e = csr.num_nodes, csr.num_edges
print(f"N={n} E'={e} R={R}")


def chain(name, fn, reps=R):
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    us = 1e3 * sorted(ts)[2] / reps
    print(f"{name:58s} {us:7.2f} us / launch-group")
    return us


x1 = torch.randn(16, device=dev)
chain("axpy 16 elements (launch + dependency floor)", lambda: lib.axpy_(x1, x1, 0.0))
for c in (8, 64):
    a, b = torch.randn(n, c, device=dev), torch.randn(n, c, device=dev)
    chain(f"axpy [N,{c}]", lambda: lib.axpy_(a, b, 0.5))
for k, c in ((64, 64), (64, 32), (32, 16), (16, 8), (8, 16), (36, 64), (8, 1), (128, 64)):
    x, w = torch.randn(n, k, device=dev), torch.randn(c, k, device=dev) * 0.1
    bias = torch.zeros(c, device=dev)
    out = torch.empty(n, c, device=dev)
    chain(f"dense_fwd {k}->{c} (+bias, relu)", lambda: lib.dense_fwd([x], w, bias, None, 1, out=out))
    a1, a2 = torch.randn(c, device=dev), torch.randn(c, device=dev)
    if c > 1:
        chain(f"dense_fwd {k}->{c} (conv lin + att dots)", lambda: lib.dense_fwd([x], w, att=(a1, a2), out=out))
    gz = torch.randn(n, c, device=dev)
    gx = torch.empty(n, k, device=dev)
    chain(f"dgrad {c}->{k} (transposed, gated)", lambda: lib.dense_fwd([gz], w, transposed=True, out=gx, gate=x))
for c in (8, 16, 32, 64):
    h, s, d = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
    b, a1, a2 = torch.zeros(c, device=dev), torch.randn(c, device=dev), torch.randn(c, device=dev)
    one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    g = torch.randn(n, c, device=dev)
    chain(f"gat_fwd C={c}", lambda: lib.gat_fwd(csr, h, s, d, b))
    chain(f"gat_fwd_gn + apply C={c} (2 launches)", lambda: lib.gat_fwd_gn(csr, h, s, d, b, one, zero, one, None, 0.8, 1, 2))
    chain(f"gat_fwd_gn alone C={c}", lambda: lib.gat_fwd_gn(csr, h, s, d, b, one, zero, one, None, 0.8, 1, 2, apply=False))
    o, m, z = lib.gat_fwd(csr, h, s, d, b)
    x1_, stats = lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2)
    chain(f"graphnorm_fwd C={c} (2 launches)", lambda: lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2))
    chain(f"gat_bwd C={c} (2 launches)", lambda: lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2))
    chain(f"gn_bwd_moments + gat_bwd_gn C={c} (3 launches)", lambda: lib.gat_bwd_gn(csr, g, o, x1_, one, one, stats, 1.25, h, s, d, m, z, a1, a2))
    Ht, St, Dt = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
    chain(f"gat_bwd2 C={c} (2 launches)", lambda: lib.gat_bwd2(csr, Ht, St, Dt, g, h, s, d, m, z, a1, a2))
    go, dp, bst = lib.graphnorm_bwd(g, o, x1_, one, one, stats, 1.25)
    chain(f"graphnorm_bwd2 C={c} (2 launches)", lambda: lib.graphnorm_bwd2(Ht, g, o, x1_, one, one, stats, bst, 1.25))
