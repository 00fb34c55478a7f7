import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from building_gan_b200 import graph, lib, synth
dev="cuda"
pairs = [synth.large_grid_pair(900 + i) for i in range(10)]
_, vb = graph.collate_fn(pairs); csr = vb.bg_csr.to(dev); n, e = csr.num_nodes, csr.num_edges
flush = torch.empty((256 << 20) // 4, device=dev)
def t(fn, reps=7):
    ts=[]
    for _ in range(reps):
        flush.zero_(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]*1e3
for c in (1, 8, 32, 64, 128):
    h, s, d = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
    b, a1, a2 = torch.zeros(c, device=dev), torch.randn(c, device=dev), torch.randn(c, device=dev)
    g = torch.randn(n, c, device=dev)
    o, m, z = lib.gat_fwd(csr, h, s, d, b); lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2)
    tf = t(lambda: lib.gat_fwd(csr, h, s, d, b)); tb = t(lambda: lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2))
    bf = 4*(2*n*c+5*n+e+c+1); bb = 4*(4*n*c+8*n+3*e+c+2)
    print(f"BIG_N={os.environ.get('BG_GAT_BIG_N','off')} C={c:4d} fwd {tf:8.1f}us {bf/tf/1e3:7.1f}GB/s | bwd {tb:8.1f}us {bb/tb/1e3:7.1f}GB/s")
