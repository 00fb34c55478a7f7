"""gat_fwd / gat_bwd at N=1e6 (10 x 1e5-voxel grids), C=64 - for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import graph, lib, synth
dev = "cuda"
C = int(os.environ.get("C", 64))
pairs = [synth.large_grid_pair(900 + i) for i in range(10)]
_, vb = graph.collate_fn(pairs)
csr = vb.bg_csr.to(dev)
n = csr.num_nodes
h, s, d = torch.randn(n, C, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
b, a1, a2 = torch.zeros(C, device=dev), torch.randn(C, device=dev), torch.randn(C, device=dev)
g = torch.randn(n, C, device=dev)
flush = torch.empty((256 << 20) // 4, device=dev)
for _ in range(3):
    flush.zero_()
    o, m, z = lib.gat_fwd(csr, h, s, d, b)
    lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2)
torch.cuda.synchronize()
print("ok", n, csr.num_edges, C)
