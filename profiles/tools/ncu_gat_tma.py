"""ncu target: gat_fwd at HBM size (N = 1e6, C = 64): register-path kernel, then the TMA-gather (tile::gather4) variant."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import graph, lib, synth
dev = torch.device("cuda", 0)
_, vb = graph.collate_fn([synth.large_grid_pair(900 + i) for i in range(10)])
csr = vb.bg_csr.to(dev)
n, c = csr.num_nodes, 64
h, s, d, b = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev), torch.zeros(c, device=dev)
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
for _ in range(2):
    flush.zero_()
    lib.gat_fwd(csr, h, s, d, b)
    flush.zero_()
    lib.gat_fwd_tma(csr, h, s, d, b)
torch.cuda.synchronize()
