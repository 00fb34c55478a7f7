"""gat_fwd at HBM size (10 x 1e5-voxel grids, N = 1e6): register-path kernel vs the TMA-gather variant, cold L2, CUDA events."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import graph, lib, synth
from building_gan_b200.benchmarks import _peak, _time
dev = torch.device("cuda", 0)
peak, _ = _peak()
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
for ngraphs in (1, 10):
    _, vb = graph.collate_fn([synth.large_grid_pair(900 + i) for i in range(ngraphs)])
    csr = vb.bg_csr.to(dev)
    n, e = csr.num_nodes, csr.num_edges
    for c in (64, 128):
        h, s, d, b = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev), torch.zeros(c, device=dev)
        by = 4 * (2 * n * c + 5 * n + e + c + 1)
        row = {}
        for name, fn in (("register", lambda: lib.gat_fwd(csr, h, s, d, b)), ("tma_gather4", lambda: lib.gat_fwd_tma(csr, h, s, d, b))):
            fn()
            t = _time(fn, flush)
            row[name] = {"us": round(t * 1e6, 1), "GBs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)}
        print(json.dumps({"N": n, "C": c, **row}))
