"""GraphNorm kernels at N=1e6 (cold L2), guidance only."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from building_gan_b200 import lib
from building_gan_b200.benchmarks import _time, _peak
dev = torch.device("cuda", 0)
lib.load()
peak, _ = _peak()
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
n = 1_000_000
for c in (16, 64, 128):
    o, g = torch.randn(n, c, device=dev), torch.randn(n, c, device=dev)
    one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    x1, stats = lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2)
    lib.graphnorm_bwd(g, o, x1, one, one, stats, 1.25)
    tf = _time(lambda: lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2), flush)
    tb = _time(lambda: lib.graphnorm_bwd(g, o, x1, one, one, stats, 1.25), flush)
    bf, bb = 4 * (2 * n * c + 4 * c), 4 * (3 * n * c + 6 * c)
    print(f"C={c:3d} fwd {tf*1e6:7.1f} us {bf/tf/1e9:7.1f} GB/s ({bf/tf/1e9/peak:.3f})  bwd {tb*1e6:7.1f} us {bb/tb/1e9:7.1f} GB/s ({bb/tb/1e9/peak:.3f})", flush=True)

# fused aggregation + statistics vs separate kernels (N=1e6)
from building_gan_b200 import graph, synth
pairs = [synth.large_grid_pair(900 + i) for i in range(10)]
_, vb = graph.collate_fn(pairs)
csr = vb.bg_csr.to(dev)
n, e = csr.num_nodes, csr.num_edges
for c in (16, 64, 128):
    h, s, d = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
    b = torch.zeros(c, device=dev); one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    lib.gat_fwd_gn(csr, h, s, d, b, one, zero, one, None, 0.8, 1, 2)
    def sep():
        o, m, z = lib.gat_fwd(csr, h, s, d, b)
        lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2)
    t_sep = _time(sep, flush)
    t_fused = _time(lambda: lib.gat_fwd_gn(csr, h, s, d, b, one, zero, one, None, 0.8, 1, 2), flush)
    t_gat = _time(lambda: lib.gat_fwd(csr, h, s, d, b), flush)
    by = 4 * (2 * n * c + 5 * n + e + c + 1) + 4 * (2 * n * c + 4 * c)
    print(f"C={c:3d} gat_fwd alone {t_gat*1e6:7.1f} us | gat_fwd + graphnorm_fwd separate {t_sep*1e6:7.1f} us ({by/t_sep/1e9/peak:.3f})  fused {t_fused*1e6:7.1f} us ({by/t_fused/1e9/peak:.3f})", flush=True)
