"""Launch-geometry sweep of the aggregation kernels at N=1e6 (cold L2), guidance only."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from building_gan_b200 import graph, lib, synth
from building_gan_b200.benchmarks import _time, _peak
dev = torch.device("cuda", 0)
L = lib.load()
peak, _ = _peak()
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
pairs = [synth.large_grid_pair(900 + i) for i in range(10)]
_, vb = graph.collate_fn(pairs)
csr = vb.bg_csr.to(dev)
n, e = csr.num_nodes, csr.num_edges
cs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16,64,128".split(","))]
cfgs = [(512, 2, 64), (512, 2, 128), (512, 2, 256), (1024, 1, 64), (1024, 1, 128), (1024, 1, 256), (1024, 1, 512)]
for c in cs:
    h, s, d = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
    b, a1, a2 = torch.zeros(c, device=dev), torch.randn(c, device=dev), torch.randn(c, device=dev)
    g = torch.randn(n, c, device=dev)
    o, m, z = lib.gat_fwd(csr, h, s, d, b)
    lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2)
    by_f = 4 * (2 * n * c + 5 * n + e + c + 1)
    by_b = 4 * (4 * n * c + 8 * n + 3 * e + c + 2)
    for thr, cps, kb in cfgs:
        for pf in (0, 128):
            L.bg_tune(0, thr); L.bg_tune(1, cps); L.bg_tune(2, pf); L.bg_tune(3, int(pf > 0)); L.bg_tune(5, kb)
            tf = _time(lambda: lib.gat_fwd(csr, h, s, d, b), flush, reps=5)
            tb = _time(lambda: lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2), flush, reps=5)
            print(f"C={c:3d} thr={thr:4d} cps={cps} kb={kb} pf={pf}: fwd {tf*1e6:8.1f} us {by_f/tf/1e9:7.1f} GB/s ({by_f/tf/1e9/peak:.3f})"
                  f"   bwd {tb*1e6:8.1f} us {by_b/tb/1e9:7.1f} GB/s ({by_b/tb/1e9/peak:.3f})", flush=True)
