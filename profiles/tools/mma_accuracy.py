"""Accuracy of the dense kernels against fp64 on the same problem: warp-MMA 3xTF32 vs tiled FFMA (vs tcgen05 where eligible)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import lib
dev = "cuda"
torch.manual_seed(0)
n = 15145
for k, c in ((64, 64), (64, 32), (32, 16), (16, 8), (128, 64)):
    for kind in ("randn", "relu-like (half zeros, positive)", "large common offset"):
        x = torch.randn(n, k, dtype=torch.float64)
        if kind.startswith("relu"):
            x = torch.relu(x)
        elif kind.startswith("large"):
            x = x + 30.0
        W = torch.randn(c, k, dtype=torch.float64) * 0.2
        ref = x @ W.t()
        out = {}
        for name, on in (("mma", True), ("ffma", False)):
            lib.set_dense_mma(on)
            y = lib.dense_fwd([x.float().to(dev)], W.float().to(dev))["out"].double().cpu()
            err = (y - ref).abs()
            out[name] = (float(err.max() / ref.abs().max()), float((err / ref.abs().clamp_min(1e-3 * ref.abs().max())).max()),
                         float(((y - ref)).mean() / ref.abs().mean()))
        lib.set_dense_mma(True)
        print(f"{k:3d}->{c:2d} {kind:34s} mma max-rel {out['mma'][0]:.2e} elementwise {out['mma'][1]:.2e} bias {out['mma'][2]:+.1e} | "
              f"ffma max-rel {out['ffma'][0]:.2e} elementwise {out['ffma'][1]:.2e} bias {out['ffma'][2]:+.1e}")
