"""ncu target: the dense kernels at the training step's size (N = 15145): row-per-thread 64->64, tiled 64->64, tcgen05 128->128."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import lib
dev = torch.device("cuda", 0)
n = 15145
x64, w64 = torch.randn(n, 64, device=dev), torch.randn(64, 64, device=dev) * 0.1
x128, w128 = torch.randn(n, 128, device=dev), torch.randn(128, 128, device=dev) * 0.1
b64, b128 = torch.zeros(64, device=dev), torch.zeros(128, device=dev)
g128, be128 = torch.ones(128, device=dev), torch.zeros(128, device=dev)
o64, o128 = torch.empty(n, 64, device=dev), torch.empty(n, 128, device=dev)
for _ in range(3):
    lib.set_dense_mma(True)
    lib.dense_fwd([x64], w64, b64, None, 1, out=o64)       # warp-MMA 3xTF32
    lib.set_dense_mma(False)
    lib.set_rowdense(2)
    lib.dense_fwd([x64], w64, b64, None, 1, out=o64)       # row-per-thread FFMA
    lib.set_rowdense(0)
    lib.dense_fwd([x64], w64, b64, None, 1, out=o64)       # tiled FFMA
    lib.set_rowdense(1)
    lib.set_dense_mma(True)
    lib.dense_fwd([x128], w128, b128, (g128, be128), 2, out=o128)  # tcgen05 3xTF32
torch.cuda.synchronize()
