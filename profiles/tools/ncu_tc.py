import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import lib
dev = "cuda"; N = 15145
for cin in (128, 524):
    x = torch.randn(N, cin, device=dev); W = torch.randn(128, cin, device=dev); b = torch.randn(128, device=dev)
    for _ in range(3):
        lib.dense_fwd([x], W, b, None, 1)
torch.cuda.synchronize(); print("ok")
