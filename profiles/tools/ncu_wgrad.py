import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from building_gan_b200 import lib
dev = "cuda"
N = 15145
gz = torch.randn(N, 64, device=dev); x = torch.randn(N, 64, device=dev)
dW = torch.empty(64, 65, device=dev)
for _ in range(4):
    lib.dense_wgrad(gz, [x, None], dW=dW)
torch.cuda.synchronize()
print("ok")
