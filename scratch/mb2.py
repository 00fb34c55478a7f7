import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from building_gan_b200 import lib
from torch.profiler import profile, ProfilerActivity
dev = "cuda"
for N in (1000, 4000, 15145, 60000):
    for cout, k in ((2, 32), (64, 64), (128, 128)):
        gz = torch.randn(N, cout, device=dev); x = torch.randn(N, k, device=dev)
        dW = torch.empty(cout, k, device=dev)
        for _ in range(3): lib.dense_wgrad(gz, [x], dW=dW)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10): lib.dense_wgrad(gz, [x], dW=dW)
            torch.cuda.synchronize()
        d = {e.key[:30]: e.self_device_time_total / e.count for e in prof.key_averages()}
        print(N, cout, k, {k_: round(v, 1) for k_, v in d.items()})
