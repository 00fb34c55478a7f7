"""Host-side profile of graphed training steps with the GPU drained between steps (torch.profiler, CPU + CUDA runtime calls)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from building_gan_b200 import Configuration, graphs, step as _step
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from building_gan_b200.optim import Adam
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
cfg = Configuration()
torch.manual_seed(777)
G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
host = bench._make_batches(0, 1, 32, pin=False)
lb, vb = bench._clone_to(*host[0], dev)
gs = graphs.GraphedStep(G, D, og, od, cfg)
for _ in range(6):
    gs(lb, vb, sync_losses=False)
torch.cuda.synchronize()
acc = {}
def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[label or name] = acc.get(label or name, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, w)
wrap(_step, "SideLoss"); wrap(_step, "generator_loss"); wrap(torch.Tensor, "backward"); wrap(gs.opt_g, "step", "opt_g.step")
wrap(gs, "_capture_sampling"); wrap(gs, "_capture_critic"); wrap(gs, "_run")
wrap(torch.cuda.CUDAGraph, "replay")
K = 5
for _ in range(K):
    gs(lb, vb, sync_losses=False)
    torch.cuda.synchronize()
print({k: round(1e3 * v / K, 3) for k, v in acc.items()}, "ms per step, GPU drained between steps")
with profile(activities=[ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        gs(lb, vb, sync_losses=False)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
