"""Quick kernel-time breakdown of one training step with torch.profiler (CUPTI) - guidance only."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch, time
import bench
from building_gan_b200 import Configuration, lib, step
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
cfg = Configuration()
torch.manual_seed(777)
G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
og = torch.optim.Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS)
od = torch.optim.Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
host = bench._make_batches(0, 1, 32, pin=False)
lb, vb = bench._clone_to(*host[0], dev)
for _ in range(3):
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"] if hasattr(prof.key_averages()[0], "device_type") else prof.key_averages()
rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print("total device us", tot)
for e in rows[:32]:
    if e.self_device_time_total <= 0: continue
    print(f"{e.key[:80]:80s} {e.count:6d} {e.self_device_time_total:10.0f} us {e.self_device_time_total/max(e.count,1):8.2f} {100*e.self_device_time_total/tot:5.1f}%")
