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
from building_gan_b200.optim import Adam as _A
_A = torch.optim.Adam if os.environ.get("ADAM") == "torch" else _A
og = _A(G.parameters(), lr=2e-4, betas=cfg.BETAS)
od = _A(D.parameters(), lr=2e-4, betas=cfg.BETAS)
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
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if str(e.device_type).endswith("CUDA"):
        agg[e.name][0] += 1; agg[e.name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if "bg::" in k)
print(f"GPU kernel time total {tot/1e3:.2f} ms, launches {sum(v[0] for v in agg.values())}; libbgb200: {ours/1e3:.2f} ms in {sum(v[0] for k, v in agg.items() if 'bg::' in k)} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[:90]:90s} {c:5d} {t:9.0f} us {t/c:8.2f} {100*t/tot:5.1f}%")
