"""Per-step wall times of the graphed step over the bench's 6 cycling batches (diagnostic)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from building_gan_b200 import Configuration, graphs
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from building_gan_b200.optim import Adam
dev = torch.device("cuda", 0)
cfg = Configuration()
torch.manual_seed(777)
G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
host = bench._make_batches(0, 6, 32, pin=True)
res = [bench._clone_to(lb, vb, dev) for lb, vb in host]
gs = graphs.GraphedStep(G, D, og, od, cfg)
clocks = None
if os.environ.get("CLOCKS", "0") == "1":
    clocks = bench._Clocks(0); clocks.start()
flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
ts = []
for i in range(40):
    t = time.perf_counter()
    flush.zero_()
    gs(*res[i % 6], sync_losses=False)
    if i % 10 == 9:
        torch.cuda.synchronize()
    ts.append(time.perf_counter() - t)
torch.cuda.synchronize()
if clocks: print(clocks.stop())
print(" ".join(f"{1e3*t:.1f}" for t in ts))
print("mem allocated %.0f MB reserved %.0f MB" % (torch.cuda.memory_allocated() / 2**20, torch.cuda.memory_reserved() / 2**20))
