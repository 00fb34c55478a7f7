import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from building_gan_b200 import Configuration, lib, step
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
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
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
