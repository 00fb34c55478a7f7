"""One training step under cudaProfilerStart/Stop for ncu (--profile-from-start off)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
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
if os.environ.get("GRAPH", "1") == "1":
    from building_gan_b200.graphs import GraphedStep
    gs = GraphedStep(G, D, og, od, cfg)
    run = lambda: gs(lb, vb, sync_losses=False)
else:
    run = lambda: step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False)
for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("N", vb.num_nodes, "E'", vb.bg_csr.num_edges)
