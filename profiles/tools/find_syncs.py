"""Where does the host synchronise inside one graphed training step?  (torch.cuda.set_sync_debug_mode + warnings as errors)"""
import sys, os, warnings, traceback
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
host = bench._make_batches(0, 1, 32, pin=False)
lb, vb = bench._clone_to(*host[0], dev)
gs = graphs.GraphedStep(G, D, og, od, cfg)
for _ in range(4):
    gs(lb, vb, sync_losses=False)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    gs(lb, vb, sync_losses=False)
torch.cuda.set_sync_debug_mode("default")
print(len(w), "synchronising calls in one step")
for x in w:
    print(x.filename, x.lineno, str(x.message)[:120])
torch.cuda.set_sync_debug_mode("error")
try:
    gs(lb, vb, sync_losses=False)
except Exception:
    traceback.print_exc()
