"""Host vs GPU time of the graphed training step, with a host-side breakdown of one step."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from building_gan_b200 import Configuration, lib, step, graphs, models
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from building_gan_b200.optim import Adam
dev = torch.device("cuda", 0)
cfg = Configuration()
torch.manual_seed(777)
G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
host = bench._make_batches(0, 1, int(os.environ.get("BATCH", "32")), pin=False)
lb, vb = bench._clone_to(*host[0], dev)
gs = graphs.GraphedStep(G, D, og, od, cfg)
for _ in range(4):
    gs(lb, vb, sync_losses=False)
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
for _ in range(K):
    gs(lb, vb, sync_losses=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/K:.2f} ms/step, complete {1e3*(t2-t0)/K:.2f} ms/step")
# host-side breakdown: wrap the pieces
import types
acc = {}
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, w)
wrap(gs, "_capture_sampling"); wrap(gs, "_capture_critic"); wrap(gs, "_run")
torch.cuda.synchronize()
for _ in range(K):
    gs(lb, vb, sync_losses=False)
    torch.cuda.synchronize()  # drained GPU: pure host costs
print({k: round(1e3 * v / K, 3) for k, v in acc.items()}, "ms per step (GPU drained between steps)")
# GPU time of a step alone (drained before, synchronised after)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t = time.perf_counter()
    gs(lb, vb, sync_losses=False)
    th = time.perf_counter() - t
    torch.cuda.synchronize()
    ts.append((th, time.perf_counter() - t))
print("single step: host %.2f ms, host+drain %.2f ms" % (1e3 * sum(a for a, _ in ts) / 10, 1e3 * sum(b for _, b in ts) / 10))
