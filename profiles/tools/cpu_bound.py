"""Is the training step CPU- or GPU-bound?  Enqueue time vs completion time, plus a per-phase CPU breakdown."""
import sys, os, time
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
for _ in range(3):
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False, overlap=(os.environ.get("OVERLAP", "1") == "1"))
torch.cuda.synchronize()
K = 10
t0 = time.perf_counter()
for _ in range(K):
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False, overlap=(os.environ.get("OVERLAP", "1") == "1"))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/K:.2f} ms/step, complete {1e3*(t2-t0)/K:.2f} ms/step")
# CPU-only cost of pieces (GPU drained before each measurement so nothing blocks)
def cpu_ms(fn, n=2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    dt = time.perf_counter() - t
    torch.cuda.synchronize()
    return 1e3 * dt / n
z = torch.randn(1, vb.num_nodes, cfg.Z_DIM, device=dev)
def gfwd():
    with torch.no_grad():
        return G(lb, vb, z)
_, hard, soft = gfwd()
hard, soft = hard.unsqueeze(0), soft.unsqueeze(0)
print("G fwd (no grad) cpu ms", cpu_ms(gfwd))
def dloss():
    od.zero_grad()
    l = step.discriminator_loss(D, lb, vb, hard, soft, cfg, rng="device")
    l.backward()
print("critic loss+backward cpu ms", cpu_ms(dloss))
print("adam D cpu ms", cpu_ms(od.step))
print("adam G cpu ms", cpu_ms(og.step))
def gupd():
    logits, h, s = G(lb, vb, z)
    og.zero_grad()
    l = step.generator_loss(D, lb, vb, logits, h.unsqueeze(0), cfg)
    l.backward()
print("generator loss+backward cpu ms", cpu_ms(gupd, 1))
