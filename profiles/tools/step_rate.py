"""Steady-state rate of the graphed training step over the bench's six batch shapes: BLOCKS timed blocks of STEPS steps after a
long warm-up (graph pools / allocator at steady state), CUDA events, no L2 flush.  The quick A/B harness for step-level changes."""
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
host = bench._make_batches(0, 6, int(os.environ.get("BATCH", "32")), pin=False)
res = [bench._clone_to(*h, dev) for h in host]
gs = graphs.GraphedStep(G, D, og, od, cfg)
AHEAD = os.environ.get("AHEAD", "1") == "1"
STEPS, BLOCKS = int(os.environ.get("STEPS", "30")), int(os.environ.get("BLOCKS", "3"))
for i in range(24):
    gs(*res[i % 6], sync_losses=False)
torch.cuda.synchronize()
out = []
for b in range(BLOCKS):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(STEPS):
        gs(*res[i % 6], sync_losses=False, next_batch=res[(i + 1) % 6] if AHEAD else None)
    e1.record()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    out.append((e0.elapsed_time(e1) / STEPS, 1e3 * th / STEPS))
if os.environ.get("DRAIN") == "1":  # pure host cost per step: GPU drained before every step, host time of the call only
    hs = []
    for i in range(STEPS):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gs(*res[i % 6], sync_losses=False, next_batch=res[(i + 1) % 6] if AHEAD else None)
        hs.append(1e3 * (time.perf_counter() - t0))
    torch.cuda.synchronize()
    hs.sort()
    print("host ms per step with the GPU drained before each step: median %.3f, min %.3f" % (hs[len(hs) // 2], hs[0]))
print("ms/step (device, host enqueue) per block:", [(round(a, 3), round(b, 3)) for a, b in out],
      "best %.3f ms = %.1f steps/s" % (min(a for a, _ in out), 1e3 / min(a for a, _ in out)))
