"""Host-side profile of PIPELINED graphed training steps over the bench's six batch shapes (CPU activities: which runtime calls block)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from building_gan_b200 import Configuration, graphs
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from building_gan_b200.optim import Adam
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
cfg = Configuration()
torch.manual_seed(777)
G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
host = bench._make_batches(0, 6, 32, pin=False)
res = [bench._clone_to(*h, dev) for h in host]
gs = graphs.GraphedStep(G, D, og, od, cfg)
for i in range(24):
    gs(*res[i % 6], sync_losses=False)
torch.cuda.synchronize()
per = []
for i in range(36):
    t = time.perf_counter()
    gs(*res[i % 6], sync_losses=False)
    per.append(round(1e3 * (time.perf_counter() - t), 2))
torch.cuda.synchronize()
print("host ms per step:", per)
print("memory: reserved %.1f MB, allocated %.1f MB, num_alloc_retries %d" % (torch.cuda.memory_reserved() / 2**20, torch.cuda.memory_allocated() / 2**20,
      torch.cuda.memory_stats()["num_alloc_retries"]), "segments", torch.cuda.memory_stats()["segment.all.current"])
with profile(activities=[ProfilerActivity.CPU]) as prof:
    for i in range(12):
        gs(*res[i % 6], sync_losses=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=50))
print("segments after", torch.cuda.memory_stats()["segment.all.current"], "reserved %.1f MB" % (torch.cuda.memory_reserved() / 2**20))
