"""GPU timeline of one training step: busy time (union of kernel intervals), idle gaps and what they wait for."""
import sys, os, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from building_gan_b200 import Configuration, lib, step
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
OV = os.environ.get("OVERLAP", "1") == "1"
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
    step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False, overlap=OV)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses=False, overlap=OV)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if str(e.device_type).endswith("CUDA") and e.device_time > 0]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, cur_end, gaps = 0.0, t0, []
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, e.name[:60]))
        busy += en - s
        cur_end = en
    elif en > cur_end:
        busy += en - cur_end
        cur_end = en
span = t1 - t0
print(f"span {span/1e3:.2f} ms for 2 steps; GPU busy (union) {busy/1e3:.2f} ms = {100*busy/span:.1f}%; kernels {len(ev)}; sum of kernel times {sum(e.device_time for e in ev)/1e3:.2f} ms")
per = collections.defaultdict(lambda: [0, 0.0, 1e30, 0.0])
for e in ev:
    k = getattr(e, "stream", None) if hasattr(e, "stream") else None
    r = per[k]; r[0] += 1; r[1] += e.device_time; r[2] = min(r[2], e.time_range.start); r[3] = max(r[3], e.time_range.end)
print("per stream: " + "; ".join(f"{k}: {v[0]} kernels, {v[1]/1e3:.2f} ms busy, active {(v[2]-t0)/1e3:.2f}..{(v[3]-t0)/1e3:.2f} ms" for k, v in per.items()))
hist = collections.Counter()
tot_gap = sum(g for g, _ in gaps)
for g, _ in gaps:
    hist["<2us" if g < 2 else "2-5us" if g < 5 else "5-20us" if g < 20 else "20-100us" if g < 100 else ">100us"] += g
print("idle total %.2f ms; by gap size (ms):" % (tot_gap / 1e3), {k: round(v / 1e3, 2) for k, v in hist.items()})
by = collections.defaultdict(float)
for g, n in gaps:
    if g >= 5: by[n] += g
print("idle >=5us attributed to the NEXT kernel:")
for n, g in sorted(by.items(), key=lambda kv: -kv[1])[:14]:
    print(f"  {g/1e3:7.2f} ms  {n}")
cpu = [e for e in prof.events() if not str(e.device_type).endswith("CUDA")]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in cpu:
    agg[e.name][0] += 1; agg[e.name][1] += e.cpu_time
print("top CPU-side ops (self+children, us):")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {t/1e3:8.2f} ms {c:6d}  {n[:70]}")
