"""Per-stream GPU timeline of graphed training steps (chrome trace of torch.profiler, kernels grouped by stream)."""
import sys, os, json, collections, tempfile
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
host = bench._make_batches(0, 1, 32, pin=False)
lb, vb = bench._clone_to(*host[0], dev)
gs = graphs.GraphedStep(G, D, og, od, cfg)
for _ in range(4):
    gs(lb, vb, sync_losses=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        gs(lb, vb, sync_losses=False)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "gt.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ks)
print(f"3 steps: span {(t1-t0)/1e3:.2f} ms, {len(ks)} GPU ops, sum of durations {sum(e['dur'] for e in ks)/1e3:.2f} ms")
per = collections.defaultdict(list)
for e in ks:
    per[e["args"].get("stream")].append(e)
for s, v in sorted(per.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in v)
    names = collections.Counter()
    for e in v: names[e["name"].split("<")[0].split("(")[0][:40]] += e["dur"]
    top = ", ".join(f"{n} {d/1e3:.2f}" for n, d in names.most_common(4))
    print(f"stream {s}: {len(v)} ops, busy {busy/1e3:.2f} ms ({100*busy/(t1-t0):.0f}% of span); top: {top}")
# union busy + concurrency histogram
evs = []
for e in ks:
    evs.append((e["ts"], 1)); evs.append((e["ts"] + e["dur"], -1))
evs.sort()
lvl, last, hist = 0, t0, collections.Counter()
for t, d in evs:
    hist[min(lvl, 4)] += t - last
    last, lvl = t, lvl + d
print("time with k kernels in flight (ms):", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
# middle step: per-kernel-name totals
mid0, mid1 = t0 + (t1 - t0) / 3, t0 + 2 * (t1 - t0) / 3
names = collections.Counter(); cnt = collections.Counter()
for e in ks:
    if mid0 <= e["ts"] < mid1:
        n = e["name"].split("(")[0][:60]
        names[n] += e["dur"]; cnt[n] += 1
print("middle third, by kernel:")
for n, d in names.most_common(25):
    print(f"  {d/1e3:7.3f} ms {cnt[n]:5d}  {n}")
