"""Kernel sequence of one captured critic update on the stream that carries the gradient-penalty chain (diagnostic)."""
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
    gs(lb, vb, sync_losses=False)
    gs(lb, vb, sync_losses=False)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "gc.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
adams = [e for e in ks if "adam_flat" in e["name"]]
# second step: adam launches 6..11 (5 critic + 1 generator); window = critic update #3 of the second step
a0, a1 = adams[7], adams[8]
w0, w1 = a0["ts"] + a0["dur"], a1["ts"] + a1["dur"]
win = [e for e in ks if w0 <= e["ts"] < w1]
print(f"critic update window {(w1-w0)/1e3:.3f} ms, {len(win)} GPU ops")
step0 = adams[5]["ts"] + adams[5]["dur"]
json.dump([[round(e["ts"] - step0, 2), e["dur"], e["args"].get("stream"), e["name"][:70]] for e in ks if e["ts"] >= step0],
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "gpurun_out", "s4_step_ops.json"), "w"))
st = collections.Counter(e["args"].get("stream") for e in win if "bwd2" in e["name"])
chain_stream = st.most_common(1)[0][0]
print("per stream in window:", {s: (len(v := [e for e in win if e['args'].get('stream') == s]), round(sum(e['dur'] for e in v) / 1e3, 3)) for s in set(e['args'].get('stream') for e in win)})
chain = [e for e in win if e["args"].get("stream") == chain_stream]
prev = w0
tot_gap = tot_dur = 0
for e in chain:
    gap = e["ts"] - prev
    tot_gap += max(gap, 0); tot_dur += e["dur"]
    print(f"{(e['ts']-w0)/1e3:8.3f} +{gap:6.1f} us gap  {e['dur']:6.1f} us  {e['name'][:90]}")
    prev = e["ts"] + e["dur"]
print(f"chain stream {chain_stream}: {len(chain)} ops, busy {tot_dur/1e3:.3f} ms, gaps {tot_gap/1e3:.3f} ms")
