#!/usr/bin/env python
"""Benchmark of the Building-GAN hot path on B200 (contract: see the task statement / DESIGN.md).

Default workload = BASELINE.json configs[1]: the reference's ``train.py`` step on "6types-like"
synthetic buildings, batch 32 per GPU: one step = trainer.py:467-495 = 5 critic updates (G forward
without grad, D(real), D(fake), gradient penalty with its second-order backward, Adam) + 1
generator update (G forward, D(fake), backward, Adam).  Metric: G+D train steps/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [...]                               # the reference algorithm on the host CPU
    python bench.py --workload sample|c4 [...]                           # extra reports (not the driver's line)

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM.  ``e2e``: the same step through the
public API from pinned HOST buffers (H2D of the batch + CSR every step, the losses read back with
.item() like trainer.py:479,493).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH = 32
NUM_BATCHES = 6          # distinct synthetic batches cycled through (each has its own N / E)
L2_FLUSH_BYTES = 256 << 20
# the ONE description of the timed workload, printed by both arms (the driver compares the two strings)
WORKLOAD = ("train.py step (trainer.py:467-495: 5 critic updates + 1 generator update), batch 32 per GPU, 6types-like synthetic "
            "buildings (mean ~400 voxels, 6-neighbour irregular grids)")
FP32_FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FMA lanes x 2 flop x max SM clock (no fp32 figure in MEASURED_PEAKS.json)


def _building_ids(rank: int, b: int, batch: int):
    base = 4001 + rank * 100_000
    return [base + b * batch + i for i in range(batch)]


def _env_int(name, default):
    return int(os.environ.get(name, default))


def _shard_ids(rank: int, world: int, b: int, batch: int):
    """Data-parallel batches: global batch b = batch x world consecutive buildings, cut into `world` contiguous shards balanced by
    voxel count (dist.shard_by_nodes) - BASELINE config 5 at 8 GPUs (global batch 256), its scaled-down versions at 2 / 4."""
    from building_gan_b200.dist import shard_by_nodes
    from workloads import synth as wsynth
    ids = [4001 + b * batch * world + i for i in range(batch * world)]
    shards = shard_by_nodes([wsynth.voxel_count(i) for i in ids], world)
    return [ids[k] for k in shards[rank]]


def _make_batches(rank: int, num_batches: int, batch: int, pin: bool, world: int = 1):
    from building_gan_b200 import graph, synth
    out = []
    for b in range(num_batches):
        ids = _building_ids(rank, b, batch) if world == 1 else _shard_ids(rank, world, b, batch)
        pairs = [synth.building_pair_fast(i) for i in ids]
        lb, vb = graph.collate_fn(pairs)
        if pin:
            lb, vb = lb.pin_memory(), vb.pin_memory()
        out.append((lb, vb))
    return out


def _batch_bytes(lb, vb) -> int:
    total = 0
    for b in (lb, vb):
        for v in b._fields.values():
            if isinstance(v, torch.Tensor):
                total += v.numel() * v.element_size()
    csr = vb._fields.get("bg_csr")
    if csr is not None:
        for f in csr.FIELDS:
            t = getattr(csr, f)
            total += t.numel() * t.element_size()
    return total


def _clone_to(lb, vb, device):
    from building_gan_b200.graph import Batch
    out = []
    for b in (lb, vb):
        nb = Batch.__new__(Batch)
        object.__setattr__(nb, "_fields", dict(b._fields))
        object.__setattr__(nb, "_cuts", object.__getattribute__(b, "_cuts"))
        object.__setattr__(nb, "_starts", object.__getattribute__(b, "_starts"))
        object.__setattr__(nb, "_pack", b.__dict__.get("_pack"))
        nb._fields.pop("_bg_cache", None)
        out.append(nb.to(device, non_blocking=True))
    return out


def _settle_heap() -> None:
    """Called (untimed) right before a timed block: collect the cyclic garbage of the setup / warm-up phase and move the
    surviving heap - models, batches, the imported libraries: ~1e6 tracked objects - to the permanent generation (gc.freeze).
    The collector stays ON; what it no longer does is re-traverse that static heap when a full collection falls into the
    block: one such pass costs 50-150 ms of host time, and the step is ~1700 driver calls with the host barely ahead of
    the GPU (single blocks of otherwise healthy runs read 91-106 steps/s instead of 130: profiles/r02b_summary.md)."""
    if os.environ.get("BG_GC_FREEZE", "1") == "1":
        gc.collect()
        gc.freeze()


class _Clocks(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.  In-process NVML (nvidia_ml_py) when
    available: a sample costs microseconds.  (Spawning `nvidia-smi` five times a second from every rank takes driver
    locks that stall kernel launches: at 8 ranks it cost the timed region 25%.)  Fallback: nvidia-smi at 1 Hz."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self._smax = index, [], threading.Event(), None
        self.nvml, self.handle, self.source = None, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        if self._smax is None:  # constant: asked once (start() is called before the timed region begins)
            self._smax = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.handle))
        self.rows.append([str(sm), str(self._smax)] + ["Active" if mask & bit else "Not Active" for _, bit in self.NVML_REASONS])

    def run(self):
        # The first NVML sample comes 60 ms into the timed region, not at its start: it then reads the clocks under load, and a
        # slow query (see below) does not land on the first step's launches, where the host has no run-ahead yet and every
        # stalled launch is idle GPU time.  A region shorter than that still gets its one sample (taken as it ends).
        if self.nvml is not None and self._stop_evt.wait(float(os.environ.get("BG_CLOCK_DELAY", "0.06"))):
            try:
                self._sample_nvml()
            except Exception:
                pass
            return
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            # 10 samples a second: on some boxes an NVML query takes ~100 ms instead of microseconds and holds a driver lock that
            # kernel launches wait for - the timed step is host-launch-sensitive (one run read 91 steps/s with the sampler's
            # first block against 131 in the block after it, profiles/r02b_summary.md)
            self._stop_evt.wait(float(os.environ.get("BG_CLOCK_PERIOD", "0.1")) if self.nvml is not None else 1.0)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        smax = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": reasons, "samples": len(sm),
                "source": self.source}


def _gat_fwd_bytes(n, e, c):
    """h, out [N,C]; s, d, m, z, rowptr [N]; col [E]; bias [C]."""
    return 4 * (2 * n * c + 5 * n + e + c + 1)


# ----------------------------------------------------------------------------------------------------
# reference arm: the oracle's restatement of trainer.py on the host cores
# ----------------------------------------------------------------------------------------------------
def _oracle_batches(num_batches: int, batch: int, rank: int = 0):
    """The SAME buildings as ``_make_batches`` (same ids, same generator: workloads/synth.py) as the oracle's own PyG-style
    ``Batch`` objects.  Nothing of the product package is imported on this path."""
    from oracle import config as oconfig, pyg as opyg
    from workloads import synth as wsynth
    out = []
    for b in range(num_batches):
        fields = [wsynth.building_fields_fast(i, oconfig.Configuration) for i in _building_ids(rank, b, batch)]
        out.append((opyg.Batch.from_data_list([opyg.Data(**f[0]) for f in fields]),
                    opyg.Batch.from_data_list([opyg.Data(**f[1]) for f in fields])))
    return out


def run_reference(args, rank: int, world: int) -> None:
    """The reference arm: the reference's algorithm (oracle restatement; torch_geometric is not installable and the
    reference has no native code to compile) on the host cores, all threads.  Imports oracle/ and workloads/ only."""
    if rank != 0:
        return
    from oracle import config as oconfig, models as omodels, trainer as otrainer
    assert "building_gan_b200" not in sys.modules, "the reference arm must not load the product package"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = oconfig.Configuration()
    cfg.BATCH_SIZE = BATCH
    cfg.DEVICE = "cpu"
    torch.manual_seed(777)
    batches = _oracle_batches(max(1, min(NUM_BATCHES, args.steps + args.warmup)), BATCH)
    G, D = omodels.OracleGenerator(cfg, 17, 12), omodels.OracleDiscriminator(cfg, 17, 12)
    og = torch.optim.Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
    od = torch.optim.Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
    for i in range(args.warmup):
        otrainer.train_step(G, D, og, od, *batches[i % len(batches)], cfg)
    _settle_heap()
    t0 = time.perf_counter()
    for i in range(args.steps):
        otrainer.train_step(G, D, og, od, *batches[(args.warmup + i) % len(batches)], cfg)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {"metric": "G+D train steps/sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "global_batch": BATCH},
            "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full steps of the oracle restatement of trainer.py:467-495 (PyG ops restated "
                                       "in torch; torch_geometric is not installable), torch threads = all host cores"},
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_package_loaded": "building_gan_b200" in sys.modules}
    print(json.dumps(line), flush=True)


def _to_oracle(batch, opyg):
    graphs = [batch[i] for i in range(batch.num_graphs)]
    return opyg.Batch.from_data_list([opyg.Data(**{k: v for k, v in g._fields.items()}) for g in graphs])


# ----------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------
def _pdl_state(lib) -> bool:
    prev = lib.set_pdl(False)
    lib.set_pdl(prev)
    return prev


def run_b200(args, rank: int, local_rank: int, world: int) -> None:
    import torch.distributed as dist
    from building_gan_b200 import Configuration, lib, step
    from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib.load()
    cfg = Configuration()
    cfg.BATCH_SIZE = BATCH
    torch.manual_seed(777)
    G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
    # H15: the reference's Adam(lr, betas) (train.py:36-37).  "flat" = building_gan_b200.optim.Adam, the same optimiser as
    # one kernel launch over the models' flat buffers; "torch" = torch.optim.Adam (foreach) on the same models.
    if args.adam == "flat":
        from building_gan_b200.optim import Adam
    else:
        Adam = torch.optim.Adam
    og = Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
    od = Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
    OVERLAP = not args.no_overlap
    grad_sync, sync_kind = None, "none"
    if world > 1:
        from building_gan_b200.dist import GradSync, PeerSync
        sync_kind = "nccl all-reduce (AVG) of the flat gradient bucket, then one-launch Adam"
        grad_sync = GradSync(world)
        if args.adam == "flat" and not args.nccl_sync:
            try:  # one-shot peer-memory all-reduce fused with the Adam step (csrc/bg_p2p.cu)
                peer = PeerSync()
                peer.attach(D, od)
                peer.attach(G, og)
                grad_sync = peer
                sync_kind = "bg_p2p_allreduce_adam: one-shot NVLink peer-memory all-reduce fused with the Adam step (one launch per update)"
            except Exception as exc:  # symmetric memory not available on this box: NCCL path
                sync_kind += f" (peer-memory path unavailable: {repr(exc)[:120]})"
    host = _make_batches(rank, NUM_BATCHES, BATCH, pin=True, world=world)
    resident = [_clone_to(lb, vb, dev) for lb, vb in host]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    h2d = sum(_batch_bytes(lb, vb) for lb, vb in host) / len(host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, fin=None, start=0):
        """``start``: index of the block's first step = the number of untimed calls of ``fn`` before it, so that the batch sequence
        simply continues: the last untimed step has then announced (next_batch=) exactly the batch the first timed step runs,
        as a prefetching loader does at every step - otherwise each block would open with a capture on an idle GPU."""
        _settle_heap()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(start, start + steps):
            flush.zero_()  # L2 flush between timed iterations (inside the timed region, ~40 us)
            fn(i)
        if fin is not None:
            fin()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    gstep = None
    if OVERLAP and args.adam == "flat" and not args.no_graph:
        from building_gan_b200.graphs import GraphedStep
        gstep = GraphedStep(G, D, og, od, cfg, grad_sync=grad_sync)  # critic update + sampling pass captured once per step, replayed N_CRITIC times

    def step_resident(i):
        lb, vb = resident[i % len(resident)]
        if gstep is not None:
            # next_batch: the following step's graphs are captured during this one (a prefetching loader knows the next batch)
            return gstep(lb, vb, sync_losses=False, next_batch=resident[(i + 1) % len(resident)])
        return step.train_step(G, D, og, od, lb, vb, cfg, rng="device", grad_sync=grad_sync, sync_losses=False, overlap=OVERLAP)

    result_sink = []
    pending = []  # (pinned host buffer, copy-done event) of steps whose losses are still on their way to the host

    def drain(keep: int):
        while len(pending) > keep:
            buf, ev = pending.pop(0)
            ev.synchronize()
            vals = buf.tolist()
            result_sink.append((vals[:-1], vals[-1]))  # the step's 6 losses (trainer.py:479,493) as floats

    ahead = {}  # step index -> its batch, already on its way to the device (one step of prefetch, like a data loader's)
    ring = [(torch.empty(cfg.N_CRITIC + 1, dtype=torch.float32, pin_memory=True), torch.cuda.Event()) for _ in range(4)]

    def step_e2e(i):
        if gstep is not None:
            # pinned host -> device, one batch per step: step i+1's copy is enqueued during step i (prefetch depth 1), so that
            # its graphs can be captured during step i as well; the first batch of a timed block is copied here
            lb, vb = ahead.pop(i) if i in ahead else _clone_to(*host[i % len(host)], dev)
            ahead.clear()
            ahead[i + 1] = _clone_to(*host[(i + 1) % len(host)], dev)
            # the 6 losses of EVERY step are read back (one async D2H into pinned memory per step); the host looks at step i's
            # floats two steps later, so capture and execution overlap.  fin_e2e() collects the last one
            # inside the timed region.
            gstep(lb, vb, sync_losses=False, next_batch=ahead[i + 1])
            buf, ev = ring[i % len(ring)]  # pinned result buffers allocated once (no cudaHostAlloc inside the timed region)
            buf.copy_(gstep.last_losses, non_blocking=True)
            ev.record()
            pending.append((buf, ev))
            # the host reads step i's floats once step i+2 is enqueued: two steps of run-ahead, the same depth GraphedStep itself
            # allows (it waits for the graphs of step i-2 before it drops them); with one step the host stalled on step i-1's
            # completion right after enqueueing step i and every bit of host jitter was idle GPU time
            drain(int(os.environ.get("BG_E2E_DEPTH", "2")))
            return
        lb, vb = _clone_to(*host[i % len(host)], dev)  # pinned host -> device, every step
        d_losses, g_loss, _ = step.train_step(G, D, og, od, lb, vb, cfg, rng="device", grad_sync=grad_sync, sync_losses="step", overlap=OVERLAP)
        result_sink.append((d_losses, g_loss))  # the step's 6 losses (trainer.py:479,493) read back as floats, one D2H

    # untimed warm-up: the W steps asked for; with CUDA-graph replay at least one pass over every distinct batch shape plus
    # the eager first step, so that the graph memory pools reach their steady-state size before the clock starts (a first
    # capture at a new, larger N grows the pool with cudaMalloc: 30-700 ms once per pool, profiles/tools/graph_steps.py)
    # (round 2: TWO passes - the first timed block of a full run once came out 13 % slower than the blocks after it)
    # and at least as many steps as the timed block itself, so that clocks / power state / pools are where the timed block keeps them)
    n_warm = max(args.warmup, 2 * NUM_BATCHES + 2, args.steps) if gstep is not None else args.warmup
    for i in range(n_warm):
        step_resident(i)
    clocks = _Clocks(local_rank)
    clocks.start()
    launches0 = lib.LAUNCHES
    ms = timed(step_resident, args.steps, start=n_warm)
    launches = lib.LAUNCHES - launches0
    clock_info = clocks.stop()
    # the same step with the trainer's per-step metrics call (trainer.py:497 -> step.compute_metrics: one
    # confusion-matrix kernel + macro scores on the device, no sklearn, no D2H) included
    metric_sink = []

    def step_with_metrics(i):
        out = step_resident(i)
        metric_sink.append(step.compute_metrics(resident[i % len(resident)][1], out[2], cfg))
        if len(metric_sink) > 4:
            metric_sink.pop(0)

    step_with_metrics(0)
    ms_metrics = timed(step_with_metrics, args.steps, start=1)
    n_e2e_warm = min(args.warmup, 2)
    for i in range(n_e2e_warm):
        step_e2e(i)
    drain(0)
    n_before = len(result_sink)
    ms_e2e = timed(step_e2e, args.steps, fin=lambda: drain(0), start=n_e2e_warm)
    assert len(result_sink) - n_before == args.steps, "every timed end-to-end step must have delivered its losses to the host"

    # ---- roofline of the aggregation kernel at the bench shapes (rank 0): the 46 gat_fwd launches of one step's layer
    # widths (6 generator passes x 14 + 16 discriminator passes x 6 widths), CUDA-graph replayed so the CPU launch cost
    # is out of the picture, timed with CUDA events on the launching stream, L2 flushed before every replay.
    roofline, ops, roof_step, agg = None, None, None, None
    if rank == 0:
        roof_step = _gat_roofline(resident[0][1], G, D, flush)
        if not args.no_hbm_roofline:
            roofline, agg = _hbm_roofline(dev, flush)
        if roofline is None:
            roofline = roof_step
    # one extra step under the CUPTI profiler (kernel shares).  With several ranks the step contains collectives (the fused
    # peer-memory exchange sits inside optimizer.step()), so EVERY rank runs it; rank 0 reports.
    if world > 1 or rank == 0:
        run_one = (lambda: gstep(*resident[0], sync_losses=False)) if gstep is not None else \
            (lambda: step.train_step(G, D, og, od, *resident[0], cfg, rng="device", grad_sync=grad_sync, sync_losses=False, overlap=OVERLAP))
        ops = _kernel_shares(run_one)
    cpu_baseline, torch_gpu, dropin, parity, extra = None, None, None, None, {}
    if rank == 0:
        roof_dense = _dense_rooflines(resident[0][1], G, D, flush)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dropin = _dropin_block(cfg, host, dev, args.steps, flush)
        parity = _parity_block(G, D, resident[0], cfg, lib)
        cpu_baseline = _cpu_baseline(host)
        torch_gpu = _torch_gpu_baseline(host, dev)
        if not args.no_extra_workloads:
            extra = _extra_workloads(cfg, dev, flush)

    if rank == 0:
        val = world * args.steps / (ms * 1e-3)
        e2e_val = world * args.steps / (ms_e2e * 1e-3)
        torch_val = torch_gpu.get("value") if isinstance(torch_gpu, dict) else None
        line = {"metric": "G+D train steps/sec", "value": round(val, 3), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "global_batch": BATCH * world, "batches_cycled": NUM_BATCHES, "warmup_requested": args.warmup,
                           "semantics": ("single process" if world == 1 else
                                         f"DistributedDataParallel-equivalent: {world} ranks x batch {BATCH}, rank-local GraphNorm / type-table "
                                         f"statistics and batch-mean losses, gradients averaged (NOT the same function as one GPU at batch "
                                         f"{BATCH * world}: GraphNorm batch=None couples all nodes of a batch, models.py:73,83,193,203)"),
                           "l2": f"flushed between timed iterations ({L2_FLUSH_BYTES >> 20} MiB write)",
                           "rng": "z / GP mix drawn on device; dropout masks + Gumbel noise from in-kernel Philox (BG_RNG=philox)",
                           "adam": ("building_gan_b200.optim.Adam (torch.optim.Adam semantics, one launch over flat buffers)"
                                    if args.adam == "flat" else "torch.optim.Adam (foreach)"),
                           "overlap": ("independent passes on 4 streams (step.Lanes): sampling passes up front, D(real)/D(fake) beside "
                                       "the gradient-penalty pass" if OVERLAP else "single stream"),
                           "cuda_graphs": ("critic update and sampling pass captured once per step, replayed N_CRITIC times "
                                           "(graphs.GraphedStep); the next step's graphs are captured during the current step "
                                           "(next_batch=: prefetch depth 1; e2e: the next batch's H2D copy is enqueued one step ahead, "
                                           "inside the timed region)" if gstep is not None else "none"),
                           "gc": "cyclic garbage collected and the heap frozen (gc.freeze) before each timed block, outside it; the collector stays on",
                           "pdl": _pdl_state(lib),
                           "grads": os.environ.get("BG_GRADS", "bucket"), "executor": os.environ.get("BG_EXECUTOR", "native"),
                           "gradient_exchange": sync_kind,
                           "sharding": ("global batch of 32 x world buildings per step, contiguous shards balanced by voxel count "
                                        "(dist.shard_by_nodes)" if world > 1 else "none"),
                           "parallelism": f"dp{world}" if world > 1 else "single"},
                "e2e": {"value": round(world * args.steps / (ms_e2e * 1e-3), 3), "unit": "steps/s",
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 * (cfg.N_CRITIC + 1),
                        "ms_per_step": round(ms_e2e / args.steps, 3)},
                "with_metrics": {"value": round(world * args.steps / (ms_metrics * 1e-3), 3), "unit": "steps/s",
                                 "ms_per_step": round(ms_metrics / args.steps, 3),
                                 "note": "resident batches, trainer.py:497 metrics call included (device-side macro F1 / precision / "
                                         "recall / accuracy, step.compute_metrics)"},
                "gpu_launches": (ops["libbgb200_launches_per_step"] * args.steps
                                 if isinstance(ops, dict) and "libbgb200_launches_per_step" in ops else launches),
                "gpu_launches_source": "kernels of libbgb200.so counted by CUPTI over one step x steps (torch's own kernels "
                                       "excluded); fallback: the Python layer's per-pass estimate",
                "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "parity": parity,
                # the three rates the north star names, each with its configuration spelled out
                "dropin": dropin,
                "fast_path": {"value": round(e2e_val, 3), "unit": "steps/s",
                              "config": "graphs.GraphedStep + optim.Adam (one-launch Adam) + step.Lanes + device RNG: the `e2e` number of this line"},
                "torch_b200_baseline": torch_gpu,
                "x_over_torch_b200": ({"fast_path": round(e2e_val / torch_val, 2),
                                       "dropin": (round(dropin["value"] / torch_val, 2) if isinstance(dropin, dict) and dropin.get("value") else None),
                                       "target": 20.0, "denominator": "torch_b200_baseline.value (the reference algorithm as plain torch ops on this B200)"}
                                      if torch_val else None),
                "roofline_step_shapes": roof_step, "roofline_dense": roof_dense if rank == 0 else None,
                "aggregation_hbm": agg, "kernels": ops, **extra}
        print(json.dumps(line), flush=True)


def _gat_roofline(vb, G, D, flush):
    from building_gan_b200 import lib
    csr = vb.bg_csr
    n, e, dev = csr.num_nodes, csr.num_edges, vb.x.device
    widths = [c.cout for c in G._convs] * 6 + [c.cout for c in D._convs] * 16
    bufs = {}
    for c in set(widths):
        bufs[c] = (torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev), torch.zeros(c, device=dev))
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for c in set(widths):
            lib.gat_fwd(csr, *bufs[c])
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            for c in widths:
                lib.gat_fwd(csr, *bufs[c])
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    total_bytes = sum(_gat_fwd_bytes(n, e, c) for c in widths)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    peak, src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    ach = total_bytes / (ms * 1e-3) / 1e9
    return {"kernel": "gat_fwd_kernel<C>: GATConv edge-softmax + aggregation, the step's 180 launches (6 G passes x 14 + 16 D "
                      "passes x 6 layer widths)", "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "peak_source": src,
            "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None, "avg_launch_us": round(1e3 * ms / len(widths), 3),
            "algorithmic_bytes_per_launch": round(total_bytes / len(widths)),
            "note": f"N={n}, E'={e}: the batch-32 working set (<= 8 MB per layer) is L2-resident and every launch is "
                    "latency-bound, so this fraction is NOT an HBM number; cold-L2 HBM-sized runs (N=1e5/1e6): "
                    "`bench.py --workload c4`, results in profiles/"}


def _hbm_roofline(dev, flush):
    """BASELINE config 4 / SURVEY section 8(d) cold-cache rule: the aggregation kernels on a 10-graph batch of 1e5-voxel
    irregular grids (N = 1e6, compulsory bytes >= 4x L2), L2 flushed before every launch, CUDA events on the launching
    stream, median of 7.  Returns (roofline object for gat_fwd at C=64, per-kernel table for C in 16/64/128)."""
    from building_gan_b200 import graph, lib, synth
    from building_gan_b200.benchmarks import _peak, _time
    peak, src = _peak()
    pairs = [synth.large_grid_pair(900 + i) for i in range(10)]
    _, vb = graph.collate_fn(pairs)
    csr = vb.bg_csr.to(dev)
    n, e = csr.num_nodes, csr.num_edges
    table, roof = {}, None
    for c in (16, 64, 128):
        h, s, d = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
        b, a1, a2 = torch.zeros(c, device=dev), torch.randn(c, device=dev), torch.randn(c, device=dev)
        g = torch.randn(n, c, device=dev)
        one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        o, m, z = lib.gat_fwd(csr, h, s, d, b)
        x1, stats = lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2)
        for _ in range(2):
            lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2)
            lib.graphnorm_bwd(g, o, x1, one, one, stats, 1.25)
        cases = (("gat_fwd", lambda: lib.gat_fwd(csr, h, s, d, b), _gat_fwd_bytes(n, e, c)),
                 ("gat_bwd", lambda: lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2), 4 * (4 * n * c + 8 * n + 3 * e + c + 2)),
                 ("graphnorm_fwd", lambda: lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2), 4 * (2 * n * c + 4 * c)),
                 # what the model passes run: statistics fused into the aggregation epilogue + one elementwise pass
                 ("gat_fwd+graphnorm_fwd_fused", lambda: lib.gat_fwd_gn(csr, h, s, d, b, one, zero, one, None, 0.8, 1, 2),
                  _gat_fwd_bytes(n, e, c) + 4 * (2 * n * c + 4 * c)),
                 ("graphnorm_bwd", lambda: lib.graphnorm_bwd(g, o, x1, one, one, stats, 1.25), 4 * (3 * n * c + 6 * c)),
                 # what the model passes run: GraphNorm backward moments + its elementwise half fused into the aggregation backward
                 ("graphnorm_bwd+gat_bwd_fused", lambda: lib.gat_bwd_gn(csr, g, o, x1, one, one, stats, 1.25, h, s, d, m, z, a1, a2),
                  4 * (3 * n * c + 6 * c) + 4 * (4 * n * c + 8 * n + 3 * e + c + 2)))
        row = {}
        for name, fn, by in cases:
            t = _time(fn, flush)
            row[name] = {"us": round(t * 1e6, 1), "GBs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)}
            if name == "gat_fwd" and c == 64:
                roof = {"kernel": "gat_fwd_kernel<64, pipelined>: GATConv edge-softmax + aggregation (BASELINE config 4: 10 x 1e5-voxel "
                                  "irregular 6-neighbour grids, N=1e6, E'=6.76e6, C=64)", "bound": "hbm",
                        "achieved": round(by / t / 1e9, 1), "peak": peak, "peak_source": src + " (MEASURED_PEAKS.json hbm_gbs)",
                        "unit": "GB/s", "frac": round(by / t / 1e9 / peak, 4), "traffic": _ncu_traffic("gat_fwd_kernel"),
                        "avg_launch_us": round(t * 1e6, 1), "algorithmic_bytes_per_launch": by,
                        "timing": "CUDA events around one launch, L2 flushed (256 MiB write) before each, median of 7",
                        "note": "the training step's own batch-32 shapes are L2-resident and latency-bound: see roofline_step_shapes"}
        table[f"C={c}"] = row
        del h, g, o, x1
    return roof, {"N": n, "E": e, "cold_l2": True, "peak_gbs": peak, "kernels": table}


def _ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` summary."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full_gat_c64_N1e6.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            col = next(i for i, h in enumerate(rows[0]) if kernel in h)
            vals = {r[0]: r for r in rows[1:]}
            mb = float(vals["dram__bytes_read.sum"][col]) + float(vals["dram__bytes_write.sum"][col])
            return {"bytes": int(mb * 1e6), "source": os.path.basename(path)}
        except Exception:
            continue
    return None


def _dropin_block(cfg, host, dev, steps, flush):
    """What "drops into trainer.py unchanged" costs on the kernels: the reference loop body (trainer.py:459-503) restated
    in trainer_helper.ReferenceTrainerHelper.train_batch - z and the GP mixing factor drawn on the CPU generator and copied
    (:298,470,484), ``.item()`` after every backward (:479,493), torch.optim.Adam (train.py:36-37), the per-building Python FAR
    loop (:362-380), sklearn metrics with their D2H copies (:497) - on fresh drop-in models, one stream, no CUDA graphs.
    Batches start in pinned host memory and are moved with ``.to(DEVICE)`` inside the timed region like trainer.py:461-462."""
    try:
        from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
        from building_gan_b200.trainer_helper import ReferenceTrainerHelper, TrainerHelper
        out = {}
        n = max(4, min(steps, 10))
        for name, cls, metrics in (("reference_loop", ReferenceTrainerHelper, True), ("reference_loop_no_metrics", ReferenceTrainerHelper, False),
                                   ("mixin", TrainerHelper, True)):
            torch.manual_seed(777)
            h = cls()
            h.configuration = cfg
            h.generator, h.discriminator = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
            h.optimizer_generator = torch.optim.Adam(h.generator.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
            h.optimizer_discriminator = torch.optim.Adam(h.discriminator.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
            for i in range(3):
                h.train_batch(*_clone_to(*host[i % len(host)], "cpu"), with_metrics=metrics)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n):
                flush.zero_()
                h.train_batch(*_clone_to(*host[(3 + i) % len(host)], "cpu"), with_metrics=metrics)
            torch.cuda.synchronize()
            out[name] = round(n / (time.perf_counter() - t0), 3)
        return {"value": out["reference_loop"], "unit": "steps/s", "steps": n,
                "config": "unchanged-trainer semantics: CPU-drawn z / GP mix + H2D, .item() after each of the 6 backwards, torch.optim.Adam "
                          "(foreach), per-building Python FAR loop, sklearn metrics call included (trainer.py:459-503), batch moved "
                          "from pinned host memory inside the timed region, single stream, no CUDA graphs, BG_RNG=philox dropout",
                "without_metrics_call": out["reference_loop_no_metrics"],
                "with_trainer_helper_mixin": {"value": out["mixin"], "config": "same loop with trainer_helper.TrainerHelper mixed in: FAR "
                                              "loop -> one segment sum, sklearn -> one confusion-matrix kernel + one D2H, fused critic-loss glue"},
                "timing": "host wall clock around the loop with a device synchronize on both sides (the loop itself syncs 6+ times per step)"}
    except Exception as exc:
        return {"error": repr(exc)[:300]}


def _parity_block(G, D, resident0, cfg, lib):
    """Parity of the TIMED configuration: the bench's own models (weights after the timed steps), one bench batch (N ~ 15 k),
    the dense mode the bench ran, eval-mode dropout, against the fp64 oracle (oracle/check.py - the checker, not the product)."""
    try:
        from building_gan_b200 import step
        from oracle import check
        lb, vb = resident0
        r = check.parity_report(G, D, lb, vb, cfg, step, gradients=True, envelope=True)
        mode = lib.set_dense_tc(1)
        lib.set_dense_tc(mode)
        r["dense_mode"] = {0: "FFMA", 1: "tcgen05 3xTF32, 4 TMEM accumulators (128-wide layers) + FFMA", 2: "tcgen05 bf16 operands (128-wide layers) + FFMA"}[mode]
        r["norm"] = "max|a - ref| / max|ref| per tensor; gradients: worst parameter tensor"
        return r
    except Exception as exc:
        return {"error": repr(exc)[:300]}


def _dense_rooflines(vb, G, D, flush):
    """Rooflines of the dense work that dominates the step's GPU time (launch list profiles/r02b_*: dense_mma_kernel 23 %,
    wgrad_mma_kernel 13.5 %, dense_tc_kernel 8 %): the step's own dense shapes through the public entry points (bg_dense_fwd /
    bg_dense_wgrad dispatch them exactly as inside the step), CUDA-graph replayed, cold L2, CUDA events; logical flops =
    2 N K Cout against the FP32 FFMA peak (148 SMs x 128 lanes x 2 x 1.965 GHz) - the warp-MMA kernels issue 3 TF32 MMAs per
    logical product - and algorithmic bytes against the measured HBM peak.  At N ~ 15 k all of them are latency-bound: the
    honest figure is the average launch time against a ~1 us launch floor and ~3 us for an elementwise pass over the same rows."""
    try:
        from building_gan_b200 import lib
        n, dev = vb.num_nodes, vb.x.device
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        shapes = [(c.cin, c.cout) for c in D._convs] + [(36, 64), (64, 64), (64, 32), (32, 16), (16, 8), (8, 1)]
        shapes += [(c.cin, c.cout) for c in G._convs] + [(64, 32), (32, 16), (16, 7)]
        out = {}
        for label, mode in (("bg_dense_fwd (dense_mma_kernel: mma.sync 3xTF32 for K % 8 == 0 and Cout in 8..64; rowdense / tiled FFMA for the rest)", "fwd"),
                            ("bg_dense_wgrad (wgrad_mma_kernel: mma.sync 3xTF32 partial sums + fold)", "wgrad")):
            bufs = []
            for k, c in shapes:
                x, w, g = torch.randn(n, k, device=dev), torch.randn(c, k, device=dev), torch.randn(n, c, device=dev)
                bufs.append((x, w, g))
            run = (lambda x, w, g: lib.dense_fwd([x], w)) if mode == "fwd" else (lambda x, w, g: lib.dense_wgrad(g, [x]))
            stream = torch.cuda.Stream()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(stream):
                for b in bufs:
                    run(*b)
                torch.cuda.synchronize()
                with torch.cuda.graph(graph, stream=stream):
                    for b in bufs:
                        run(*b)
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            flops = sum(2.0 * n * k * c for k, c in shapes)
            byts = sum(4.0 * (n * k + n * c + k * c) for k, c in shapes)
            launches = len(shapes) * (1 if mode == "fwd" else 2)
            out[label] = {"launches": launches, "avg_launch_us": round(1e3 * ms / launches, 2),
                          "tflops": round(flops / (ms * 1e-3) / 1e12, 3), "fp32_ffma_peak_tflops": round(FP32_FFMA_PEAK_TFLOPS, 1),
                          "frac_fp32_peak": round(flops / (ms * 1e-3) / 1e12 / FP32_FFMA_PEAK_TFLOPS, 4),
                          "algorithmic_GBs": round(byts / (ms * 1e-3) / 1e9, 1), "frac_hbm_peak": round(byts / (ms * 1e-3) / 1e9 / hbm, 4),
                          "bound": "latency (L2-resident operands, tens of CTAs per launch)"}
        out["shapes"] = f"N={n}; (K, Cout) of the narrow / plain dense layers of one G + one D pass: {shapes}"
        # the 128-wide Linear + LayerNorm + LeakyReLU layers of the generator: tcgen05 3xTF32 (dense_tc_kernel<128>)
        tensor_peak = peaks.get("bf16_tflops", peaks.get("dense_bf16_tflops", 0.0)) or None
        x, w, b = torch.randn(n, 128, device=dev), torch.randn(128, 128, device=dev) * 0.1, torch.zeros(128, device=dev)
        gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
        run = lambda: lib.dense_fwd([x], w, b, (gam, bet), 2, save_ln=True)
        reps = 4
        stream, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            run()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(reps):
                    run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2] / reps
        tf = 2.0 * n * 128 * 128 / (ms * 1e-3) / 1e12
        out["dense_tc_kernel<128> (tcgen05 3xTF32, Linear 128 -> 128 + LayerNorm + LeakyReLU, xhat / rstd saved)"] = {
            "launches": reps, "avg_launch_us": round(1e3 * ms, 2), "logical_tflops": round(tf, 2), "tensor_pipe_tflops": round(3 * tf, 2),
            "measured_bf16_peak_tflops": tensor_peak,
            "frac_tensor_peak": (round(3 * tf / (tensor_peak / 2), 4) if tensor_peak else None),
            "algorithmic_GBs": round(4.0 * (3 * n * 128 + 128 * 128) / (ms * 1e-3) / 1e9, 1),
            "bound": "latency (119 tiles on 148 SMs: one wave; TF32 peak taken as half the measured bf16 peak; 4 launches back to back in one CUDA graph)"}
        return out
    except Exception as exc:
        return {"error": repr(exc)[:300]}


def _extra_workloads(cfg, dev, flush):
    """BASELINE metric 2 (generated buildings/s, config 3) and config 1 (sanity single-datum overfit step, N ~ 400, where launch
    latency is everything) as extra keys of the default line, so the driver's record carries them.  Bounded: a few seconds."""
    out = {}
    try:
        from building_gan_b200 import graph, step, synth
        from building_gan_b200.graphs import GraphedStep
        from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
        from building_gan_b200.optim import Adam
        torch.manual_seed(777)
        G, D = VoxelGNNGenerator(cfg, 17, 12).to(dev), VoxelGNNDiscriminator(cfg, 17, 12).to(dev)
        # ---- config 3: generator-only sampling, batch 512 (the reference's BATCH_SIZE, config.py:63, used by Trainer.test)
        G.eval()
        lb, vb = graph.collate_fn([synth.building_pair_fast(4001 + i) for i in range(512)])
        lb, vb = lb.to(dev), vb.to(dev)
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        cur = torch.cuda.current_stream()

        def run(k):
            for s_ in streams:
                s_.wait_stream(cur)
            for i in range(k):
                with torch.cuda.stream(streams[i % 2]):
                    lab = step.sample(G, lb, vb, cfg)
            for s_ in streams:
                cur.wait_stream(s_)
            return lab
        run(4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 12
        e0.record()
        lab = run(k)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        out["sampling"] = {"metric": "generated buildings/sec", "value": round(512 * k / sec, 1), "unit": "buildings/s",
                           "config": {"workload": "BASELINE config 3: generator-only sampling (eval forward + argmax, trainer.py:769-770,:73), "
                                                  "batch 512 buildings", "voxels_per_batch": int(vb.num_nodes), "streams": 2, "passes": k},
                           "ms_per_pass": round(1e3 * sec / k, 3), "labels_checksum": int(lab.sum())}
        # the same pass through the oracle on the host cores (bounded: one pass after one warm-up)
        try:
            from oracle import check, models as omodels
            oG = omodels.OracleGenerator(cfg, 17, 12).eval()
            oG.load_state_dict({k_: v.detach().cpu() for k_, v in G.state_dict().items()})
            olb, ovb = check.to_oracle_batch(lb, torch.float32), check.to_oracle_batch(vb, torch.float32)
            torch.set_num_threads(os.cpu_count() or 1)
            with torch.no_grad():
                z = torch.randn(1, ovb.num_nodes, cfg.Z_DIM)
                oG(olb, ovb, z)
                t0 = time.perf_counter()
                oG(olb, ovb, z)[1].argmax(1)
                dt = time.perf_counter() - t0
            out["sampling"]["cpu_baseline"] = {"value": round(512 / dt, 1), "unit": "buildings/s", "cores": os.cpu_count(), "kind": "port",
                                               "sample": "one batch-512 eval forward + argmax of the oracle generator"}
        except Exception as exc:
            out["sampling"]["cpu_baseline"] = {"error": repr(exc)[:200]}
        del lb, vb
        # ---- config 1: sanity.py single-datum overfit (one building, trainer.py:459-495 in a loop)
        G.train()
        og = Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
        od = Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
        lb1, vb1 = graph.collate_fn([synth.building_pair_fast(4001)])
        lb1, vb1 = lb1.to(dev), vb1.to(dev)
        gs = GraphedStep(G, D, og, od, cfg)
        for _ in range(8):  # one eager step, then the graph pools' steady state (see the conv-type block below)
            gs(lb1, vb1, sync_losses=False)
        _settle_heap()
        torch.cuda.synchronize()
        k = 20
        e0.record()
        for _ in range(k):
            gs(lb1, vb1, sync_losses=False)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        out["config1_sanity"] = {"metric": "G+D train steps/sec", "value": round(k / sec, 2), "unit": "steps/s",
                                 "ms_per_step": round(1e3 * sec / k, 3),
                                 "config": {"workload": "BASELINE config 1: sanity.py single-datum overfit step (one building)",
                                            "voxels": int(vb1.num_nodes), "path": "graphs.GraphedStep"}}
        # ---- the three non-default conv types (models.py:22-31) through the same graphed step, batch 32
        from building_gan_b200 import Configuration
        lb32, vb32 = graph.collate_fn([synth.building_pair_fast(i) for i in _building_ids(0, 0, BATCH)])
        lb32, vb32 = lb32.to(dev), vb32.to(dev)
        conv = {}
        for kind in ("GCNCONV", "GRAPHCONV", "GATV2CONV"):
            c2 = Configuration()
            c2.GENERATOR_CONV_TYPE = c2.DISCRIMINATOR_CONV_TYPE = kind
            torch.manual_seed(777)
            G2, D2 = VoxelGNNGenerator(c2, 17, 12).to(dev), VoxelGNNDiscriminator(c2, 17, 12).to(dev)
            og2 = Adam(G2.parameters(), lr=c2.LEARNING_RATE_GENERATOR, betas=c2.BETAS)
            od2 = Adam(D2.parameters(), lr=c2.LEARNING_RATE_DISCRIMINATOR, betas=c2.BETAS)
            gs2 = GraphedStep(G2, D2, og2, od2, c2)
            # one eager step + six graphed ones: the graphs of the last two steps stay alive (graphs.GraphedStep._alive), so the two
            # private graph pools reach their steady state only after ~4 graphed steps - with fewer the timed steps still
            # meet cudaMalloc (readings of 3.8 .. 18 steps/s for the same model on different boxes)
            for _ in range(7):
                gs2(lb32, vb32, sync_losses=False)
            _settle_heap()
            torch.cuda.synchronize()
            k = 8
            e0.record()
            for _ in range(k):
                gs2(lb32, vb32, sync_losses=False)
            e1.record()
            torch.cuda.synchronize()
            conv[kind] = round(k / (e0.elapsed_time(e1) * 1e-3), 2)
            del G2, D2, og2, od2, gs2
        out["conv_types"] = {"metric": "G+D train steps/sec", "unit": "steps/s", "values": conv,
                             "config": "batch 32, graphs.GraphedStep, GENERATOR_CONV_TYPE = DISCRIMINATOR_CONV_TYPE = the key "
                                       "(op-by-op executor on the library kernels inside the captured graphs)"}
    except Exception as exc:
        out["extra_workloads_error"] = repr(exc)[:300]
    return out


def _torch_gpu_baseline(host_batches, dev):
    """The reference algorithm (oracle restatement: plain torch ops, scatter_add / index_select, autograd double backward)
    on the SAME B200 - the stand-in for "the reference's single-B200 PyTorch path" of the north star (torch_geometric
    itself is not installable).  TF32 off.  Bounded: 1 warm-up + 3 steps."""
    try:
        from building_gan_b200 import Configuration
        from oracle import models as omodels, pyg as opyg, trainer as otrainer
        torch.backends.cuda.matmul.allow_tf32 = False
        cfg = Configuration()
        torch.manual_seed(777)
        G, D = omodels.OracleGenerator(cfg, 17, 12).to(dev), omodels.OracleDiscriminator(cfg, 17, 12).to(dev)
        og = torch.optim.Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
        od = torch.optim.Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
        batches = [(_to_oracle(lb, opyg).to(dev), _to_oracle(vb, opyg).to(dev)) for lb, vb in host_batches[:2]]
        otrainer.train_step(G, D, og, od, *batches[0], cfg)
        torch.cuda.synchronize()
        n = 3
        t0 = time.perf_counter()
        for i in range(n):
            otrainer.train_step(G, D, og, od, *batches[(i + 1) % 2], cfg)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": round(n / dt, 4), "unit": "steps/s", "kind": "port on cuda:0 (plain torch ops, fp32, TF32 off)",
                "sample": f"{n} full batch-32 steps (after 1 warm-up) of the oracle restatement of trainer.py:467-495"}
    except Exception as exc:  # informational
        return {"error": repr(exc)[:200]}


def _kernel_shares(run_step):
    """GPU time per kernel of one step (torch.profiler / CUPTI) - informational, not used for `value`."""
    try:
        import collections
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run_step()
            torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0, 0.0])
        for ev in prof.events():
            if str(ev.device_type).endswith("CUDA") and ev.device_time > 0:
                name = ev.name.split("(")[0].split("<")[0].replace("void ", "")
                agg[name][0] += 1
                agg[name][1] += ev.device_time
        tot = sum(v[1] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]
        return {"gpu_ms_per_step": round(tot / 1e3, 3), "launches_per_step": sum(v[0] for v in agg.values()),
                "libbgb200_launches_per_step": sum(v[0] for k, v in agg.items() if "bg::" in k),
                "top": {k: {"launches": c, "ms": round(t / 1e3, 3), "share": round(t / tot, 3)} for k, (c, t) in top}}
    except Exception as exc:  # profiling is best-effort
        return {"error": repr(exc)}


def _cpu_baseline(host_batches):
    """The oracle (reference algorithm) on the host cores, bounded sample: 1 warm-up + 3 timed steps."""
    from building_gan_b200 import Configuration
    from oracle import models as omodels, pyg as opyg, trainer as otrainer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = Configuration()
    torch.manual_seed(777)
    G, D = omodels.OracleGenerator(cfg, 17, 12), omodels.OracleDiscriminator(cfg, 17, 12)
    og = torch.optim.Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
    od = torch.optim.Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
    batches = [(_to_oracle(lb, opyg), _to_oracle(vb, opyg)) for lb, vb in host_batches[:2]]
    otrainer.train_step(G, D, og, od, *batches[0], cfg)
    n = 3
    t0 = time.perf_counter()
    for i in range(n):
        otrainer.train_step(G, D, og, od, *batches[(i + 1) % 2], cfg)
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 4), "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} full batch-32 steps (after 1 warm-up) of the oracle restatement of trainer.py:467-495"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "sample", "c4"])
    ap.add_argument("--adam", default="flat", choices=["flat", "torch"])
    ap.add_argument("--no-graph", action="store_true", help="no CUDA-graph replay of the critic updates")
    ap.add_argument("--nccl-sync", action="store_true", help="gradient exchange through ncclAllReduce instead of the fused peer-memory kernel")
    ap.add_argument("--no-overlap", action="store_true", help="every pass on one stream in the reference's call order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-roofline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true", help="skip the sampling (config 3) / sanity (config 1) extra keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = _env_int("RANK", 0), _env_int("LOCAL_RANK", 0), _env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload != "train":
        from building_gan_b200 import benchmarks
        benchmarks.run(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
