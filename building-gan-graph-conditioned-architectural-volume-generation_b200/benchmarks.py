"""Extra benchmark workloads of bench.py (not the driver's headline line).

``c4``   BASELINE config 4: aggregation / normalisation kernels on large irregular voxel grids (one 1e5-voxel graph and
         a 10-graph 1e6-voxel batch), forward + backward (+ second-order), cold L2 (256 MiB flush before every launch),
         CUDA-event timed per launch; reports algorithmic GB/s (SURVEY section 8d byte counts) against the measured HBM peak.
``sample`` BASELINE config 3: generator-only sampling (eval forward + argmax), buildings/s.
"""
from __future__ import annotations

import json
import os
import time

import torch

from . import Configuration, graph, lib, step, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def _time(fn, flush, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e-3


def run_c4(args, dev):
    peak, src = _peak()
    flush = torch.empty((256 << 20) // 4, dtype=torch.float32, device=dev)
    out = {"workload": "c4", "peak_gbs": peak, "peak_source": src, "cases": []}
    for ngraphs in (1, 10):
        pairs = [synth.large_grid_pair(900 + i) for i in range(ngraphs)]
        _, vb = graph.collate_fn(pairs)
        csr = vb.bg_csr.to(dev)
        n, e = csr.num_nodes, csr.num_edges
        for c in (1, 2, 4, 8, 16, 32, 64, 128):
            h, s, d = torch.randn(n, c, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
            b, a1, a2 = torch.zeros(c, device=dev), torch.randn(c, device=dev), torch.randn(c, device=dev)
            g = torch.randn(n, c, device=dev)
            one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            o, m, z = lib.gat_fwd(csr, h, s, d, b)
            x1, stats = lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2)
            for _ in range(2):  # warm-up (first launches, workspace growth)
                lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2)
                lib.graphnorm_bwd(g, o, x1, one, one, stats, 1.25)
            t_f = _time(lambda: lib.gat_fwd(csr, h, s, d, b), flush)
            t_b = _time(lambda: lib.gat_bwd(csr, g, h, s, d, m, z, a1, a2), flush)
            t_nf = _time(lambda: lib.graphnorm_fwd(o, one, zero, one, None, 0.8, 1, 2), flush)
            t_nb = _time(lambda: lib.graphnorm_bwd(g, o, x1, one, one, stats, 1.25), flush)
            by_f = 4 * (2 * n * c + 5 * n + e + c + 1)
            by_b = 4 * (4 * n * c + 8 * n + 3 * e + c + 2)
            by_nf = 4 * (2 * n * c + 4 * c)
            by_nb = 4 * (3 * n * c + 6 * c)
            case = {"N": n, "E": e, "C": c}
            for name, t, by in (("gat_fwd", t_f, by_f), ("gat_bwd", t_b, by_b), ("graphnorm_fwd", t_nf, by_nf),
                                ("graphnorm_bwd", t_nb, by_nb)):
                case[name] = {"us": round(t * 1e6, 2), "alg_MB": round(by / 1e6, 2), "GBs": round(by / t / 1e9, 1),
                              "frac": round(by / t / 1e9 / peak, 3)}
            out["cases"].append(case)
            del h, g, o, x1
    print(json.dumps(out), flush=True)


def run_sample(args, dev, rank=0, world=1):
    """BASELINE config 3 / metric 2: generator-only sampling (eval forward + argmax, trainer.py:769-770 + :73), batch 512
    buildings per GPU (the reference's BATCH_SIZE, config.py:63, used by Trainer.test).  Independent batches: every rank samples
    its own buildings (no collective on the data path; weak scaling), and on one GPU consecutive batches alternate between
    ``BG_SAMPLE_STREAMS`` streams (default 2) so that the latency-bound narrow layers of one pass overlap the wide layers of
    the other.  value = buildings of all ranks / max-over-ranks device time."""
    import os
    import torch.distributed as dist
    from .models import VoxelGNNGenerator
    cfg = Configuration()
    torch.manual_seed(777)
    G = VoxelGNNGenerator(cfg, 17, 12).to(dev).eval()
    batch = 512
    pairs = [synth.building_pair_fast(4001 + rank * batch + i) for i in range(batch)]
    lb, vb = graph.collate_fn(pairs)
    lb, vb = lb.to(dev), vb.to(dev)
    ns = max(1, int(os.environ.get("BG_SAMPLE_STREAMS", "2")))
    streams = [torch.cuda.Stream(device=dev) for _ in range(ns)] if ns > 1 else [torch.cuda.current_stream()]
    cur = torch.cuda.current_stream()
    labels = [None] * len(streams)

    def run_steps(k):
        for s in streams:
            s.wait_stream(cur)
        for i in range(k):
            with torch.cuda.stream(streams[i % len(streams)]):
                labels[i % len(streams)] = step.sample(G, lb, vb, cfg)
        for s in streams:
            cur.wait_stream(s)

    run_steps(max(args.warmup, 3) * len(streams))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    voxels = vb.num_nodes
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return
    sec = ms * 1e-3
    print(json.dumps({"metric": "generated buildings/sec", "value": round(world * batch * args.steps / sec, 1), "unit": "buildings/s",
                      "n_gpus": world, "steps": args.steps, "ms_per_step": round(1e3 * sec / args.steps, 3), "scaling": "weak",
                      "config": {"workload": "generator-only sampling (eval forward + argmax), batch 512 buildings per GPU",
                                 "voxels_per_batch": voxels, "streams": len(streams)},
                      "labels_checksum": int(labels[0].sum())}), flush=True)


def run(args, rank, local_rank, world):
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib.load()
    if args.workload == "c4":
        if rank == 0:
            run_c4(args, dev)  # a single-GPU kernel sweep
    else:
        run_sample(args, dev, rank, world)
