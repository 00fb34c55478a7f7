"""Parity AT THE BENCHMARKED SIZES AND MODES (round-1 review: every earlier GPU test ran 2-4 buildings or forced the
large-graph launch geometry onto small graphs with ``bg_tune``).

* BASELINE config 2 - one batch of 32 buildings (N ~ 15 k), the size ``bench.py`` times: generator forward, critic loss
  with its second-order backward, generator loss and every parameter gradient against the fp64 oracle, in the dense mode the
  bench runs (tcgen05) and in the strict FFMA mode.
* BASELINE config 4 - one 1e5-voxel graph and the N = 1e6 ten-graph batch, NATURAL dispatch (no ``bg_tune``): the
  aggregation kernels forward / backward, the fused statistics and fused GraphNorm-backward variants, against the fp64 oracle.
  At N = 1e6 the oracle is evaluated on a closed sub-graph (a random source sample S, all destinations T of S's out-edges,
  ALL in-edges of T): exact for ``out``/``gd`` on T and for ``gh``/``gs`` on S, at a memory cost the host can afford.
* the in-kernel Philox path (what the bench times): its dropout masks and Gumbel noise are exported through the same
  kernels, injected into the oracle and into the explicit-mask path - the latter must agree BIT FOR BIT.
"""
import pytest
import torch
import torch.nn.functional as F

from building_gan_b200 import Configuration, graph, lib, models as bm, step as bstep, synth
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from oracle import check as ocheck
from oracle import pyg
from util import assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bench_batch(batch=32, first=4001):
    pairs = [synth.building_pair_fast(first + i) for i in range(batch)]
    lb, vb = graph.collate_fn(pairs)
    return lb.to(DEV), vb.to(DEV)


def _models(seed=777):
    cfg = Configuration()
    torch.manual_seed(seed)
    G, D = VoxelGNNGenerator(cfg, 17, 12).to(DEV), VoxelGNNDiscriminator(cfg, 17, 12).to(DEV)
    with torch.no_grad():  # move norm / bias parameters off their trivial initial values
        for m in list(G.modules()) + list(D.modules()):
            for name in ("bias", "mean_scale"):
                p = getattr(m, name, None)
                if isinstance(p, torch.nn.Parameter):
                    p.add_(0.1 * torch.randn_like(p))
    return cfg, G, D


# ------------------------------------------------------------------------------------------------------------
# config 2: the batch the bench times
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dense", ["tcgen05", "ffma"])
def test_batch32_model_parity(dense):
    """Tolerances (max-abs error / max magnitude, fp64 oracle): logits / losses 2x the fp32 reference's own rounding
    envelope at this size or 1e-5, whichever is larger (FFMA), 1e-4 (tcgen05 3xTF32); parameter gradients 1e-3 (the 1- and
    2-channel bottleneck blocks' GraphNorm gradients are length-N sums with ~100x cancellation).  Labels: exact away from
    Gumbel near-ties (top-2 gap <= 1e-4)."""
    lib.set_dense_tc(dense == "tcgen05")
    try:
        cfg, G, D = _models()
        lb, vb = _bench_batch()
        r = ocheck.parity_report(G, D, lb, vb, cfg, bstep)
        print(dense, r)
        env = r["fp32_oracle_logits_rel"]
        fwd_tol = 1e-4 if dense == "tcgen05" else max(1e-5, 2 * env)
        assert r["N"] > 10_000
        assert r["logits_rel"] <= fwd_tol, r
        assert r["label_soft_rel"] <= fwd_tol, r
        assert r["labels_differ_outside_ties"] == 0, r
        assert r["d_loss_rel"] <= fwd_tol and r["g_loss_rel"] <= max(fwd_tol, 2e-4 * (r["labels_differ_at_ties"] > 0)), r
        # gradients: error of the whole update direction (l2 over all parameters) and the worst single tensor; next to the
        # tolerance the reference's OWN fp32 arithmetic is measured the same way (fp32_oracle_*): a correct fp32
        # implementation cannot be asked to sit far inside that envelope
        assert r["d_grad_rel_l2"] <= max(2e-4, 4 * r["fp32_oracle_d_grad_rel_l2"]), r
        assert r["g_grad_rel_l2"] <= max(1e-3, 4 * r["fp32_oracle_g_grad_rel_l2"]), r
        assert r["d_worst_grad_rel"] <= max(1e-3, 4 * r["fp32_oracle_d_worst_grad_rel"]), r
        assert r["g_worst_grad_rel"] <= max(2e-3, 4 * r["fp32_oracle_g_worst_grad_rel"]), r
    finally:
        lib.set_dense_tc(True)


# ------------------------------------------------------------------------------------------------------------
# config 4: aggregation kernels on large graphs, natural dispatch
# ------------------------------------------------------------------------------------------------------------
def _large(ngraphs):
    pairs = [synth.large_grid_pair(900 + i) for i in range(ngraphs)]
    _, vb = graph.collate_fn(pairs)
    return vb, pyg.gat_edges(vb.edge_index, vb.num_nodes), vb.bg_csr.to(DEV)


def _closed_subgraph(edges, n, sample, seed=0):
    """(S, T, sub-edge list): S = random sources, T = destinations of S's out-edges, sub-edges = every in-edge of T."""
    g = torch.Generator().manual_seed(seed)
    S = torch.randperm(n, generator=g)[:sample]
    in_s = torch.zeros(n, dtype=torch.bool)
    in_s[S] = True
    T = torch.unique(edges[1][in_s[edges[0]]])
    in_t = torch.zeros(n, dtype=torch.bool)
    in_t[T] = True
    return S, T, edges[:, in_t[edges[1]]]


@pytest.mark.parametrize("ngraphs,C", [(1, 16), (1, 64), (1, 128), (10, 64), (10, 16)])
def test_aggregation_kernels_large_graphs_natural_dispatch(ngraphs, C):
    vb, edges, csr = _large(ngraphs)
    n = vb.num_nodes
    gen = torch.Generator().manual_seed(C + ngraphs)
    r32 = lambda *shape: torch.randn(*shape, generator=gen)
    h, s, d, b, gout = r32(n, C), r32(n), r32(n), r32(C), r32(n, C)
    a_s, a_d = r32(C), r32(C)
    w, beta, alpha = r32(C) * 0.3 + 1, r32(C) * 0.3, r32(C) * 0.3 + 1
    f = lambda t: t.to(DEV).contiguous()
    hd, sd_, dd, bd, gd_ = f(h), f(s), f(d), f(b), f(gout)
    # kernels, natural dispatch
    o, m, z = lib.gat_fwd(csr, hd, sd_, dd, bd)
    gh_tot, gsd, _, _ = lib.gat_bwd(csr, gd_, hd, sd_, dd, m, z, f(a_s), f(a_d))
    o2, m2, z2, x1, stats = lib.gat_fwd_gn(csr, hd, sd_, dd, bd, f(w), f(beta), f(alpha))
    assert torch.equal(o, o2) and torch.equal(m, m2) and torch.equal(z, z2)
    # oracle (fp64) on the whole graph (1e5) or on a closed sub-graph (1e6)
    if ngraphs == 1:
        S, T, sub = torch.arange(n), torch.arange(n), edges
    else:
        S, T, sub = _closed_subgraph(edges, n, 20_000)
    h64, s64, d64 = h.double().requires_grad_(), s.double().requires_grad_(), d.double().requires_grad_()
    out = pyg.gat_core(h64, s64, d64, sub)
    g_h, g_s, g_d = torch.autograd.grad(out, (h64, s64, d64), gout.double())
    assert_close(o.cpu()[T], (out.detach() + b.double())[T], 1e-5, f"gat_fwd N={n} C={C}")
    ref_tot = g_h + g_s[:, None] * a_s.double() + g_d[:, None] * a_d.double()
    if ngraphs == 1:
        assert_close(gh_tot.cpu(), ref_tot, 1e-5, f"gat_bwd gh_tot N={n} C={C}")
    else:  # gh_tot of a node mixes its source role (complete on S) and its destination role (complete on T)
        both = torch.zeros(n, dtype=torch.bool)
        both[S] = True
        in_t = torch.zeros(n, dtype=torch.bool)
        in_t[T] = True
        both &= in_t
        assert int(both.sum()) > 1000
        assert_close(gh_tot.cpu()[both], ref_tot[both], 1e-5, f"gat_bwd gh_tot N={n} C={C}")
    assert_close(gsd.cpu()[S, 0], g_s[S], 1e-5, "gat_bwd gs")
    assert_close(gsd.cpu()[T, 1], g_d[T], 1e-5, "gat_bwd gd")
    # fused statistics: GraphNorm over ALL rows of the (verified) aggregate, evaluated in fp64 from the definition
    o64 = o.double()
    mu = o64.mean(0)
    oh = o64 - f(alpha).double() * mu
    var = oh.pow(2).mean(0)
    x1_ref = torch.relu(f(w).double() * oh / (var + 1e-5).sqrt() + f(beta).double())
    assert_close(stats[:C], mu, 3e-5, "fused mean")
    assert_close(stats[2 * C:], var, 3e-5, "fused var")
    assert_close(x1, x1_ref, 3e-5, "fused x1")
    del o64, oh, x1_ref
    # fused GraphNorm backward + aggregation backward == the unfused pair (same constants, same expression)
    gx1 = f(r32(n, C))
    go_ref, dpar_ref, _ = lib.graphnorm_bwd(gx1, o, x1, f(w), f(alpha), stats, 1.0)
    gh_ref, gsd_ref, _, _ = lib.gat_bwd(csr, go_ref, hd, sd_, dd, m, z, f(a_s), f(a_d))
    go, dpar, gh2, gsd2 = lib.gat_bwd_gn(csr, gx1, o, x1, f(w), f(alpha), stats, 1.0, hd, sd_, dd, m, z, f(a_s), f(a_d))
    assert torch.equal(dpar, dpar_ref)
    assert_close(go, go_ref, 1e-6, "fused go")
    assert_close(gh2, gh_ref, 1e-5, "fused gh")
    assert_close(gsd2, gsd_ref, 1e-5, "fused gsd")
    # and the GraphNorm backward itself against autograd of the definition in fp64 (per-channel moments over N rows), on the
    # ReLU pattern the kernel took (among 1e7 elements a few pre-activations sit within fp32 rounding of 0 and flip)
    o_r = o.double().requires_grad_()
    mu_r = o_r.mean(0)
    oh_r = o_r - f(alpha).double() * mu_r
    y = (f(w).double() * oh_r / (oh_r.pow(2).mean(0) + 1e-5).sqrt() + f(beta).double()) * (x1 > 0)
    (go64,) = torch.autograd.grad(y, o_r, gx1.double())
    assert_close(go_ref, go64, 2e-5, "graphnorm_bwd go")


# ------------------------------------------------------------------------------------------------------------
# the in-kernel Philox path: export its draws, replay them
# ------------------------------------------------------------------------------------------------------------
def _philox_keeps(n, widths, seed, offset):
    """The keep-masks the native passes generate for (seed, offset): GraphNorm + ReLU + dropout of an all-ones activation
    with weight 0 / bias 1 is 1/0.8 where kept and 0 where dropped (element-index keyed Philox, block k at offset + k)."""
    keeps = []
    for k, c in enumerate(widths):
        ones = torch.ones(n, c, device=DEV)
        x1, _ = lib.graphnorm_fwd(ones, torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.ones(c, device=DEV), None,
                                  0.8, seed, offset + k)
        keeps.append((x1 > 0).to(torch.uint8))
    return keeps


def _philox_gumbel(n, k, seed, offset):
    """softmax(0 + g) of the in-kernel noise; log of it is g up to a per-row constant, which the softmax ignores."""
    soft, _, _ = lib.gumbel_st_fwd(torch.zeros(n, k, device=DEV), None, seed, offset + 1000)
    return soft.log()


def test_philox_path_replays_bit_exactly_and_matches_the_oracle():
    assert bm.RNG_MODE == "philox"
    cfg, G, D = _models(5)
    lb, vb = _bench_batch(4, first=31)
    n = vb.num_nodes
    G.train(), D.train()
    z = torch.randn(1, n, cfg.Z_DIM, device=DEV)
    label = torch.softmax(torch.randn(n, cfg.NUM_CLASSES, device=DEV), 1)
    seed = torch.cuda.initial_seed() & 0xFFFFFFFFFFFFFFFF
    with torch.no_grad():
        t0 = bm._philox_calls
        logits, hard, soft = G(lb, vb, z)                       # in-kernel Philox: ticket t0 + 1
        score = D(lb, vb, label.unsqueeze(0))                   # ticket t0 + 2
    gk = _philox_keeps(n, [c.cout for c in G._convs], seed, (t0 + 1) * 2048)
    gnoise = _philox_gumbel(n, cfg.NUM_CLASSES, seed, (t0 + 1) * 2048)
    dk = _philox_keeps(n, [c.cout for c in D._convs], seed, (t0 + 2) * 2048)
    for kk in gk + dk:
        assert 0.78 < float(kk.float().mean()) < 0.82
    with torch.no_grad():                                        # the explicit-mask path on the exported draws: bit for bit
        l2, h2, s2 = G(lb, vb, z, gnoise, keeps=gk)
        sc2 = D(lb, vb, label.unsqueeze(0), keeps=dk)
    assert torch.equal(score, sc2)
    assert torch.equal(logits, l2), "philox masks differ from their export"
    assert rel_err(s2, soft) <= 2e-6 and torch.equal(h2.argmax(1), hard.argmax(1))  # noise went through log(softmax)
    # oracle in train mode with the same masks / noise
    oG, oD = ocheck.oracle_twins(G, D, cfg)
    oG.train(), oD.train()
    olb, ovb = ocheck.to_oracle_batch(lb), ocheck.to_oracle_batch(vb)

    class _Fixed(torch.nn.Module):
        def __init__(self, keep):
            super().__init__()
            self.keep = keep.cpu().double()

        def forward(self, x):
            return x * self.keep * 1.25

    for k_, keep in enumerate(gk):
        setattr(oG.encoder, f"module_{4 * k_ + 3}", _Fixed(keep))
    for k_, keep in enumerate(dk):
        setattr(oD.encoder, f"module_{4 * k_ + 3}", _Fixed(keep))
    with torch.no_grad():
        ol, oh, os_ = oG(olb, ovb, z.cpu().double(), gnoise.cpu().double())
        osc = oD(olb, ovb, label.cpu().double().unsqueeze(0))
    assert_close(logits, ol, 1e-4, "philox-mode generator logits vs oracle with the exported masks")
    assert_close(soft, os_, 1e-4, "philox-mode label_soft")
    assert_close(score, osc, 2e-5, "philox-mode critic score")
    top2 = os_.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-3
    assert torch.equal(hard.argmax(1).cpu()[safe], oh.argmax(1)[safe])


# ------------------------------------------------------------------------------------------------------------
# TMA-gather variant of the aggregation forward
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ngraphs,C", [(1, 64), (1, 128), (10, 64)])
def test_gat_fwd_tma_gather_variant(ngraphs, C):
    """bg_gat_fwd_tma (cp.async.bulk.tensor tile::gather4 into an mbarrier ring) == the register-path kernel to rounding, and
    == the fp64 oracle (rel 1e-5) on a 1e5-voxel graph / a closed sub-graph of the 1e6-voxel batch; a graph with a high-degree
    row is refused loudly."""
    vb, edges, csr = _large(ngraphs)
    n = vb.num_nodes
    gen = torch.Generator().manual_seed(7 * C + ngraphs)
    h, s, d, b = (torch.randn(n, C, generator=gen), torch.randn(n, generator=gen), torch.randn(n, generator=gen), torch.randn(C, generator=gen))
    f = lambda t: t.to(DEV).contiguous()
    o_ref, m_ref, z_ref = lib.gat_fwd(csr, f(h), f(s), f(d), f(b))
    o, m, z = lib.gat_fwd_tma(csr, f(h), f(s), f(d), f(b))
    assert_close(o, o_ref.double(), 2e-6, "tma vs register path: out")
    assert torch.equal(m, m_ref)
    assert_close(z, z_ref.double(), 2e-6, "tma vs register path: z")
    if ngraphs == 1:
        T, sub = torch.arange(n), edges
    else:
        _, T, sub = _closed_subgraph(edges, n, 20_000)
    out = pyg.gat_core(h.double(), s.double(), d.double(), sub) + b.double()
    assert_close(o.cpu()[T], out[T], 1e-5, f"gat_fwd_tma N={n} C={C}")
    # twice the same bits
    o2, _, _ = lib.gat_fwd_tma(csr, f(h), f(s), f(d), f(b))
    assert torch.equal(o, o2)


def test_gat_fwd_tma_refuses_what_it_cannot_do():
    from test_kernels_gpu import _hub_graph
    n, edges, csr = _hub_graph()
    h, s, d = torch.randn(n, 64, device=DEV), torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    with pytest.raises(RuntimeError, match="max in-degree"):
        lib.gat_fwd_tma(csr, h, s, d, None)
    vb, edges, csr = _large(1)
    with pytest.raises(RuntimeError, match="C in"):
        lib.gat_fwd_tma(csr, torch.randn(vb.num_nodes, 32, device=DEV), torch.randn(vb.num_nodes, device=DEV), torch.randn(vb.num_nodes, device=DEV), None)
