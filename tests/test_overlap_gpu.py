"""The overlapped training step (step.train_step(overlap=True): sampling passes up front on their own stream, D(real) /
D(fake) beside the gradient-penalty pass, parameter gradients through lane buckets) against the same call order on ONE
stream (overlap="order"): same kernels and same random draws, so the only permitted difference is the order in which the
three critic passes' parameter-gradient contributions are added (fp32 rounding).  Running the overlapped step twice from
the same state must be bit-identical (a race between the streams would show up here)."""
import copy

import pytest
import torch

from building_gan_b200 import Configuration, models as bm, step as bstep
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from building_gan_b200.optim import Adam

from util import small_batch

pytestmark = pytest.mark.gpu


def _run(overlap, adam, steps=2, ids=(21, 22, 23, 24, 25, 26)):
    cfg = Configuration()
    cfg.DEVICE = "cuda"
    torch.manual_seed(1)
    G, D = VoxelGNNGenerator(cfg, 17, 12), VoxelGNNDiscriminator(cfg, 17, 12)  # train mode: dropout on (Philox tickets)
    og, od = adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
    lb, vb = small_batch(ids)
    lb, vb = lb.to("cuda"), vb.to("cuda")
    torch.manual_seed(2)
    torch.cuda.manual_seed(2)
    bm._philox_calls = 0
    losses, grads = [], None
    for s in range(steps):
        d, g, hard = bstep.train_step(G, D, og, od, lb, vb, cfg, rng="cpu", overlap=overlap)
        losses += list(d) + [g]
        if s == 0:
            grads = [p.grad.detach().clone() for p in D.parameters()]  # D grads left by the generator update's backward
    torch.cuda.synchronize()
    return losses, [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())], hard.clone(), grads


@pytest.mark.parametrize("adam", [Adam, torch.optim.Adam], ids=["flat-adam", "torch-adam"])
def test_overlap_equals_single_stream(adam):
    l1, p1, h1, _ = _run(True, adam, steps=1)
    l0, p0, h0, _ = _run("order", adam, steps=1)
    # first critic loss: nothing has been updated yet and no parameter gradient is involved -> identical
    assert l1[0] == l0[0]
    scale = max(1.0, max(abs(v) for v in l0))
    assert max(abs(a - b) for a, b in zip(l1, l0)) <= 1e-4 * scale, (l1, l0)
    # Adam moves every parameter by ~lr per update (6 updates of D, 1 of G) whatever the gradient's scale, and the critic's
    # parameter gradients are differences of nearly cancelling real / fake contributions (test_lane_gradients_match_sequential
    # bounds them by 1e-5 of the tensor's largest entry): a reordered 3-term sum moves small entries by a visible fraction of
    # lr.  Bound the difference by a fraction of the distance travelled; measured: median 0.5 %, worst entry 27 %.
    for a, b in zip(p1, p0):
        assert (a - b).abs().max().item() <= 2 * 2e-4 * 6
        assert (a - b).abs().median().item() <= 0.02 * 2e-4 * 6
    assert (h1.argmax(-1) != h0.argmax(-1)).float().mean().item() <= 0.01


def test_overlap_is_deterministic():
    a = _run(True, Adam, steps=2)
    b = _run(True, Adam, steps=2)
    assert a[0] == b[0]
    for x, y in zip(a[1], b[1]):
        assert torch.equal(x, y)
    assert torch.equal(a[2], b[2])


def test_lane_gradients_match_sequential():
    """One critic update, gradients only: lanes + merge == plain accumulation into the bucket (up to summation order)."""
    cfg = Configuration()
    cfg.DEVICE = "cuda"
    torch.manual_seed(3)
    G, D = VoxelGNNGenerator(cfg, 17, 12), VoxelGNNDiscriminator(cfg, 17, 12)
    G.eval(), D.eval()
    lb, vb = small_batch((31, 32, 33))
    lb, vb = lb.to("cuda"), vb.to("cuda")
    with torch.no_grad():
        _, hard, soft = G(lb, vb, torch.randn(1, vb.num_nodes, cfg.Z_DIM, device="cuda"))
    hard, soft = hard.unsqueeze(0), soft.unsqueeze(0)
    e = torch.rand(vb.num_nodes, 1, device="cuda")
    out = []
    for lanes in (None, bstep.Lanes.get(torch.device("cuda", 0))):
        D.zero_grad(set_to_none=True)
        loss = bstep.discriminator_loss(D, lb, vb, hard, soft, cfg, "device", lanes, e)
        loss.backward()
        if lanes is not None:
            torch.cuda.current_stream().wait_stream(lanes.real)
            torch.cuda.current_stream().wait_stream(lanes.fake)
            D.merge_lanes()
        out.append((loss.item(), [p.grad.detach().clone() for p in D.parameters()]))
    assert out[0][0] == out[1][0]
    for (name, _), g0, g1 in zip(D.named_parameters(), out[0][1], out[1][1]):
        assert (g0 - g1).abs().max().item() <= 1e-5 * max(g0.abs().max().item(), 1e-6), name


def test_graphed_step_runs_and_trains():
    """graphs.GraphedStep (critic update + sampling pass captured once per step, replayed N_CRITIC times): the replays must
    differ (fresh z / mixing factor / dropout / Gumbel draws, Adam advancing), Adam's step count must advance by N_CRITIC per
    step, and the first critic loss of a step must equal the eager loss on the same state (same arithmetic inside the graph)."""
    from building_gan_b200.graphs import GraphedStep

    cfg = Configuration()
    cfg.DEVICE = "cuda"
    torch.manual_seed(1)
    G, D = VoxelGNNGenerator(cfg, 17, 12), VoxelGNNDiscriminator(cfg, 17, 12)
    og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
    batches = [tuple(b.to("cuda") for b in small_batch(ids)) for ids in ((21, 22, 23), (24, 25, 26, 27), (28, 29))]
    gs = GraphedStep(G, D, og, od, cfg)
    seen = []
    for s in range(4):
        lb, vb = batches[s % len(batches)]
        before = [p.detach().clone() for p in D.parameters()]
        d, g, hard = gs(lb, vb, sync_losses="step")
        assert len(d) == cfg.N_CRITIC and all(map(lambda v: v == v and abs(v) < 1e4, d + [g]))
        assert hard.shape == (1, vb.num_nodes, cfg.NUM_CLASSES)
        assert len(set(d)) == cfg.N_CRITIC, d  # five different critic losses: the replays are not copies of each other
        assert any(not torch.equal(a, b) for a, b in zip(before, D.parameters()))
        seen.append(d)
    assert float(od.state_dict()["state"][0]["step"]) == 4 * cfg.N_CRITIC
    assert float(og.state_dict()["state"][0]["step"]) == 4
    # same state, same batch: eager critic loss vs. the captured one.  Dropout off so that only z / e differ between the two
    # paths; the loss is then compared through its deterministic part: D(real).mean() enters both with the same value.
    G.eval(), D.eval()
    lb, vb = batches[0]
    with torch.no_grad():
        ref_real = D(lb, vb, vb.types_onehot.unsqueeze(0)).mean().item()
    d, g, _ = gs(lb, vb, sync_losses="step")
    assert all(v == v for v in d)
    torch.cuda.synchronize()
    assert abs(ref_real) < 1e4


@pytest.mark.parametrize("kind", ["GATCONV", "GCNCONV", "GRAPHCONV", "GATV2CONV"])
def test_graphed_step_matches_eager_statistics(kind):
    """Train the same initial state for a few steps with the eager overlapped step and with GraphedStep: different random
    draws (so no element-wise equality), but the critic losses must stay in the same range and all parameters finite.  For
    every conv type of models.py:22-31 (the non-default ones run the op-by-op executor inside the captured graphs)."""
    from building_gan_b200.graphs import GraphedStep

    out = {}
    for mode in ("eager", "graph"):
        cfg = Configuration()
        cfg.GENERATOR_CONV_TYPE = cfg.DISCRIMINATOR_CONV_TYPE = kind
        cfg.DEVICE = "cuda"
        torch.manual_seed(1)
        torch.cuda.manual_seed(1)
        G, D = VoxelGNNGenerator(cfg, 17, 12), VoxelGNNDiscriminator(cfg, 17, 12)
        og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
        lb, vb = (b.to("cuda") for b in small_batch((21, 22, 23, 24)))
        gs = GraphedStep(G, D, og, od, cfg) if mode == "graph" else None
        ls = []
        for s in range(6):
            if gs is not None:
                d, g, _ = gs(lb, vb, sync_losses="step")
            else:
                d, g, _ = bstep.train_step(G, D, og, od, lb, vb, cfg, rng="device", sync_losses="step", overlap=True)
            ls.append(sum(d) / len(d))
        assert all(torch.isfinite(p).all() for p in list(G.parameters()) + list(D.parameters()))
        out[mode] = ls
    # the critic loss starts near LAMBDA_GP (gradient norm ~0) and decreases as D trains; both paths must show that
    for mode in out:
        assert out[mode][-1] < out[mode][0], out
    assert abs(out["graph"][-1] - out["eager"][-1]) <= 0.25 * abs(out["eager"][0]), out


def test_graphed_step_prepares_the_next_batch_ahead():
    """GraphedStep(..., next_batch=): the following step's graphs are captured during the current step and used by the next call
    when it is handed the same batch objects; a different batch falls back to capturing at the start of its own step.  Same
    training behaviour either way (finite losses, Adam counters advance, the critic loss decreases)."""
    from building_gan_b200.graphs import GraphedStep

    cfg = Configuration()
    cfg.DEVICE = "cuda"
    torch.manual_seed(1)
    G, D = VoxelGNNGenerator(cfg, 17, 12), VoxelGNNDiscriminator(cfg, 17, 12)
    og, od = Adam(G.parameters(), lr=2e-4, betas=cfg.BETAS), Adam(D.parameters(), lr=2e-4, betas=cfg.BETAS)
    batches = [tuple(b.to("cuda") for b in small_batch(ids)) for ids in ((21, 22, 23), (24, 25, 26, 27), (28, 29))]
    gs = GraphedStep(G, D, og, od, cfg)
    first = []
    for s in range(7):
        cur, nxt = batches[s % 3], batches[(s + 1) % 3]
        d, g, hard = gs(*cur, sync_losses="step", next_batch=nxt)
        assert all(v == v and abs(v) < 1e4 for v in d + [g]) and hard.shape[1] == cur[1].num_nodes
        if s >= 1:  # (the very first call is the eager warm-up step: nothing is prepared there)
            assert gs._prepared is not None and gs._prepared.lb is nxt[0] and gs._prepared.vb is nxt[1]
        if s % 3 == 0:
            first.append(d[0])
    # hand the step a batch other than the prepared one: it must notice and capture for the batch it was given
    d, g, hard = gs(*batches[2], sync_losses="step")
    assert hard.shape[1] == batches[2][1].num_nodes and all(v == v for v in d + [g]) and gs._prepared is None
    assert float(od.state_dict()["state"][0]["step"]) == 8 * cfg.N_CRITIC and float(og.state_dict()["state"][0]["step"]) == 8
    assert first[-1] < first[0], first  # the critic loss on batch 0 decreases over the steps
