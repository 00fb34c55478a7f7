"""The "drops into trainer.py unchanged" claim, checked with the reference's OWN trainer code.

Runs only where /root/reference exists (the build container; the GPU box has no reference checkout): the UNMODIFIED
``building_gan/src/trainer.py`` is imported (``oracle/make_golden._import_reference``: the pyg shim stands in for the
torch_geometric wheel, inert stubs for matplotlib / pytz / IPython) and its ``TrainerHelper`` methods are driven with

* this package's ``Batch`` objects (``graph.collate_fn``) - the duck type must satisfy every attribute, slice and
  ``voxel_graph[gi]`` access trainer.py:291-443 makes - against the oracle's ``Batch`` of the same buildings, and
* this package's restatement of those methods (``trainer_helper.ReferenceTrainerHelper``, the ``dropin`` block of
  bench.py) against the reference's, bit for bit on the CPU with the oracle models.

The models themselves need a GPU; their side of the contract (signatures, state-dict keys) is checked structurally here
and numerically in tests/test_models_gpu.py.
"""
from __future__ import annotations

import inspect
import os

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "building_gan", "src")),
                                reason="needs the reference checkout (build container only)")


@pytest.fixture(scope="module")
def ref():
    from oracle.make_golden import _import_reference
    config, data, models, trainer = _import_reference()
    return {"config": config, "data": data, "models": models, "trainer": trainer}


@pytest.fixture(scope="module")
def world(ref):
    from building_gan_b200 import graph, synth
    from oracle import models as omodels, pyg as opyg
    cfg = ref["config"].Configuration()
    cfg.DEVICE = "cpu"
    pairs = [synth.building_pair(i) for i in (4001, 4002, 4003)]
    lb, vb = graph.collate_fn(pairs)                              # the product's Batch (CPU tensors)
    olb = opyg.Batch.from_data_list([opyg.Data(**p[0]._fields) for p in pairs])
    ovb = opyg.Batch.from_data_list([opyg.Data(**p[1]._fields) for p in pairs])
    torch.manual_seed(777)
    G, D = omodels.OracleGenerator(cfg, 17, 12), omodels.OracleDiscriminator(cfg, 17, 12)
    return cfg, (lb, vb), (olb, ovb), G, D


def _helper(cls, G, D, cfg):
    h = cls.__new__(cls)
    h.generator, h.discriminator, h.configuration = G, D, cfg
    return h


def test_reference_trainer_helper_runs_on_the_product_batch(ref, world):
    """trainer.py:291-443 unmodified, product ``Batch`` in, same numbers as with the oracle's PyG-style ``Batch``."""
    cfg, (lb, vb), (olb, ovb), G, D = world
    helper = _helper(ref["trainer"].TrainerHelper, G, D, cfg)
    z = torch.randn(1, vb.num_nodes, cfg.Z_DIM)
    outs = []
    for l, v in ((lb, vb), (olb, ovb)):
        torch.manual_seed(11)
        G.zero_grad(), D.zero_grad()
        with torch.no_grad():
            _, hard, soft = G(l, v, z)
        d_loss = helper._compute_discriminator_loss(l, v, hard.unsqueeze(0), soft.unsqueeze(0))
        d_loss.backward()
        logits, hard_g, _ = G(l, v, z)
        g_loss = helper._compute_generator_loss(l, v, logits, hard_g.unsqueeze(0))
        g_loss.backward()
        met = helper._compute_metrics(v, hard_g.unsqueeze(0))
        outs.append((d_loss.detach(), g_loss.detach(), met, [p.grad.clone() for p in D.parameters()]))
    (d0, g0, m0, gr0), (d1, g1, m1, gr1) = outs
    assert torch.equal(d0, d1) and torch.equal(g0, g1)
    assert m0[0] == m1[0] and list(m0[1]) == list(m1[1]) and m0[2:] == m1[2:]
    assert all(torch.equal(a, b) for a, b in zip(gr0, gr1))
    # trainer.py:464: the data_number consistency assert of the training loop holds on the product batches too
    assert [set(d) for d in lb.data_number] == [set(d) for d in vb.data_number]


def test_restated_helper_equals_the_reference_helper(ref, world):
    """``ReferenceTrainerHelper`` (what bench.py's ``dropin`` block times) == the reference's ``TrainerHelper``, bit for bit."""
    from building_gan_b200.trainer_helper import ReferenceTrainerHelper
    cfg, (lb, vb), _, G, D = world
    theirs, ours = _helper(ref["trainer"].TrainerHelper, G, D, cfg), _helper(ReferenceTrainerHelper, G, D, cfg)
    z = torch.randn(1, vb.num_nodes, cfg.Z_DIM)
    res = []
    for h in (theirs, ours):
        torch.manual_seed(23)
        G.zero_grad(), D.zero_grad()
        with torch.no_grad():
            _, hard, soft = G(lb, vb, z)
        gp = h._compute_gradient_penalty(lb, vb, soft.unsqueeze(0))
        d_loss = h._compute_discriminator_loss(lb, vb, hard.unsqueeze(0), soft.unsqueeze(0))
        logits, hard_g, _ = G(lb, vb, z)
        g_loss = h._compute_generator_loss(lb, vb, logits, hard_g.unsqueeze(0))
        (d_loss + g_loss).backward()
        res.append((gp.detach(), d_loss.detach(), g_loss.detach(), h._compute_metrics(vb, hard_g.unsqueeze(0)),
                    [p.grad.clone() for p in list(G.parameters()) + list(D.parameters()) if p.grad is not None]))
    a, b = res
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert a[3][0] == b[3][0] and list(a[3][1]) == list(b[3][1]) and a[3][2:] == b[3][2:]
    assert len(a[4]) == len(b[4]) and all(torch.equal(x, y) for x, y in zip(a[4], b[4]))


def test_restated_train_batch_equals_the_reference_loop_body(ref, world):
    """One batch of trainer.py:459-503 - the reference's ``_train_each_epoch`` body, run through a one-batch dataloader -
    against ``ReferenceTrainerHelper.train_batch``: same losses, same metrics, same updated weights."""
    import copy
    from building_gan_b200.trainer_helper import ReferenceTrainerHelper
    cfg, (lb, vb), _, G0, D0 = world
    results = []
    for which in ("reference", "restated"):
        G, D = copy.deepcopy(G0), copy.deepcopy(D0)
        og = torch.optim.Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
        od = torch.optim.Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
        torch.manual_seed(5)
        if which == "reference":
            t = _helper(ref["trainer"].Trainer, G, D, cfg)
            t.optimizer_generator, t.optimizer_discriminator = og, od

            class _DL:
                train_dataloader = [(lb, vb)]
            t.dataloaders = _DL()
            fn = ref["trainer"].Trainer._train_each_epoch
            fn = getattr(fn, "__wrapped__", fn)
            out = fn(t)
            out = out[0] if isinstance(out, tuple) and len(out) == 2 and isinstance(out[0], tuple) else out
            results.append((out, [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())]))
        else:
            h = _helper(ReferenceTrainerHelper, G, D, cfg)
            h.optimizer_generator, h.optimizer_discriminator = og, od
            d_losses, g_loss, met = h.train_batch(lb, vb)
            results.append(((d_losses, g_loss, met), [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())]))
    (ref_out, ref_w), ((d_losses, g_loss, met), our_w) = results
    assert all(torch.equal(a, b) for a, b in zip(ref_w, our_w)), "updated weights differ from the reference loop body"
    # the reference returns epoch means (one batch here): g_loss mean, d_loss mean, then the metric means
    flat = [float(v) for v in ref_out if isinstance(v, (int, float)) or (hasattr(v, "ndim") and getattr(v, "ndim", 1) == 0)]
    # (trainer.py:504-520: torch.tensor(list).mean().item() - an fp32 round trip of the per-step floats)
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
    want = (f32(g_loss), float(torch.tensor(d_losses).mean()), f32(met[0]), min(met[1]), f32(met[2]), f32(met[3]), f32(met[4]))
    assert tuple(flat) == want, (flat, want)


def test_mixin_and_models_keep_the_reference_signatures(ref):
    """Method names / parameter lists of the mix-in == the reference TrainerHelper's; model constructors and ``forward``
    parameter names == the reference models' (models.py:15,119,159,229)."""
    from building_gan_b200 import models as pm
    from building_gan_b200.trainer_helper import ReferenceTrainerHelper, TrainerHelper
    theirs = ref["trainer"].TrainerHelper
    for name in ("_compute_gradient_penalty", "_compute_discriminator_loss", "_compute_generator_loss", "_compute_metrics"):
        want = list(inspect.signature(getattr(theirs, name)).parameters)
        for cls in (ReferenceTrainerHelper, TrainerHelper):
            assert list(inspect.signature(getattr(cls, name)).parameters) == want, (cls.__name__, name)
    for cls in ("VoxelGNNGenerator", "VoxelGNNDiscriminator"):
        r, p = getattr(ref["models"], cls), getattr(pm, cls)
        assert list(inspect.signature(r.__init__).parameters) == list(inspect.signature(p.__init__).parameters)
        rf = list(inspect.signature(r.forward).parameters)
        pf = list(inspect.signature(p.forward).parameters)
        assert pf[:len(rf)] == rf, (cls, rf, pf)  # extra trailing keyword arguments (noise / mask injection) are optional
        extra = [inspect.signature(p.forward).parameters[k] for k in pf[len(rf):]]
        assert all(e.default is not inspect.Parameter.empty for e in extra)


def test_state_dict_keys_match_the_reference_models(ref):
    """Same parameter names, order and shapes as the reference modules built on the pyg shim (states.pt interchange,
    trainer.py:715-736) - for every conv type of models.py:22-31."""
    from building_gan_b200 import Configuration
    from building_gan_b200 import models as pm
    for kind in ("GATCONV", "GCNCONV", "GRAPHCONV", "GATV2CONV"):
        rcfg = ref["config"].Configuration()
        rcfg.DEVICE = "cpu"
        rcfg.GENERATOR_CONV_TYPE = rcfg.DISCRIMINATOR_CONV_TYPE = kind
        pcfg = Configuration()
        pcfg.DEVICE = "cpu"
        pcfg.GENERATOR_CONV_TYPE = pcfg.DISCRIMINATOR_CONV_TYPE = kind
        for cls in ("VoxelGNNGenerator", "VoxelGNNDiscriminator"):
            torch.manual_seed(1)
            r = getattr(ref["models"], cls)(rcfg, 17, 12)
            torch.manual_seed(1)
            p = getattr(pm, cls)(pcfg, 17, 12)
            rs, ps = r.state_dict(), p.state_dict()
            assert list(rs) == list(ps), (kind, cls)
            for k in rs:
                assert rs[k].shape == ps[k].shape, (kind, cls, k)
                assert torch.equal(rs[k], ps[k]), f"{kind} {cls} {k}: same seed must give the same initial weights"
