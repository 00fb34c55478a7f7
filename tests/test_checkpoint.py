"""`states.pt` interchange (trainer.py:628-636 load, :715-736 save).

The checkpoint is a dict of the two models' state_dicts, the two Adam state_dicts and the generator's cosine scheduler
state_dict.  Optimiser state is keyed by parameter POSITION, so interchange needs the same parameter order and shapes,
not just the same names.  Checked against the oracle models (whose keys are pinned on the unmodified reference by
tests/test_oracle_golden.py), for every conv type, on the CPU (construction and state handling need no GPU)."""
import io

import pytest
import torch

from building_gan_b200 import Configuration
from building_gan_b200.models import VoxelGNNDiscriminator, VoxelGNNGenerator
from oracle import models as omodels


def _bundle(G, D, cfg):
    og = torch.optim.Adam(G.parameters(), lr=cfg.LEARNING_RATE_GENERATOR, betas=cfg.BETAS)
    od = torch.optim.Adam(D.parameters(), lr=cfg.LEARNING_RATE_DISCRIMINATOR, betas=cfg.BETAS)
    sg = torch.optim.lr_scheduler.CosineAnnealingLR(og, T_max=10)
    return og, od, sg


def _fake_train(G, D, og, od, sg, seed):
    """Populate Adam's exp_avg / exp_avg_sq / step with recognisable values (no forward pass needed)."""
    g = torch.Generator().manual_seed(seed)
    for model, opt in ((G, og), (D, od)):
        for p in model.parameters():
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
    sg.step()


@pytest.mark.parametrize("kind", ["GATCONV", "GCNCONV", "GRAPHCONV", "GATV2CONV"])
@pytest.mark.parametrize("direction", ["ours->reference", "reference->ours"])
def test_states_pt_round_trip(kind, direction):
    cfg = Configuration()
    cfg.DEVICE = "cpu"
    cfg.GENERATOR_CONV_TYPE = cfg.DISCRIMINATOR_CONV_TYPE = kind
    torch.manual_seed(3)
    ours = (VoxelGNNGenerator(cfg, 17, 12), VoxelGNNDiscriminator(cfg, 17, 12))
    ref = (omodels.OracleGenerator(cfg, 17, 12), omodels.OracleDiscriminator(cfg, 17, 12))
    src, dst = (ours, ref) if direction == "ours->reference" else (ref, ours)
    # same parameter ORDER and shapes (optimizer state is positional)
    for a, b in zip(src, dst):
        na, nb = list(a.named_parameters()), list(b.named_parameters())
        assert [k for k, _ in na] == [k for k, _ in nb]
        assert [tuple(v.shape) for _, v in na] == [tuple(v.shape) for _, v in nb]
    sog, sod, ssg = _bundle(*src, cfg)
    _fake_train(*src, sog, sod, ssg, seed=5)
    buf = io.BytesIO()
    torch.save({"epoch_start": 3, "generator": src[0].state_dict(), "discriminator": src[1].state_dict(),
                "optimizer_generator": sog.state_dict(), "optimizer_discriminator": sod.state_dict(),
                "scheduler_generator": ssg.state_dict()}, buf)  # trainer.py:715-736
    buf.seek(0)
    states = torch.load(buf, weights_only=False)
    dog, dod, dsg = _bundle(*dst, cfg)
    dst[0].load_state_dict(states["generator"])  # trainer.py:630-634
    dst[1].load_state_dict(states["discriminator"])
    dog.load_state_dict(states["optimizer_generator"])
    dod.load_state_dict(states["optimizer_discriminator"])
    dsg.load_state_dict(states["scheduler_generator"])
    for a, b in zip(src, dst):
        for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
            assert torch.equal(v, w), k
    for so, do, sm, dm in ((sog, dog, src[0], dst[0]), (sod, dod, src[1], dst[1])):
        for ps, pd in zip(sm.parameters(), dm.parameters()):
            for key in ("exp_avg", "exp_avg_sq"):
                assert torch.equal(so.state[ps][key], do.state[pd][key])
            assert float(so.state[ps]["step"]) == float(do.state[pd]["step"]) == 1.0
    assert dsg.last_epoch == ssg.last_epoch == 1 and dog.param_groups[0]["lr"] == sog.param_groups[0]["lr"]
