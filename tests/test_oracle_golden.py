"""The oracle (and the product's data path) against vectors recorded from the UNMODIFIED reference
(tests/golden/reference_small.pt, made by oracle/make_golden.py from /root/reference).  CPU only."""
import os

import pytest
import torch

from building_gan_b200 import Configuration, graph, synth
from oracle import models as omodels
from oracle import pyg
from oracle import trainer as otrainer

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_small.pt"), weights_only=False)


def _batches(batch_cls=None):
    pairs = [synth.building_pair(i) for i in GOLD["ids"]]
    if batch_cls is None:
        return graph.collate_fn(pairs)
    return batch_cls.from_data_list([pyg.Data(**p[0]._fields) for p in pairs]), batch_cls.from_data_list(
        [pyg.Data(**p[1]._fields) for p in pairs])


@pytest.mark.parametrize("kind", ["product", "oracle"])
def test_preprocessing_and_collation_bit_exact(kind):
    """raw JSON -> processed tensors -> Data -> Batch: every tensor field equals what the reference's
    DataCreator + GraphDataset + collate_fn produced (data.py:16-163,215-461), dtype included."""
    lb, vb = _batches(None if kind == "product" else pyg.Batch)
    for name, batch in (("local", lb), ("voxel", vb)):
        for key, ref in GOLD[name].items():
            got = getattr(batch, key)
            assert got.dtype == ref.dtype, f"{name}.{key}: dtype {got.dtype} != {ref.dtype}"
            assert torch.equal(got, ref), f"{name}.{key} differs"
    assert vb.data_number == GOLD["voxel_data_number"] and lb.data_number == GOLD["local_data_number"]
    assert vb.num_graphs == 3 and vb.num_nodes == GOLD["voxel"]["x"].shape[0]
    one = vb[1]
    for key, ref in GOLD["voxel_graph1"].items():
        assert torch.equal(getattr(one, key), ref), f"voxel_batch[1].{key} differs"
    assert one.num_nodes == GOLD["voxel_graph1"]["x"].shape[0]


def _models():
    cfg = Configuration()
    G = omodels.OracleGenerator(cfg, 17, 12)
    D = omodels.OracleDiscriminator(cfg, 17, 12)
    G.load_state_dict(GOLD["G_state"], strict=True)   # same keys/shapes as the reference modules
    D.load_state_dict(GOLD["D_state"], strict=True)
    return cfg, G, D


def test_models_eval_forward_matches_reference():
    cfg, G, D = _models()
    lb, vb = _batches(pyg.Batch)
    G.eval(), D.eval()
    torch.manual_seed(1234)
    logits, hard, soft = G(lb, vb, GOLD["z"])
    e = GOLD["eval"]
    assert torch.allclose(logits, e["logits"], rtol=0, atol=1e-6)
    assert torch.allclose(soft, e["label_soft"], rtol=0, atol=1e-6)
    assert torch.equal(hard.argmax(1), e["label_hard"].argmax(1))
    assert torch.allclose(D(lb, vb, vb.types_onehot.unsqueeze(0)), e["d_real"], rtol=0, atol=1e-6)
    assert torch.allclose(D(lb, vb, e["label_hard"].unsqueeze(0)), e["d_fake"], rtol=0, atol=1e-6)


def test_parameter_counts_and_keys():
    _, G, D = _models()
    assert sum(p.numel() for p in G.parameters()) == 274185
    assert sum(p.numel() for p in D.parameters()) == 15665
    assert "encoder.module_0.lin.weight" in G.state_dict() and "encoder.module_1.mean_scale" in G.state_dict()


def _grad_close(model, ref_grads, tol):
    for k, p in model.named_parameters():
        if k not in ref_grads:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        scale = max(float(ref_grads[k].abs().max()), 1e-12)
        err = float((p.grad - ref_grads[k]).abs().max()) / scale
        assert err <= tol, f"{k}: rel err {err:.2e}"


def test_critic_loss_and_grads_match_reference_trainer():
    """trainer.py:291-332 with dropout active: same seed => same masks, same CPU draws."""
    cfg, G, D = _models()
    lb, vb = _batches(pyg.Batch)
    G.train(), D.train()
    torch.manual_seed(4321)
    with torch.no_grad():
        _, hard, soft = G(lb, vb, GOLD["z"])
    D.zero_grad()
    loss = otrainer.discriminator_loss(D, lb, vb, hard.unsqueeze(0), soft.unsqueeze(0), cfg)
    loss.backward()
    assert abs(float(loss) - float(GOLD["critic"]["d_loss"])) <= 1e-5 * abs(float(GOLD["critic"]["d_loss"]))
    _grad_close(D, GOLD["critic"]["grads"], 1e-4)


def test_generator_loss_and_grads_match_reference_trainer():
    cfg, G, D = _models()
    lb, vb = _batches(pyg.Batch)
    G.train(), D.train()
    torch.manual_seed(999)
    G.zero_grad(), D.zero_grad()
    logits, hard, _ = G(lb, vb, GOLD["z"])
    loss = otrainer.generator_loss(D, lb, vb, logits, hard.unsqueeze(0), cfg)
    loss.backward()
    assert abs(float(loss) - float(GOLD["gen"]["g_loss"])) <= 1e-5 * abs(float(GOLD["gen"]["g_loss"]))
    _grad_close(G, GOLD["gen"]["grads"], 1e-4)
