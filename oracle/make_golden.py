"""Generate tests/golden/*.pt by running the UNMODIFIED reference sources.

Run in the build container only (``python -m oracle.make_golden``): /root/reference does not exist
on the GPU box, so the vectors are committed as fixtures.  What runs here is the reference's own
code - ``building_gan/src/{config,data,models,trainer}.py`` imported from /root/reference - with

* ``oracle/pyg_shim`` standing in for the absent ``torch-geometric==2.6.1`` wheel (the op
  semantics inside the shim are the oracle's restatement, see oracle/pyg.py: that part stays
  "parity unpinned"), and
* inert stubs for matplotlib / pytz / IPython, which trainer.py imports at module scope
  (trainer.py:6,10,15,18,20) but the hot path never calls.

Pinned by the fixtures: raw JSON -> processed tensors (data.py:215-391,16-77), Data fields and
collation (data.py:118-163), generator/discriminator composition (models.py:14-245), critic loss,
gradient penalty and generator loss incl. the FAR loop (trainer.py:291-385), and parameter
gradients of both losses.
"""
from __future__ import annotations

import functools
import json
import os
import sys
import tempfile
import types
from unittest import mock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
    sys.path.insert(0, REF)
    import sklearn.metrics  # noqa: F401  (real; pulls pandas in BEFORE the stubs below could confuse its optional-dependency probe)
    import torch.utils.tensorboard  # noqa: F401
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "pytz", "IPython", "IPython.display",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        sys.modules.setdefault(name, mock.MagicMock(name=name))
    from building_gan.src import config, data, models, trainer  # noqa: E402  (the reference's own modules)
    return config, data, models, trainer


def _tensor_fields(batch):
    return {k: v.clone() for k, v in batch._store.items() if isinstance(v, torch.Tensor)}


def main() -> None:
    config, data, models, trainer = _import_reference()
    from workloads import synth  # synthetic raw JSON (the real dataset is an LFS pointer); neutral code, no product import

    os.makedirs(OUT, exist_ok=True)
    ids = [4001, 4002, 4003]
    with tempfile.TemporaryDirectory() as tmp:
        cfg = config.Configuration()
        cfg.DATA_PATH = os.path.join(tmp, "raw")
        cfg.GLOBAL_GRAPH_DATA_PATH = os.path.join(cfg.DATA_PATH, "global_graph_data")
        cfg.LOCAL_GRAPH_DATA_PATH = os.path.join(cfg.DATA_PATH, "local_graph_data")
        cfg.VOXEL_GRAPH_DATA_PATH = os.path.join(cfg.DATA_PATH, "voxel_data")
        cfg.SAVE_DATA_PATH = os.path.join(tmp, "processed")
        for p in (cfg.GLOBAL_GRAPH_DATA_PATH, cfg.LOCAL_GRAPH_DATA_PATH, cfg.VOXEL_GRAPH_DATA_PATH):
            os.makedirs(p)
        for i in ids:
            g, l, v = synth.raw_building(i)
            for path, prefix, obj in ((cfg.GLOBAL_GRAPH_DATA_PATH, "graph_global_", g),
                                      (cfg.LOCAL_GRAPH_DATA_PATH, "graph_local_", l),
                                      (cfg.VOXEL_GRAPH_DATA_PATH, "voxel_", v)):
                with open(os.path.join(path, f"{prefix}{i:06d}.json"), "w") as f:
                    json.dump(obj, f)
        data.DataCreator(cfg).create()  # reference preprocessing, data.py:394-461
        with mock.patch.object(torch, "load", functools.partial(torch.load, weights_only=False)):
            dataset = data.GraphDataset(cfg)  # reference Data construction, data.py:80-148
        pairs = [dataset[i] for i in range(len(ids))]
        local_b, voxel_b = data.GraphDataset.collate_fn(pairs)  # reference collation, data.py:156-163

    golden = {"ids": ids, "local": _tensor_fields(local_b), "voxel": _tensor_fields(voxel_b),
              "voxel_data_number": voxel_b.data_number, "local_data_number": local_b.data_number}
    g1 = voxel_b[1]
    golden["voxel_graph1"] = {k: v.clone() for k, v in g1._store.items() if isinstance(v, torch.Tensor)}

    cfg = config.Configuration()
    torch.manual_seed(777)
    G = models.VoxelGNNGenerator(cfg, local_b.x.shape[1], voxel_b.x.shape[1])
    D = models.VoxelGNNDiscriminator(cfg, local_b.x.shape[1], voxel_b.x.shape[1])
    with torch.no_grad():  # move the norm / bias parameters off their trivial init
        for m in list(G.modules()) + list(D.modules()):
            for name in ("bias", "mean_scale"):
                p = getattr(m, name, None)
                if isinstance(p, torch.nn.Parameter):
                    p.add_(0.1 * torch.randn_like(p))
    golden["G_state"] = {k: v.clone() for k, v in G.state_dict().items()}
    golden["D_state"] = {k: v.clone() for k, v in D.state_dict().items()}
    n = voxel_b.num_nodes
    z = torch.randn(1, n, cfg.Z_DIM)
    golden["z"] = z.clone()

    # -- eval-mode forward (no dropout): RNG consumed only by the Gumbel noise
    G.eval(), D.eval()
    torch.manual_seed(1234)
    logits, hard, soft = G(local_b, voxel_b, z)
    golden["eval"] = {"logits": logits.detach().clone(), "label_hard": hard.detach().clone(),
                      "label_soft": soft.detach().clone(),
                      "d_real": D(local_b, voxel_b, voxel_b.types_onehot.unsqueeze(0)).detach().clone(),
                      "d_fake": D(local_b, voxel_b, hard.detach().unsqueeze(0)).detach().clone()}

    # -- train-mode losses through the reference TrainerHelper methods (dropout masks + CPU RNG draws)
    helper = trainer.TrainerHelper.__new__(trainer.TrainerHelper)
    helper.generator, helper.discriminator, helper.configuration = G, D, cfg
    G.train(), D.train()
    torch.manual_seed(4321)
    with torch.no_grad():
        _, hard_c, soft_c = G(local_b, voxel_b, z)
    D.zero_grad()
    d_loss = helper._compute_discriminator_loss(local_b, voxel_b, hard_c.unsqueeze(0), soft_c.unsqueeze(0))
    d_loss.backward()
    golden["critic"] = {"d_loss": d_loss.detach().clone(),
                        "grads": {k: p.grad.clone() for k, p in D.named_parameters()}}
    torch.manual_seed(999)
    G.zero_grad(), D.zero_grad()
    logits, hard_g, soft_g = G(local_b, voxel_b, z)
    g_loss = helper._compute_generator_loss(local_b, voxel_b, logits, hard_g.unsqueeze(0))
    g_loss.backward()
    golden["gen"] = {"g_loss": g_loss.detach().clone(),
                     "grads": {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}}
    # eval-mode losses (no dropout) - what the CUDA path is compared with without mask injection
    G.eval(), D.eval()
    torch.manual_seed(555)
    D.zero_grad()
    with torch.no_grad():
        _, hard_e, soft_e = G(local_b, voxel_b, z)
    d_loss_e = helper._compute_discriminator_loss(local_b, voxel_b, hard_e.unsqueeze(0), soft_e.unsqueeze(0))
    d_loss_e.backward()
    golden["critic_eval"] = {"d_loss": d_loss_e.detach().clone(),
                             "grads": {k: p.grad.clone() for k, p in D.named_parameters()}}
    torch.save(golden, os.path.join(OUT, "reference_small.pt"))
    size = os.path.getsize(os.path.join(OUT, "reference_small.pt"))
    print(f"wrote tests/golden/reference_small.pt ({size / 1e6:.2f} MB): N={n}, d_loss={float(d_loss):.6f}, "
          f"g_loss={float(g_loss):.6f}")


if __name__ == "__main__":
    main()
