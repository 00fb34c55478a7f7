"""CPU/torch ORACLE for the Building-GAN voxel-graph message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference /
``cpu_baseline`` leg may import it, and there only as the checker or as the
timed reference arm - never as the thing shipped.  The product package
(``building_gan_b200``) must not import this package.

What it is: a plain-PyTorch restatement of

* the reference's ``building_gan/src/models.py:14-245`` (generator and
  discriminator), ``building_gan/src/trainer.py:291-385,459-495`` (WGAN-GP
  losses, gradient penalty, one optimisation step) and
  ``building_gan/src/data.py:156-163`` (collation), and
* the third-party ops those files call, which are NOT vendored under
  ``/root/reference``: ``torch-geometric==2.6.1`` (``requirements.txt:11``):
  ``GATConv``, ``GraphNorm``, ``GCNConv``, ``GraphConv``, ``GATv2Conv``,
  ``nn.Sequential``, ``Data``/``Batch.from_data_list``, ``utils.softmax``,
  ``utils.scatter`` (pure-torch branch, no ``torch_scatter`` pinned).

PARITY PINNING STATUS
---------------------
* Model/loss composition (concat orders, layer order, loss glue, gradient
  penalty): PINNED.  ``oracle/make_golden.py`` imports the UNMODIFIED
  reference ``models.py``/``trainer.py`` from ``/root/reference`` (with
  ``oracle/pyg_shim`` standing in for the absent torch_geometric wheel) and
  records input/output vectors into ``tests/golden/``; ``tests/test_oracle_golden.py``
  checks this restatement against them.
* torch_geometric op semantics (GATConv / GraphNorm / Batch): **parity
  unpinned** - the wheel cannot be installed here (no network) and the
  reference ships no golden vectors, tests or readable checkpoint (all LFS
  pointers).  They are restated from PyG 2.6.1's published algorithm and
  validated against dense-matrix definitions, hand-computed 3-node known
  answers and fp64 gradcheck (``tests/test_oracle_pyg.py``).
"""
