"""The host-side half of `step.compute_metrics`: macro precision / recall / F1 from confusion counts must equal
sklearn called the way trainer.py:387-443 calls it (average="macro", zero_division=0).  Pure torch: no GPU needed (the
counts themselves come from bg_segment_confusion, covered by tests/test_trainer_ops_gpu.py)."""
import numpy as np
import pytest
import torch
from sklearn import metrics as skm

from building_gan_b200 import step


def _cm(y, yp, k=7):
    cm = torch.zeros(k, k, dtype=torch.int32)
    for t, p in zip(y, yp):
        cm[t, p] += 1
    return cm


@pytest.mark.parametrize("seed,classes_true,classes_pred", [(0, 7, 7), (1, 3, 7), (2, 7, 2), (3, 1, 1), (4, 5, 4)])
def test_macro_scores_from_confusion_counts(seed, classes_true, classes_pred):
    rng = np.random.default_rng(seed)
    y = rng.integers(0, classes_true, 500)
    yp = rng.integers(0, classes_pred, 500)
    prec, rec, f1 = step._prf(_cm(y, yp))
    assert abs(float(prec) - skm.precision_score(y, yp, average="macro", zero_division=0)) < 1e-12
    assert abs(float(rec) - skm.recall_score(y, yp, average="macro", zero_division=0)) < 1e-12
    assert abs(float(f1) - skm.f1_score(y, yp, average="macro", zero_division=0)) < 1e-12


def test_batched_confusion_matrices_and_empty_segment():
    rng = np.random.default_rng(9)
    cms, want = [], []
    for n in (40, 0, 7):
        y, yp = rng.integers(0, 7, n), rng.integers(0, 7, n)
        cms.append(_cm(y, yp))
        want.append(skm.f1_score(y, yp, average="macro", zero_division=0) if n else 0.0)
    _, _, f1 = step._prf(torch.stack(cms))
    assert f1.shape == (3,)
    for a, b in zip(f1.tolist(), want):
        assert abs(a - b) < 1e-12
