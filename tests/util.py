"""Shared helpers for the parity tests."""
import torch

import building_gan_b200 as bg
from building_gan_b200 import graph, synth


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max |a - ref| / max(|ref|, tiny): the fp32 parity metric (north_star: rel 1e-5)."""
    ref = ref.double().cpu()
    a = a.double().cpu()
    scale = max(ref.abs().max().item(), 1e-30)
    return (a - ref).abs().max().item() / scale


def assert_close(a, ref, tol=1e-5, what=""):
    assert a.shape == ref.shape, f"{what}: shape {tuple(a.shape)} != {tuple(ref.shape)}"
    e = rel_err(a, ref)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


def small_batch(ids=(11, 12, 13), shuffle=False):
    pairs = [synth.building_pair(i, shuffle=shuffle) for i in ids]
    return graph.collate_fn(pairs)
