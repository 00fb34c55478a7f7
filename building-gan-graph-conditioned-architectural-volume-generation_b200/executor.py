"""Hand-composed forward / backward / second-order-backward passes of the generator and the
discriminator on the libbgb200 kernels.

One pass = a fixed sequence of C-ABI calls (no torch ops on activations, no host syncs, no
data-dependent control flow), so it is CUDA-graph capturable.  The math mirrors the reference
layer by layer (models.py:119-155 generator, :229-245 discriminator); the derivative structure is
the one autograd builds for trainer.py:291-385:

* ``*_forward``   saves what the backward needs;
* ``*_backward``  = first-order VJP (input gradient + parameter gradients);
* ``disc_backward2`` = VJP of the discriminator's *input-gradient* (WGAN-GP, trainer.py:306-312
  create_graph=True): a second-order sweep through the backward ops followed by a first-order
  sweep through the forward ops with the per-layer cotangents injected.

Parameter gradients are written into a flat fp32 buffer laid out like ``ParamLayout`` (one bucket
per model: also the NCCL all-reduce payload).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import lib
from .lib import ACT_LRELU, ACT_NONE, ACT_RELU

KEEP_P = 0.8          # nn.Dropout(0.2) hard-coded in the reference (models.py:75,85,195,205)
KEEP_SCALE = 1.0 / KEEP_P


class ParamLayout:
    """Maps parameter names to slices of one flat fp32 buffer (grad bucket).  ``groups`` lists names that
    must sit back to back without padding (GraphNorm's weight/bias/mean_scale, GATConv's att_src/att_dst:
    the kernels emit them as one [3,C] / [2,C] block); every other slice starts 16-byte aligned."""

    def __init__(self, named_params: Sequence[Tuple[str, Tensor]], groups: Sequence[Sequence[str]] = ()):
        follow = {}
        for g in groups:
            for a, b in zip(g[:-1], g[1:]):
                follow[b] = a
        self.names, self.shapes, self.offsets, off = [], {}, {}, 0
        prev = None
        for name, p in named_params:
            if follow.get(name) != prev or prev is None:
                off = (off + 3) // 4 * 4
            self.names.append(name)
            self.shapes[name] = tuple(p.shape)
            self.offsets[name] = off
            off += p.numel()
            prev = name
        self.total = (off + 3) // 4 * 4

    def view(self, flat: Tensor, name: str) -> Tensor:
        shape = self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        return flat[self.offsets[name]: self.offsets[name] + n].view(shape)


class DenseSpec:
    """One Linear (+LayerNorm +activation).  ``prefix`` is the state-dict prefix of the Linear;
    ``ln_prefix`` of its LayerNorm (or None)."""

    def __init__(self, lin: str, ln: Optional[str], act: int):
        self.lin, self.ln, self.act = lin, ln, act


class ConvSpec:
    def __init__(self, conv: str, norm: str, cin: int, cout: int):
        self.conv, self.norm, self.cin, self.cout = conv, norm, cin, cout


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def dense_forward(P: Dict[str, Tensor], spec: DenseSpec, segs, save: bool):
    ln = (P[spec.ln + ".weight"], P[spec.ln + ".bias"]) if spec.ln else None
    res = lib.dense_fwd(segs, P[spec.lin + ".weight"], P[spec.lin + ".bias"], ln, spec.act, save_ln=save and ln is not None)
    res["segs"] = segs if save else None
    return res


def dense_backward(P, G: Optional[Dict[str, Tensor]], spec: DenseSpec, saved, gout: Tensor,
                   in_cols: Optional[Sequence[Tuple[int, int]]], accumulate: bool):
    """Returns (gz, [input-gradient per requested column window]).  Parameter gradients go to G."""
    W = P[spec.lin + ".weight"]
    if spec.ln:
        gz, dg, db = lib.ln_act_bwd(gout, saved["out"], spec.act, saved["xhat"], saved["rstd"], P[spec.ln + ".weight"],
                                    None if G is None else G[spec.ln + ".weight"],
                                    None if G is None else G[spec.ln + ".bias"], accumulate)
    elif spec.act != ACT_NONE:
        gz = lib.ln_act_bwd(gout, saved["out"], spec.act)[0]
    else:
        gz = gout
    if G is not None:
        _wgrad_with_bias(gz, saved["segs"], G[spec.lin + ".weight"], G[spec.lin + ".bias"], accumulate)
    gins = []
    if in_cols is not None:
        for lo, hi in in_cols:
            gins.append(lib.dense_fwd([gz], W, transposed=True, cols=(lo, hi))["out"])
    return gz, gins


def _wgrad_with_bias(gz: Tensor, segs, dW: Tensor, db: Tensor, accumulate: bool) -> None:
    """dW = gz^T X and db = gz^T 1 in one launch: the ones segment is appended as the last K column and
    routed to db by the fold kernel."""
    lib.dense_wgrad(gz, list(segs) + [None], dW=dW, accumulate=accumulate, dbias=db)


def conv_forward(P, spec: ConvSpec, csr, x: Tensor, keep, save: bool):
    """keep: None (eval), an explicit uint8 mask, or ("philox", seed, offset) for in-kernel masks."""
    a_s, a_d = P[spec.conv + ".att_src"].view(-1), P[spec.conv + ".att_dst"].view(-1)
    lin = lib.dense_fwd([x], P[spec.conv + ".lin.weight"], att=(a_s, a_d))
    o, m, z = lib.gat_fwd(csr, lin["out"], lin["s"], lin["d"], P[spec.conv + ".bias"])
    gnp = (P[spec.norm + ".weight"], P[spec.norm + ".bias"], P[spec.norm + ".mean_scale"])
    if keep is None:
        x1, stats = lib.graphnorm_fwd(o, *gnp, None, 1.0)
    elif isinstance(keep, tuple):
        x1, stats = lib.graphnorm_fwd(o, *gnp, None, KEEP_P, keep[1], keep[2])
    else:
        x1, stats = lib.graphnorm_fwd(o, *gnp, keep, KEEP_P)
    sv = None
    if save:
        sv = dict(x=x, h=lin["out"], s=lin["s"], d=lin["d"], o=o, m=m, z=z, x1=x1, stats=stats,
                  scale=KEEP_SCALE if keep is not None else 1.0)
    return x1, sv


def conv_backward(P, G, spec: ConvSpec, csr, sv, gx1: Optional[Tensor], accumulate: bool, inject=None, keep_for_bwd2=False):
    """First-order backward of one [GATConv, GraphNorm, ReLU, Dropout] block.  ``inject`` = (ot, ht)
    cotangents added at o and h (second-order sweep).  Returns gx (gradient at the block input)."""
    a_s, a_d = P[spec.conv + ".att_src"].view(-1), P[spec.conv + ".att_dst"].view(-1)
    nrm = spec.norm
    if gx1 is not None:
        dpar = None if G is None else G["__gn__" + nrm]
        go, dpar, bstats = lib.graphnorm_bwd(gx1, sv["o"], sv["x1"], P[nrm + ".weight"], P[nrm + ".mean_scale"], sv["stats"],
                                             sv["scale"], dpar, accumulate)
        if inject is not None:
            lib.axpy_(go, inject[0])
    else:
        go, bstats = inject[0], None
    gh, gsd, Pe, DU = lib.gat_bwd(csr, go, sv["h"], sv["s"], sv["d"], sv["m"], sv["z"], a_s, a_d)
    if G is not None:
        lib.dense_wgrad(go, [None], dW=G[spec.conv + ".bias"].view(-1, 1), accumulate=accumulate)
        lib.dense_wgrad(gsd, [sv["h"]], dW=G["__att__" + spec.conv], accumulate=accumulate)
    if inject is not None:
        lib.axpy_(gh, inject[1])
    if G is not None:
        lib.dense_wgrad(gh, [sv["x"]], dW=G[spec.conv + ".lin.weight"], accumulate=accumulate)
    gx = lib.dense_fwd([gh], P[spec.conv + ".lin.weight"], transposed=True)["out"]
    if keep_for_bwd2:
        sv["b_gx1"], sv["b_go"], sv["b_gh"], sv["b_gsd"], sv["b_bstats"] = gx1, go, gh, gsd, bstats
    return gx


def conv_backward2(P, G2, spec: ConvSpec, csr, sv, Xt: Tensor):
    """Second-order step through (lin-bwd, gat-bwd, gn-bwd) of one block, in that order.  Xt = cotangent
    on the block's input gradient gx.  Returns (cotangent on gx1, (ot, ht) to inject at o / h)."""
    a_s, a_d = P[spec.conv + ".att_src"].view(-1), P[spec.conv + ".att_dst"].view(-1)
    W = P[spec.conv + ".lin.weight"]
    lin = lib.dense_fwd([Xt], W, att=(a_s, a_d))                       # Ht = Xt W^T, St, Dt
    lib.dense_wgrad(sv["b_gh"], [Xt], dW=G2[spec.conv + ".lin.weight"], accumulate=True)
    lib.dense_wgrad(sv["b_gsd"], [lin["out"]], dW=G2["__att__" + spec.conv], accumulate=True)
    gt, ht, sdt = lib.gat_bwd2(csr, lin["out"], lin["s"], lin["d"], sv["b_go"], sv["h"], sv["s"], sv["d"], sv["m"], sv["z"],
                               a_s, a_d)
    lib.dense_wgrad(sdt, [sv["h"]], dW=G2["__att__" + spec.conv], accumulate=True)
    nrm = spec.norm
    gx1t, ot, _ = lib.graphnorm_bwd2(gt, sv["b_gx1"], sv["o"], sv["x1"], P[nrm + ".weight"], P[nrm + ".mean_scale"],
                                     sv["stats"], sv["b_bstats"], sv["scale"], G2["__gn__" + nrm], True)
    return gx1t, (ot, ht)


# ------------------------------------------------------------------------------------------------
# grad-bucket helpers: GraphNorm's three vectors are produced as one [3,C] block and the two
# attention vectors as one [2,C] block; the bucket holds them contiguously under alias keys.
# ------------------------------------------------------------------------------------------------
def grad_views(layout: ParamLayout, flat: Tensor, convs: Sequence[ConvSpec]) -> Dict[str, Tensor]:
    G = {name: layout.view(flat, name) for name in layout.names}
    for c in convs:
        n = c.cout
        o = layout.offsets[c.norm + ".weight"]
        assert layout.offsets[c.norm + ".bias"] == o + n and layout.offsets[c.norm + ".mean_scale"] == o + 2 * n
        G["__gn__" + c.norm] = flat[o: o + 3 * n].view(3, n)
        o = layout.offsets[c.conv + ".att_src"]
        assert layout.offsets[c.conv + ".att_dst"] == o + n
        G["__att__" + c.conv] = flat[o: o + 2 * n].view(2, n)
    return G


def conv_groups(convs: Sequence[ConvSpec]) -> List[List[str]]:
    out = []
    for c in convs:
        out.append([c.norm + ".weight", c.norm + ".bias", c.norm + ".mean_scale"])
        out.append([c.conv + ".att_src", c.conv + ".att_dst"])
    return out
