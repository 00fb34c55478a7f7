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
    """One [conv, GraphNorm, ReLU, Dropout] block.  ``kind`` in GATCONV / GCNCONV / GRAPHCONV / GATV2CONV
    (reference models.py:22-31,166-175)."""

    def __init__(self, conv: str, norm: str, cin: int, cout: int, kind: str = "GATCONV"):
        self.conv, self.norm, self.cin, self.cout, self.kind = conv, norm, cin, cout, kind


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


def _colsum(rows: Tensor, dst: Tensor, accumulate: bool) -> None:
    """dst[C] (+)= column sums of rows[N,C] (deterministic: the weight-gradient kernel against a column of ones)."""
    lib.dense_wgrad(rows, [None], dW=dst.view(-1, 1), accumulate=accumulate)


def _gcn_weights(csr) -> Tensor:
    w = getattr(csr, "_gcn_w", None)
    if w is None or w.device != csr.device:
        w = lib.gcn_norm(csr)
        csr._gcn_w = w
    return w


# ---- aggregation part of a block, per conv type --------------------------------------------------
# forward:   (o, saved)                                       o = conv(x) incl. its bias
# backward:  gx from go (= d loss / d o); ``inj`` = cotangent(s) injected at the conv's intermediate (second-order
#            sweep); parameter gradients into G; with keep_for_bwd2 the intermediates are kept in ``sv``
# backward2: (gt, inj) from Xt (= cotangent on gx): gt = cotangent on go, inj = what the injected first-order sweep
#            adds at the intermediate; direct parameter cotangents are accumulated into G2
def _agg_forward(P, spec: ConvSpec, csr, x: Tensor):
    cv = spec.conv
    if spec.kind == "GATCONV":
        a_s, a_d = P[cv + ".att_src"].view(-1), P[cv + ".att_dst"].view(-1)
        lin = lib.dense_fwd([x], P[cv + ".lin.weight"], att=(a_s, a_d))
        o, m, z = lib.gat_fwd(csr, lin["out"], lin["s"], lin["d"], P[cv + ".bias"])
        return o, dict(h=lin["out"], s=lin["s"], d=lin["d"], m=m, z=z)
    if spec.kind == "GCNCONV":
        h = lib.dense_fwd([x], P[cv + ".lin.weight"])["out"]
        return lib.spmm(csr, h, _gcn_weights(csr), P[cv + ".bias"]), dict(h=h)
    if spec.kind == "GRAPHCONV":
        if getattr(csr, "num_input_self_loops", 0):
            raise RuntimeError("GRAPHCONV: the voxel graph has self loops in edge_index; the CSR strips them and GraphConv "
                               "(which adds none) would lose their contribution")
        agg = lib.spmm(csr, x, None, None, self_loops=False)
        o = lib.dense_fwd([agg], P[cv + ".lin_rel.weight"], P[cv + ".lin_rel.bias"])["out"]
        lib.axpy_(o, lib.dense_fwd([x], P[cv + ".lin_root.weight"])["out"])
        return o, dict(agg=agg)
    if spec.kind == "GATV2CONV":
        xl = lib.dense_fwd([x], P[cv + ".lin_l.weight"], P[cv + ".lin_l.bias"])["out"]
        xr = lib.dense_fwd([x], P[cv + ".lin_r.weight"], P[cv + ".lin_r.bias"])["out"]
        o, logit, m, z = lib.gatv2_fwd(csr, xl, xr, P[cv + ".att"].view(-1), P[cv + ".bias"])
        return o, dict(xl=xl, xr=xr, logit=logit, m=m, z=z)
    raise ValueError(f"Invalid conv_type: {spec.kind}")


def _agg_backward(P, G, spec: ConvSpec, csr, sv, go: Tensor, accumulate: bool, inj, keep_for_bwd2: bool) -> Tensor:
    cv, x = spec.conv, sv["x"]
    if spec.kind == "GATCONV":
        a_s, a_d = P[cv + ".att_src"].view(-1), P[cv + ".att_dst"].view(-1)
        gh, gsd, Pe, DU = lib.gat_bwd(csr, go, sv["h"], sv["s"], sv["d"], sv["m"], sv["z"], a_s, a_d)
        if G is not None:
            _colsum(go, G[cv + ".bias"], accumulate)
            lib.dense_wgrad(gsd, [sv["h"]], dW=G["__att__" + cv], accumulate=accumulate)
        if inj is not None:
            lib.axpy_(gh, inj)
        if G is not None:
            lib.dense_wgrad(gh, [x], dW=G[cv + ".lin.weight"], accumulate=accumulate)
        if keep_for_bwd2:
            sv["b_gh"], sv["b_gsd"] = gh, gsd
        return lib.dense_fwd([gh], P[cv + ".lin.weight"], transposed=True)["out"]
    if spec.kind == "GCNCONV":
        gh = lib.spmm(csr, go, _gcn_weights(csr), None, transpose=True)
        if inj is not None:
            lib.axpy_(gh, inj)
        if G is not None:
            _colsum(go, G[cv + ".bias"], accumulate)
            lib.dense_wgrad(gh, [x], dW=G[cv + ".lin.weight"], accumulate=accumulate)
        if keep_for_bwd2:
            sv["b_gh"] = gh
        return lib.dense_fwd([gh], P[cv + ".lin.weight"], transposed=True)["out"]
    if spec.kind == "GRAPHCONV":
        g_agg = lib.dense_fwd([go], P[cv + ".lin_rel.weight"], transposed=True)["out"]
        gx = lib.dense_fwd([go], P[cv + ".lin_root.weight"], transposed=True)["out"]
        lib.axpy_(gx, lib.spmm(csr, g_agg, None, None, transpose=True, self_loops=False))
        if G is not None:
            _wgrad_with_bias(go, [sv["agg"]], G[cv + ".lin_rel.weight"], G[cv + ".lin_rel.bias"], accumulate)
            lib.dense_wgrad(go, [x], dW=G[cv + ".lin_root.weight"], accumulate=accumulate)
        return gx
    if spec.kind == "GATV2CONV":
        att = P[cv + ".att"].view(-1)
        gxl, gxr, garow = lib.gatv2_bwd(csr, go, sv["xl"], sv["xr"], att, sv["logit"], sv["m"], sv["z"])
        if inj is not None:
            lib.axpy_(gxl, inj[0])
            lib.axpy_(gxr, inj[1])
        if G is not None:
            _colsum(go, G[cv + ".bias"], accumulate)
            _colsum(garow, G[cv + ".att"], accumulate)
            _wgrad_with_bias(gxl, [x], G[cv + ".lin_l.weight"], G[cv + ".lin_l.bias"], accumulate)
            _wgrad_with_bias(gxr, [x], G[cv + ".lin_r.weight"], G[cv + ".lin_r.bias"], accumulate)
        if keep_for_bwd2:
            sv["b_gxl"], sv["b_gxr"] = gxl, gxr
        gx = lib.dense_fwd([gxl], P[cv + ".lin_l.weight"], transposed=True)["out"]
        return lib.axpy_(gx, lib.dense_fwd([gxr], P[cv + ".lin_r.weight"], transposed=True)["out"])
    raise ValueError(f"Invalid conv_type: {spec.kind}")


def _agg_backward2(P, G2, spec: ConvSpec, csr, sv, Xt: Tensor):
    cv = spec.conv
    if spec.kind == "GATCONV":
        a_s, a_d = P[cv + ".att_src"].view(-1), P[cv + ".att_dst"].view(-1)
        lin = lib.dense_fwd([Xt], P[cv + ".lin.weight"], att=(a_s, a_d))                       # Ht = Xt W^T, St, Dt
        lib.dense_wgrad(sv["b_gh"], [Xt], dW=G2[cv + ".lin.weight"], accumulate=True)
        lib.dense_wgrad(sv["b_gsd"], [lin["out"]], dW=G2["__att__" + cv], accumulate=True)
        gt, ht, sdt = lib.gat_bwd2(csr, lin["out"], lin["s"], lin["d"], sv["b_go"], sv["h"], sv["s"], sv["d"], sv["m"], sv["z"],
                                   a_s, a_d)
        lib.dense_wgrad(sdt, [sv["h"]], dW=G2["__att__" + cv], accumulate=True)
        return gt, ht
    if spec.kind == "GCNCONV":  # gx = gh W, gh = A^T go: linear in go, no dependence on the activations
        Ht = lib.dense_fwd([Xt], P[cv + ".lin.weight"])["out"]
        lib.dense_wgrad(sv["b_gh"], [Xt], dW=G2[cv + ".lin.weight"], accumulate=True)
        return lib.spmm(csr, Ht, _gcn_weights(csr), None), None
    if spec.kind == "GRAPHCONV":  # gx = go W_root + S^T (go W_rel)
        SXt = lib.spmm(csr, Xt, None, None, self_loops=False)
        gt = lib.dense_fwd([Xt], P[cv + ".lin_root.weight"])["out"]
        lib.axpy_(gt, lib.dense_fwd([SXt], P[cv + ".lin_rel.weight"])["out"])
        lib.dense_wgrad(sv["b_go"], [Xt], dW=G2[cv + ".lin_root.weight"], accumulate=True)
        lib.dense_wgrad(sv["b_go"], [SXt], dW=G2[cv + ".lin_rel.weight"], accumulate=True)
        return gt, None
    if spec.kind == "GATV2CONV":
        att = P[cv + ".att"].view(-1)
        Hl = lib.dense_fwd([Xt], P[cv + ".lin_l.weight"])["out"]
        Hr = lib.dense_fwd([Xt], P[cv + ".lin_r.weight"])["out"]
        lib.dense_wgrad(sv["b_gxl"], [Xt], dW=G2[cv + ".lin_l.weight"], accumulate=True)
        lib.dense_wgrad(sv["b_gxr"], [Xt], dW=G2[cv + ".lin_r.weight"], accumulate=True)
        gt, cxl, cxr, carow = lib.gatv2_bwd2(csr, Hl, Hr, sv["b_go"], sv["xl"], sv["xr"], att, sv["logit"], sv["m"], sv["z"])
        _colsum(carow, G2[cv + ".att"], True)
        return gt, (cxl, cxr)
    raise ValueError(f"Invalid conv_type: {spec.kind}")


def conv_forward(P, spec: ConvSpec, csr, x: Tensor, keep, save: bool):
    """keep: None (eval), an explicit uint8 mask, or ("philox", seed, offset) for in-kernel masks."""
    o, agg = _agg_forward(P, spec, csr, x)
    gnp = (P[spec.norm + ".weight"], P[spec.norm + ".bias"], P[spec.norm + ".mean_scale"])
    if keep is None:
        x1, stats = lib.graphnorm_fwd(o, *gnp, None, 1.0)
    elif isinstance(keep, tuple):
        x1, stats = lib.graphnorm_fwd(o, *gnp, None, KEEP_P, keep[1], keep[2])
    else:
        x1, stats = lib.graphnorm_fwd(o, *gnp, keep, KEEP_P)
    sv = None
    if save:
        sv = dict(x=x, o=o, x1=x1, stats=stats, scale=KEEP_SCALE if keep is not None else 1.0, **agg)
    return x1, sv


def conv_backward(P, G, spec: ConvSpec, csr, sv, gx1: Optional[Tensor], accumulate: bool, inject=None, keep_for_bwd2=False):
    """First-order backward of one [conv, GraphNorm, ReLU, Dropout] block.  ``inject`` = (ot, it) cotangents added at o
    and at the conv's intermediate (second-order sweep).  Returns gx (gradient at the block input)."""
    nrm = spec.norm
    if gx1 is not None:
        dpar = None if G is None else G["__gn__" + nrm]
        go, dpar, bstats = lib.graphnorm_bwd(gx1, sv["o"], sv["x1"], P[nrm + ".weight"], P[nrm + ".mean_scale"], sv["stats"],
                                             sv["scale"], dpar, accumulate)
        if inject is not None:
            lib.axpy_(go, inject[0])
    else:
        go, bstats = inject[0], None
    gx = _agg_backward(P, G, spec, csr, sv, go, accumulate, None if inject is None else inject[1], keep_for_bwd2)
    if keep_for_bwd2:
        sv["b_gx1"], sv["b_go"], sv["b_bstats"] = gx1, go, bstats
    return gx


def conv_backward2(P, G2, spec: ConvSpec, csr, sv, Xt: Tensor):
    """Second-order step through (conv-bwd, gn-bwd) of one block, in that order.  Xt = cotangent on the block's input
    gradient gx.  Returns (cotangent on gx1, (ot, it) to inject at o / the conv's intermediate)."""
    gt, it = _agg_backward2(P, G2, spec, csr, sv, Xt)
    nrm = spec.norm
    gx1t, ot, _ = lib.graphnorm_bwd2(gt, sv["b_gx1"], sv["o"], sv["x1"], P[nrm + ".weight"], P[nrm + ".mean_scale"],
                                     sv["stats"], sv["b_bstats"], sv["scale"], G2["__gn__" + nrm], True)
    return gx1t, (ot, it)


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
        if c.kind == "GATCONV":
            o = layout.offsets[c.conv + ".att_src"]
            assert layout.offsets[c.conv + ".att_dst"] == o + n
            G["__att__" + c.conv] = flat[o: o + 2 * n].view(2, n)
    return G


def conv_groups(convs: Sequence[ConvSpec]) -> List[List[str]]:
    out = []
    for c in convs:
        out.append([c.norm + ".weight", c.norm + ".bias", c.norm + ".mean_scale"])
        if c.kind == "GATCONV":
            out.append([c.conv + ".att_src", c.conv + ".att_dst"])
    return out
