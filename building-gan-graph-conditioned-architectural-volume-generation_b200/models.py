"""Drop-in ``VoxelGNNGenerator`` / ``VoxelGNNDiscriminator`` (reference building_gan/src/models.py:14-245)
on the libbgb200 kernels.

Same constructor and ``forward`` signatures, same sub-module names / parameter shapes (state-dict
interchange), same initialisation order, same train/eval semantics (dropout masks
and Gumbel noise are drawn from torch's device generator, z / the GP mixing factor come from the
caller), and full autograd support: first order for the generator, first AND second order for the
discriminator (``torch.autograd.grad(..., create_graph=True)`` of WGAN-GP, trainer.py:306-312).
Everything between the inputs and the outputs runs in hand-written sm_100a kernels; there is no
torch / CPU fallback - on a machine without the built library or without CUDA ``forward`` raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import os
import weakref
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import executor as ex
from . import lib
from .executor import ConvSpec, DenseSpec
from .graph import csr_of
from .lib import ACT_LRELU, ACT_NONE, ACT_RELU


# ------------------------------------------------------------------------------------------------
# parameter holders with torch_geometric's names and initialisers (never called as modules)
# ------------------------------------------------------------------------------------------------
def _glorot(t: Tensor) -> None:
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-bound, bound)


def _kaiming(t: Tensor) -> None:
    """torch_geometric.nn.inits.kaiming_uniform(value, fan=in_channels, a=sqrt(5)): U(+-1/sqrt(fan_in)), one draw."""
    bound = 1.0 / math.sqrt(t.size(-1))
    with torch.no_grad():
        t.uniform_(-bound, bound)


def _uniform_fan(bias: Tensor, fan: int) -> None:
    """torch_geometric.nn.inits.uniform(size=in_channels, bias): U(+-1/sqrt(in_channels))."""
    bound = 1.0 / math.sqrt(fan)
    with torch.no_grad():
        bias.uniform_(-bound, bound)


class _PygLinear(nn.Module):
    """Parameters of torch_geometric.nn.dense.Linear(in, out, bias, weight_initializer): ``weight`` [out, in] drawn by
    ``glorot`` or (weight_initializer=None) kaiming-uniform(a=sqrt 5); ``bias`` [out] drawn U(+-1/sqrt(in)) (PyG's
    ``bias_initializer=None`` branch - NOT zeros).  ``reset_parameters()`` consumes the RNG exactly like PyG's: the convs
    call it a second time from their own ``reset_parameters()`` (PyG 2.6.1 ``Linear.__init__`` already drew once)."""

    def __init__(self, cin: int, cout: int, bias: bool = False, init: str = "glorot"):
        super().__init__()
        self._init = init
        self.weight = nn.Parameter(torch.empty(cout, cin))
        if bias:
            self.bias = nn.Parameter(torch.empty(cout))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        (_glorot if self._init == "glorot" else _kaiming)(self.weight)
        if self.bias is not None:
            _uniform_fan(self.bias, self.weight.size(1))


class GATConv(nn.Module):
    """Parameters of tgnn.GATConv(in, out) (heads=1): att_src[1,1,C], att_dst[1,1,C], bias[C], lin.weight."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        _check_widths("GATConv", in_channels, out_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _PygLinear(in_channels, out_channels)
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin.reset_parameters()  # PyG's reset_parameters() re-draws lin before the attention vectors
        _glorot(self.att_src)
        _glorot(self.att_dst)


def _check_widths(name: str, in_channels: int, out_channels: int) -> None:
    if out_channels not in lib.SUPPORTED_WIDTHS or in_channels > 128:
        raise ValueError(f"{name}({in_channels}, {out_channels}): libbgb200 supports widths {lib.SUPPORTED_WIDTHS}")


class GCNConv(nn.Module):
    """Parameters of tgnn.GCNConv(in, out): bias[C] (zeros), lin.weight[C, Cin] (glorot, no bias; drawn by Linear.__init__
    and again by GCNConv.reset_parameters())."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        _check_widths("GCNConv", in_channels, out_channels)
        self.lin = _PygLinear(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin.reset_parameters()


class GraphConv(nn.Module):
    """Parameters of tgnn.GraphConv(in, out): lin_rel.{weight,bias}, lin_root.weight - PyG Linears with the default
    initialisers (kaiming-uniform weight, U(+-1/sqrt(in)) bias), each drawn at construction and again by
    GraphConv.reset_parameters()."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        _check_widths("GraphConv", in_channels, out_channels)
        self.lin_rel = _PygLinear(in_channels, out_channels, bias=True, init="kaiming")
        self.lin_root = _PygLinear(in_channels, out_channels, init="kaiming")
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()


class GATv2Conv(nn.Module):
    """Parameters of tgnn.GATv2Conv(in, out) (heads=1, share_weights=False): att[1,1,C], bias[C] (zeros),
    lin_l.{weight,bias}, lin_r.{weight,bias} (glorot weights, U(+-1/sqrt(in)) biases; drawn twice like PyG)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        _check_widths("GATv2Conv", in_channels, out_channels)
        self.lin_l = _PygLinear(in_channels, out_channels, bias=True)
        self.lin_r = _PygLinear(in_channels, out_channels, bias=True)
        self.att = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        _glorot(self.att)


class GraphNorm(nn.Module):
    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.mean_scale = nn.Parameter(torch.ones(channels))


_CONVS = {"GCNCONV": GCNConv, "GRAPHCONV": GraphConv, "GATCONV": GATConv, "GATV2CONV": GATv2Conv}


class _GnnStack(nn.Module):
    """Registers children as ``module_{i}`` exactly like tgnn.Sequential (models.py:90,210)."""

    def __init__(self, widths: List[int], kind: str = "GATCONV"):
        super().__init__()
        self.specs: List[ConvSpec] = []
        conv = _CONVS[kind]
        i = 0
        for a, b in zip(widths[:-1], widths[1:]):
            setattr(self, f"module_{i}", conv(a, b))
            setattr(self, f"module_{i + 1}", GraphNorm(b))
            setattr(self, f"module_{i + 2}", nn.ReLU(True))
            setattr(self, f"module_{i + 3}", nn.Dropout(0.2))
            self.specs.append(ConvSpec(f"module_{i}", f"module_{i + 1}", a, b, kind))
            i += 4


def _ln_mlp(widths: List[int], final_plain: Optional[int] = None) -> nn.Sequential:
    mods: List[nn.Module] = []
    for a, b in zip(widths[:-1], widths[1:]):
        mods += [nn.Linear(a, b), nn.LayerNorm(b), nn.LeakyReLU(0.2)]
    if final_plain is not None:
        mods.append(nn.Linear(widths[-1], final_plain))
    return nn.Sequential(*mods)


def _hourglass(hidden: int, repeat: int) -> List[int]:
    down = [hidden // (2 ** k) for k in range(repeat + 1)]
    return down + down[-2::-1]


def _param_list(model) -> List[nn.Parameter]:
    """``list(model.parameters())`` cached on the model: walking the module tree costs ~0.1 ms and every pass of the
    training step (46 per step) needs the list.  nn.Module keeps the Parameter objects across .to()/.cuda()/load_state_dict
    (their .data is swapped in place); re-registering a parameter by hand must be followed by ``model._bg_params = None``."""
    cached = model.__dict__.get("_bg_params")
    if cached is None:
        cached = list(model.parameters())
        model.__dict__["_bg_params"] = cached
    return cached


_LIVE = weakref.WeakSet()


def live_models():
    """The generator / discriminator instances alive in this process (optim.Adam finds the owner of a parameter list here)."""
    return list(_LIVE)


def _executor_for(kind: str) -> str:
    return EXECUTOR if kind == "GATCONV" else "python"


def _conv_kind(kind: str) -> str:
    if kind not in _CONVS:
        raise ValueError(f"Invalid conv_type: {kind}")  # models.py:31,175
    return kind


# "native" (default): one C call per pass (csrc/bg_passes.cu).  "python": the same passes composed op by op in
# executor.py (kept as the readable specification of the pass structure and for debugging).  The native pass executor
# covers the reference's default conv type (GATCONV, config.py:89,93); models built with GCNCONV / GRAPHCONV / GATV2CONV
# run the op-by-op executor on the same kernel library (per-model ``_executor``).
EXECUTOR = os.environ.get("BG_EXECUTOR", "native")
# "philox" (default): dropout masks / Gumbel noise generated inside the kernels (Philox4x32-10, seeded from
# torch.cuda.initial_seed()).  "torch": drawn with torch's device generator in the reference's draw order
#  - bit-identical RNG stream to the reference running on the same GPU, ~200 extra launches
# per training step.
RNG_MODE = os.environ.get("BG_RNG", "philox")
# "bucket" (default): the backward passes accumulate parameter gradients IN THE KERNELS, straight into ``p.grad`` - every
# ``p.grad`` of a model is a view of one flat fp32 bucket (also the NCCL all-reduce payload).  This is what ``loss.backward()``
# + ``optimizer.step()`` (the only pattern trainer.py uses) needs, and it removes ~900 tiny torch add kernels per training
# step.  A backward that runs under ``create_graph=True`` (the WGAN-GP ``torch.autograd.grad(..., inputs=interpolated)``,
# trainer.py:306-312) computes input gradients only, exactly like autograd's ``only_inputs``.  NOT supported in this mode:
# ``torch.autograd.grad`` w.r.t. parameters and ``backward(create_graph=True)``.
# "autograd": parameter gradients are returned to autograd like any op (fully general, slower).
GRAD_MODE = os.environ.get("BG_GRADS", "bucket")
_philox_calls = 0


def _philox_ticket():
    """(seed, offset) for one model call: offsets are disjoint per call (a call uses offset .. offset+1000)."""
    global _philox_calls
    _philox_calls += 1
    return torch.cuda.initial_seed() & 0xFFFFFFFFFFFFFFFF, _philox_calls * 2048


def _model_desc(c, local_dim: int, voxel_dim: int) -> "lib.BgModelDesc":
    md = lib.BgModelDesc()
    md.local_dim, md.voxel_dim, md.num_classes, md.z_dim = local_dim, voxel_dim, c.NUM_CLASSES, c.Z_DIM
    md.le_dim, md.le_layers = c.LOCAL_ENCODER_HIDDEN_DIM, c.LOCAL_GRAPH_ENCODER_REPEAT + 1
    md.g_hidden, md.g_mlp_layers, md.g_repeat = c.GENERATOR_HIDDEN_DIM, c.GENERATOR_MLP_ENCODER_REPEAT + 1, c.GENERATOR_ENCODER_REPEAT
    md.d_hidden, md.d_repeat = c.DISCRIMINATOR_HIDDEN_DIM, c.DISCRIMINATOR_ENCODER_REPEAT
    return md


class _NativeState:
    """Per-model host-side constants of the native executor: model descriptor, grad-bucket offsets, cached
    parameter-pointer table (invalidated by nn.Module._apply, i.e. .to()/.cuda()/.float())."""

    def __init__(self, model, md, layout):
        self.md, self.layout = md, layout
        self.goff = (C.c_int64 * len(layout.names))(*[layout.offsets[n] for n in layout.names])
        self._ptrs, self._key = None, None
        self.bucket, self.views, self.anchor = None, None, None
        self.pflat = None
        self.lane, self.lane_buckets, self.dirty = None, {}, []

    # -- lanes: passes of ONE model running concurrently on different CUDA streams (step.py, critic update) ---------------
    # The backward kernels add parameter gradients into the bucket with plain read-modify-write stores, so two passes in
    # flight at once need separate destinations: a pass whose forward ran inside ``with model.lane(k)`` delivers its parameter
    # gradients into the private bucket of lane k (overwrite mode: no zeroing needed), and ``merge_lanes()`` - called by the
    # owner of the streams once they are joined - adds the dirty lane buckets into the real one in a fixed order.
    def lane_bucket(self, lane, device):
        """(flat, accumulate) for a backward pass that ran in ``lane``."""
        b = self.lane_buckets.get(lane)
        if b is None or b.device != device:
            b = torch.zeros(self.layout.total, dtype=torch.float32, device=device)
            self.lane_buckets[lane] = b
        again = lane in self.dirty
        if not again:
            self.dirty.append(lane)
        return b, again

    def merge_lanes(self, params) -> None:
        if not self.dirty:
            return
        bucket = self.bind_grads(params)
        for lane in sorted(self.dirty):
            lib.axpy_(bucket, self.lane_buckets[lane])
        self.dirty = []

    def flat_params(self, params) -> Tensor:
        """Re-home every ``p.data`` as a view of ONE flat fp32 buffer laid out like the gradient bucket (optim.Adam steps the
        flat buffers with one launch).  Idempotent; redone if .to()/.cuda()/load_state_dict(assign=True) replaced the storage."""
        lay, pf = self.layout, self.pflat
        first, last = lay.names[0], lay.names[-1]
        if pf is not None and pf.device == params[0].device and params[0].data_ptr() == pf.data_ptr() + 4 * lay.offsets[first] \
                and params[-1].data_ptr() == pf.data_ptr() + 4 * lay.offsets[last]:
            return pf
        pf = torch.zeros(lay.total, dtype=torch.float32, device=params[0].device)
        with torch.no_grad():
            for p, n in zip(params, lay.names):
                v = lay.view(pf, n)
                v.copy_(p.data)
                p.data = v
        self.pflat = pf
        return pf

    def get_anchor(self, device) -> Tensor:
        """1-element leaf that carries the autograd edge of a bucket-mode pass (parameters are not autograd inputs)."""
        if self.anchor is None or self.anchor.device != device:
            self.anchor = torch.zeros(1, device=device, requires_grad=True)
        return self.anchor

    def install_bucket(self, bucket: Tensor, params) -> Tensor:
        """Use an externally allocated flat buffer (symmetric memory: dist.PeerSync) as the gradient bucket."""
        assert bucket.numel() == self.layout.total and bucket.dtype == torch.float32
        self.bucket = bucket
        self.views = [self.layout.view(bucket, n) for n in self.layout.names]
        for p, v in zip(params, self.views):
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
        return bucket

    def bind_grads(self, params) -> Tensor:
        """Make every ``p.grad`` a view of the flat bucket (zeroing what ``zero_grad(set_to_none=True)`` dropped)."""
        dev = params[0].device
        if self.bucket is None or self.bucket.device != dev:
            self.bucket = torch.zeros(self.layout.total, dtype=torch.float32, device=dev)
            self.views = [self.layout.view(self.bucket, n) for n in self.layout.names]
            for p, v in zip(params, self.views):
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
            return self.bucket
        missing = [p.grad is None for p in params]
        if all(missing):
            self.bucket.zero_()
            for p, v in zip(params, self.views):
                p.grad = v
        elif any(missing) or params[0].grad.data_ptr() != self.views[0].data_ptr():
            for p, v, m in zip(params, self.views, missing):
                if m:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
                p.grad = v
        return self.bucket

    def ptrs(self, params):
        key = (params[0].data_ptr(), params[-1].data_ptr(), len(params))
        if self._key != key:
            self._ptrs, self._key = lib.ptr_array(params), key
        return self._ptrs


def _batch_in(bc) -> "lib.BgBatchIn":
    b = getattr(bc, "c_in", None)
    if b is None:
        b = lib.BgBatchIn()
        b.table, b.type32, b.vx = bc.table.data_ptr(), bc.type32.data_ptr(), bc.vx.data_ptr()
        bc.c_in = b
    return b


# ------------------------------------------------------------------------------------------------
# per-batch cache of data-only quantities (hoisted out of every forward call)
# ------------------------------------------------------------------------------------------------
class _BatchCtx:
    __slots__ = ("csr", "type32", "vx", "table", "key", "local_ref", "n", "real_onehot", "c_in")


def _tkey(t: Tensor):
    """Identity of a tensor's CONTENT as far as it can be known without reading it: storage address, in-place version
    counter, shape.  A cache entry also keeps a reference to the keyed tensor, so the address cannot be recycled."""
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype)


def _ctx_key(local_graph, voxel_graph):
    return (id(local_graph), _tkey(local_graph.x), _tkey(local_graph.type), _tkey(voxel_graph.x), _tkey(voxel_graph.type))


def _real_onehot(bc, label: Tensor) -> Tensor:
    """Float copy of an integer one-hot label tensor (the real sample is int64, trainer.py:319), cached per batch and keyed by
    the tensor's address, version and shape (a fresh tensor at a recycled address, or an in-place edit, misses)."""
    key = _tkey(label)
    hit = bc.real_onehot
    if hit is None or hit[0] != key:
        hit = (key, label.to(bc.vx.device, torch.float32).contiguous(), label)  # [2]: keeps the keyed tensor alive
        bc.real_onehot = hit
    return hit[1]


def _batch_ctx(local_graph, voxel_graph, num_types: int) -> _BatchCtx:
    ctx = getattr(voxel_graph, "_bg_cache", None) if not hasattr(voxel_graph, "_fields") else voxel_graph._fields.get("_bg_cache")
    dev = voxel_graph.x.device
    if dev.type != "cuda":
        raise RuntimeError("building_gan_b200 models run on CUDA only (no CPU fallback): move the batch with .to('cuda')")
    key = _ctx_key(local_graph, voxel_graph)
    if ctx is None or ctx.key != key or ctx.vx.device != dev:
        ctx = _BatchCtx()
        ctx.csr = csr_of(voxel_graph)
        ctx.type32 = voxel_graph.type.to(torch.int32).contiguous()
        ctx.vx = voxel_graph.x.to(torch.float32).contiguous()
        ctx.table = lib.type_table(local_graph.x.to(torch.float32).contiguous(), local_graph.type, num_types)
        ctx.key, ctx.local_ref = key, local_graph  # the reference keeps id(local_graph) from being recycled
        ctx.n = int(ctx.vx.shape[0])
        ctx.real_onehot = None
        ctx.c_in = None
        try:
            voxel_graph._bg_cache = ctx
        except Exception:
            pass
    return ctx


def prepare_batch(local_graph, voxel_graph, num_types: int) -> _BatchCtx:
    """Build every lazily cached per-batch tensor now (type table, CSR handles, the float copy of the real one-hot labels):
    nothing is allocated-and-cached during the model calls afterwards (required before CUDA-graph capture, graphs.py)."""
    bc = _batch_ctx(local_graph, voxel_graph, num_types)
    label = voxel_graph.types_onehot
    if label.dtype != torch.float32:
        _real_onehot(bc, label)
    _batch_in(bc)
    bc.csr.c_struct()
    return bc


def _draw_keeps(n: int, specs: List[ConvSpec], training: bool, device) -> List[Optional[Tensor]]:
    """Dropout keep-masks, one per block, drawn with torch's own dropout kernel in layer order so the
    device RNG stream is consumed exactly as the reference consumes it."""
    if not training:
        return [None] * len(specs)
    cmax = max(s.cout for s in specs)
    ones = torch.ones(n * cmax, device=device)
    return [torch.native_dropout(ones[: n * s.cout].view(n, s.cout), 0.2, True)[1].view(torch.uint8) for s in specs]


# ------------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------------
class VoxelGNNGenerator(nn.Module):
    def __init__(self, configuration, local_graph_dim: int, voxel_graph_dim: int):
        super().__init__()
        c = configuration
        self.configuration = c
        self.local_graph_dim, self.voxel_graph_dim = local_graph_dim, voxel_graph_dim
        kind = _conv_kind(c.GENERATOR_CONV_TYPE)
        self._kind = kind
        le, gh = c.LOCAL_ENCODER_HIDDEN_DIM, c.GENERATOR_HIDDEN_DIM
        self.matched_features_encoder = _ln_mlp([local_graph_dim] + [le] * (c.LOCAL_GRAPH_ENCODER_REPEAT + 1))
        self.mlp_encoder = _ln_mlp([le + voxel_graph_dim + c.Z_DIM] + [gh] * (c.GENERATOR_MLP_ENCODER_REPEAT + 1))
        widths = _hourglass(gh, c.GENERATOR_ENCODER_REPEAT)
        self.encoder = _GnnStack(widths, kind)
        self.decoder = _ln_mlp([le + voxel_graph_dim + c.Z_DIM + widths[-1] + gh, gh, gh // 2, gh // 4, gh // 8],
                               final_plain=c.NUM_CLASSES)
        self._le, self._gh, self._enc_out = le, gh, widths[-1]
        self._menc = [DenseSpec(f"matched_features_encoder.{3 * i}", f"matched_features_encoder.{3 * i + 1}", ACT_LRELU)
                      for i in range(c.LOCAL_GRAPH_ENCODER_REPEAT + 1)]
        self._mlp = [DenseSpec(f"mlp_encoder.{3 * i}", f"mlp_encoder.{3 * i + 1}", ACT_LRELU)
                     for i in range(c.GENERATOR_MLP_ENCODER_REPEAT + 1)]
        self._convs = [ConvSpec("encoder." + s.conv, "encoder." + s.norm, s.cin, s.cout, s.kind) for s in self.encoder.specs]
        self._dec = [DenseSpec(f"decoder.{3 * i}", f"decoder.{3 * i + 1}", ACT_LRELU) for i in range(4)]
        self._dec.append(DenseSpec("decoder.12", None, ACT_NONE))
        self._names = [n for n, _ in self.named_parameters()]
        self._layout = ex.ParamLayout(list(self.named_parameters()), ex.conv_groups(self._convs))
        self._native = _NativeState(self, _model_desc(c, local_graph_dim, voxel_graph_dim), self._layout)
        assert kind != "GATCONV" or lib.load().bg_gen_num_params(C.byref(self._native.md)) == len(self._names)
        self.to(c.DEVICE)
        _LIVE.add(self)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_bg_params"] = None  # .to() / .cuda() may re-create the Parameter objects
        return super()._apply(fn, *args, **kwargs)

    def forward(self, local_graph, voxel_graph, z, gumbel_noise: Optional[Tensor] = None, keeps=None):
        """(logits, label_hard, label_soft), reference models.py:119-155.  ``gumbel_noise`` [N,7] / ``keeps`` (one
        uint8 [N,C] keep-mask per conv block) inject the random draws explicitly (parity tests); by default they
        are drawn according to RNG_MODE."""
        lib.load()
        bc = _batch_ctx(local_graph, voxel_graph, self.configuration.NUM_CLASSES)
        zz = z.squeeze(0).to(bc.vx.device, torch.float32).contiguous()
        seed, offset = _philox_ticket()
        EXECUTOR = _executor_for(self._kind)
        if keeps is None and self.training and (RNG_MODE == "torch" or EXECUTOR == "python"):
            keeps = _draw_keeps(bc.n, self._convs, True, bc.vx.device)
        if gumbel_noise is None and (RNG_MODE == "torch" or EXECUTOR == "python"):
            gumbel_noise = -torch.empty(bc.n, self.configuration.NUM_CLASSES, device=bc.vx.device).exponential_().log()
        if gumbel_noise is not None:
            gumbel_noise = gumbel_noise.to(torch.float32).contiguous()
        params = _param_list(self)
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        fn = _GenFn if EXECUTOR == "python" else _GenNativeFn
        if EXECUTOR == "python" and keeps is None:
            keeps = [None] * len(self._convs)
        if EXECUTOR != "python" and GRAD_MODE == "bucket":
            logits, hard, soft = fn.apply(self, bc, zz, gumbel_noise, keeps, need, (seed, offset, True),
                                          self._native.get_anchor(zz.device))
        else:
            logits, hard, soft = fn.apply(self, bc, zz, gumbel_noise, keeps, need, (seed, offset, False), *params)
        return logits, hard, soft

    # -- passes -----------------------------------------------------------------------------------
    def _forward_pass(self, P: Dict[str, Tensor], bc: _BatchCtx, zz: Tensor, noise: Tensor, keeps, save: bool):
        sv = {"menc": [], "mlp": [], "conv": [], "dec": []}
        e = bc.table
        for spec in self._menc:                                     # 7-row encoder: row-wise ops commute with the gather
            r = ex.dense_forward(P, spec, [e], save)
            sv["menc"].append(r)
            e = r["out"]
        enc_seg = (e, bc.type32)
        segs = [enc_seg, bc.vx, zz]
        for spec in self._mlp:
            r = ex.dense_forward(P, spec, segs, save)
            sv["mlp"].append(r)
            segs = [r["out"]]
        x = segs[0]
        h = x
        for spec, keep in zip(self._convs, keeps):
            h, s = ex.conv_forward(P, spec, bc.csr, h, keep, save)
            sv["conv"].append(s)
        segs = [h, x, enc_seg, bc.vx, zz]
        for spec in self._dec:
            r = ex.dense_forward(P, spec, segs, save)
            sv["dec"].append(r)
            segs = [r["out"]]
        logits = segs[0]
        soft, hard, amax = lib.gumbel_st_fwd(logits, noise)
        sv["soft"] = soft
        return logits, hard, soft, sv

    def _backward_pass(self, P, bc: _BatchCtx, sv, g_logits, g_hard, g_soft) -> Tensor:
        flat = torch.empty(self._layout.total, dtype=torch.float32, device=bc.vx.device)
        G = ex.grad_views(self._layout, flat, self._convs)
        le, gh, eo = self._le, self._gh, self._enc_out
        if g_hard is not None or g_soft is not None:
            gl = lib.gumbel_st_bwd(g_hard, g_soft, sv["soft"])
            if g_logits is not None:
                lib.axpy_(gl, g_logits)
        else:
            gl = g_logits
        g = gl
        for spec, saved in zip(self._dec[:0:-1], sv["dec"][:0:-1]):
            _, (g,) = ex.dense_backward(P, G, spec, saved, g, [(0, saved["segs"][0].shape[1])], False)
        _, (g_enc_out, g_x_skip, g_e1) = ex.dense_backward(P, G, self._dec[0], sv["dec"][0], g,
                                                           [(0, eo), (eo, eo + gh), (eo + gh, eo + gh + le)], False)
        g = g_enc_out
        for spec, saved in zip(self._convs[::-1], sv["conv"][::-1]):
            g = ex.conv_backward(P, G, spec, bc.csr, saved, g, False)
        lib.axpy_(g, g_x_skip)
        for spec, saved in zip(self._mlp[:0:-1], sv["mlp"][:0:-1]):
            _, (g,) = ex.dense_backward(P, G, spec, saved, g, [(0, gh)], False)
        _, (g_e2,) = ex.dense_backward(P, G, self._mlp[0], sv["mlp"][0], g, [(0, le)], False)
        k = self.configuration.NUM_CLASSES
        ge = lib.type_scatter_sum(g_e1, bc.type32, k)
        lib.axpy_(ge, lib.type_scatter_sum(g_e2, bc.type32, k))
        for i in range(len(self._menc) - 1, -1, -1):
            spec, saved = self._menc[i], sv["menc"][i]
            _, gins = ex.dense_backward(P, G, spec, saved, ge, [(0, saved["segs"][0].shape[1])] if i > 0 else None, False)
            if i > 0:
                ge = gins[0]
        return flat


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: VoxelGNNGenerator, bc, zz, noise, keeps, need, ticket, *params):
        P = dict(zip(model._names, params))
        logits, hard, soft, sv = model._forward_pass(P, bc, zz, noise, keeps, save=need)
        if getattr(model, "debug_keep_saved", False):
            model.debug_saved = sv  # test hook: the activation patterns of the last forward
        ctx.model, ctx.bc, ctx.sv, ctx.P = model, bc, sv if need else None, P if need else None
        return logits, hard, soft

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_logits, g_hard, g_soft):
        if ctx.sv is None:
            raise RuntimeError("generator backward called but the forward ran without grad")
        cont = lambda t: None if t is None else t.contiguous()
        flat = ctx.model._backward_pass(ctx.P, ctx.bc, ctx.sv, cont(g_logits), cont(g_hard), cont(g_soft))
        lay = ctx.model._layout
        return (None, None, None, None, None, None, None) + tuple(lay.view(flat, n) for n in ctx.model._names)


def _saved_views(model, ws: Tensor, n: int, last: Tensor, is_gen: bool):
    """Test hook: the post-activation tensors of the last native forward as views into its workspace, in the
    dict shape the python executor produces ({"menc"/"mlp"/"pre", "conv", "dec"} -> [{"out"| "x1": tensor}])."""
    L, st = lib.load(), model._native
    off = (C.c_int64 * 64)()
    cnt = (L.bg_gen_ws_offsets if is_gen else L.bg_disc_ws_offsets)(C.byref(st.md), n, off, 64)
    offs = list(off[:cnt])
    wsf = ws.view(torch.float32)

    def take(rows, width):
        o = offs.pop(0) // 4
        return wsf[o: o + rows * width].view(rows, width)

    k = model.configuration.NUM_CLASSES
    sv = {}
    if is_gen:
        sv["menc"] = [{"out": take(k, model._le)} for _ in model._menc]
        sv["mlp"] = [{"out": take(n, model._gh)} for _ in model._mlp]
    else:
        dh = model.configuration.DISCRIMINATOR_HIDDEN_DIM
        sv["pre"] = [{"out": take(n, dh)} for _ in model._pre]
    sv["conv"] = [{"x1": take(n, c.cout)} for c in model._convs]
    widths = ([model._gh, model._gh // 2, model._gh // 4, model._gh // 8, k] if is_gen else
              [dh // 2, dh // 4, dh // 8, 1])
    sv["dec"] = [{"out": take(n, w)} for w in widths]
    sv["dec"][-1]["out"] = last
    return sv


def _launches_gen_fwd(model) -> int:
    return len(model._menc) + len(model._mlp) + 4 * len(model._convs) + len(model._dec) + 1


class _GenNativeFn(torch.autograd.Function):
    """Generator forward/backward as two C calls (bg_gen_forward / bg_gen_backward).  ``tensors`` = the parameters
    (GRAD_MODE "autograd") or the 1-element anchor (GRAD_MODE "bucket")."""

    @staticmethod
    def forward(ctx, model: VoxelGNNGenerator, bc, zz, noise, keeps, need, ticket, *tensors):
        L, st = lib.load(), model._native
        params = _param_list(model)
        dev, n, e, k = zz.device, bc.n, bc.csr.num_edges, model.configuration.NUM_CLASSES
        ws = lib.u8_buffer(L.bg_gen_fwd_ws(C.byref(st.md), n, e), dev)
        red = lib.workspace(lib.RED_BYTES, dev)
        logits = torch.empty(n, k, dtype=torch.float32, device=dev)
        hard, soft = torch.empty_like(logits), torch.empty_like(logits)
        kp = None if not keeps or keeps[0] is None else lib.ptr_array(keeps)
        lib._check(L.bg_gen_forward(C.byref(st.md), st.ptrs(params), C.byref(bc.csr.c_struct()), C.byref(_batch_in(bc)),
                                    zz.data_ptr(), lib._p(noise), kp, int(model.training), ticket[0], ticket[1], ws.data_ptr(),
                                    ws.numel(), red.data_ptr(), red.numel() * 4, logits.data_ptr(), hard.data_ptr(),
                                    soft.data_ptr(), lib._stream()))
        lib.pass_launches(_launches_gen_fwd(model))
        if getattr(model, "debug_keep_saved", False):
            model.debug_saved = _saved_views(model, ws, n, logits, True)
        ctx.need, ctx.bucket_mode = need, ticket[2]
        if need:
            ctx.model, ctx.bc, ctx.ws, ctx.zz, ctx.training = model, bc, ws, zz, model.training
            ctx.save_for_backward(logits, soft)
        return logits, hard, soft

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_logits, g_hard, g_soft):
        if not ctx.need:
            raise RuntimeError("generator backward called but the forward ran without grad")
        L, model, bc = lib.load(), ctx.model, ctx.bc
        st = model._native
        params = _param_list(model)
        logits, soft = ctx.saved_tensors
        dev, n, e = logits.device, bc.n, bc.csr.num_edges
        flat = st.bind_grads(params) if ctx.bucket_mode else torch.empty(st.layout.total, dtype=torch.float32, device=dev)
        tmp = lib.u8_buffer(L.bg_gen_bwd_ws(C.byref(st.md), n, e), dev)
        red = lib.workspace(lib.RED_BYTES, dev)
        cont = lambda t: None if t is None else t.contiguous()
        g_logits, g_hard, g_soft = cont(g_logits), cont(g_hard), cont(g_soft)
        lib._check(L.bg_gen_backward(C.byref(st.md), st.ptrs(params), C.byref(bc.csr.c_struct()), C.byref(_batch_in(bc)),
                                     ctx.zz.data_ptr(), ctx.ws.data_ptr(), logits.data_ptr(), soft.data_ptr(), lib._p(g_logits),
                                     lib._p(g_hard), lib._p(g_soft), int(ctx.training), flat.data_ptr(), st.goff,
                                     int(ctx.bucket_mode), tmp.data_ptr(), tmp.numel(), red.data_ptr(), red.numel() * 4,
                                     lib._stream()))
        lib.pass_launches(2 + 3 * (len(model._menc) + len(model._mlp) + len(model._dec)) + 6 * len(model._convs) + 8)
        if ctx.bucket_mode:
            return (None,) * 8
        return (None,) * 7 + tuple(st.layout.view(flat, nm) for nm in model._names)


# ------------------------------------------------------------------------------------------------
# discriminator
# ------------------------------------------------------------------------------------------------
class VoxelGNNDiscriminator(nn.Module):
    def __init__(self, configuration, local_graph_dim: int, voxel_graph_dim: int):
        super().__init__()
        c = configuration
        self.configuration = c
        self.local_graph_dim, self.voxel_graph_dim = local_graph_dim, voxel_graph_dim
        kind = _conv_kind(c.DISCRIMINATOR_CONV_TYPE)
        self._kind = kind
        dh = c.DISCRIMINATOR_HIDDEN_DIM
        self.mlp_encoder = nn.Sequential(nn.Linear(local_graph_dim + voxel_graph_dim + c.NUM_CLASSES, dh), nn.ReLU(True),
                                         nn.Linear(dh, dh), nn.ReLU(True))
        self.encoder = _GnnStack(_hourglass(dh, c.DISCRIMINATOR_ENCODER_REPEAT), kind)
        tail: List[nn.Module] = [nn.Linear(dh, dh // 2), nn.ReLU(True), nn.Linear(dh // 2, dh // 4), nn.ReLU(True),
                                 nn.Linear(dh // 4, dh // 8), nn.ReLU(True), nn.Linear(dh // 8, 1)]
        if not c.USE_WGANGP:
            tail.append(nn.Sigmoid())  # models.py:222-223; applied to the kernel path's score in forward()
        self.decoder = nn.Sequential(*tail)
        self._pre = [DenseSpec("mlp_encoder.0", None, ACT_RELU), DenseSpec("mlp_encoder.2", None, ACT_RELU)]
        self._convs = [ConvSpec("encoder." + s.conv, "encoder." + s.norm, s.cin, s.cout, s.kind) for s in self.encoder.specs]
        self._dec = [DenseSpec("decoder.0", None, ACT_RELU), DenseSpec("decoder.2", None, ACT_RELU),
                     DenseSpec("decoder.4", None, ACT_RELU), DenseSpec("decoder.6", None, ACT_NONE)]
        self._names = [n for n, _ in self.named_parameters()]
        self._layout = ex.ParamLayout(list(self.named_parameters()), ex.conv_groups(self._convs))
        self._label_lo = local_graph_dim + voxel_graph_dim
        self._native = _NativeState(self, _model_desc(c, local_graph_dim, voxel_graph_dim), self._layout)
        assert kind != "GATCONV" or lib.load().bg_disc_num_params(C.byref(self._native.md)) == len(self._names)
        self.to(c.DEVICE)
        _LIVE.add(self)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_bg_params"] = None  # .to() / .cuda() may re-create the Parameter objects
        return super()._apply(fn, *args, **kwargs)

    @contextlib.contextmanager
    def lane(self, key: int):
        """Passes started inside this context may run concurrently (on another CUDA stream) with other passes of this model:
        their backward delivers the parameter gradients into a private bucket (``_NativeState.lane_bucket``) until
        ``merge_lanes()``.  Used by step.discriminator_loss for the D(real) / D(fake) passes of a critic update."""
        prev, self._native.lane = self._native.lane, key
        try:
            yield self
        finally:
            self._native.lane = prev

    @contextlib.contextmanager
    def input_grads_only(self):
        """Backward passes of this model that START inside this context deliver the input gradient only (native executor,
        BG_GRADS=bucket): no weight-gradient launches, ``p.grad`` untouched.  For the generator update (trainer.py:486-491):
        the reference's ``g_loss.backward()`` also fills the critic's ``.grad``, which the next ``optimizer_d.zero_grad()``
        discards unread - the fast path (graphs.GraphedStep) skips producing them."""
        st = getattr(self, "_native", None)
        if st is None or _executor_for(self._kind) != "native":  # op-by-op executor: parameter gradients go through autograd
            yield self
            return
        prev, st.skip_param_grads = getattr(st, "skip_param_grads", False), True
        try:
            yield self
        finally:
            st.skip_param_grads = prev

    def merge_lanes(self) -> None:
        """Add the lane buckets into ``p.grad`` (call on the main stream after the lane streams were joined)."""
        self._native.merge_lanes(_param_list(self))

    def forward(self, local_graph, voxel_graph, label_hard, keeps=None):
        """Per-voxel critic score [N,1], reference models.py:229-245.  ``keeps`` injects explicit dropout masks."""
        lib.load()
        bc = _batch_ctx(local_graph, voxel_graph, self.configuration.NUM_CLASSES)
        label = label_hard.squeeze(0)
        if label.dtype != torch.float32:  # the real sample is an int64 one-hot (trainer.py:319); torch.cat promotes it
            label = _real_onehot(bc, label)
        label = label.to(bc.vx.device).contiguous()
        seed, offset = _philox_ticket()
        EXECUTOR = _executor_for(self._kind)
        if keeps is None and self.training and (RNG_MODE == "torch" or EXECUTOR == "python"):
            keeps = _draw_keeps(bc.n, self._convs, True, bc.vx.device)
        params = _param_list(self)
        need = torch.is_grad_enabled() and (label.requires_grad or any(p.requires_grad for p in params))
        if EXECUTOR == "python":
            score = _DiscFn.apply(self, bc, keeps if keeps is not None else [None] * len(self._convs), need, label, *params)
        elif GRAD_MODE == "bucket":
            score = _DiscNativeFn.apply(self, bc, keeps, need, (seed, offset, True), label,
                                        self._native.get_anchor(label.device))
        else:
            score = _DiscNativeFn.apply(self, bc, keeps, need, (seed, offset, False), label, *params)
        # vanilla-GAN critic (USE_WGANGP=False): the [N,1] score goes through the decoder's final Sigmoid (one tiny
        # elementwise op; its autograd backward feeds g_score of the kernel path)
        return score if self.configuration.USE_WGANGP else torch.sigmoid(score)

    # -- passes -----------------------------------------------------------------------------------
    def _forward_pass(self, P, bc: _BatchCtx, label: Tensor, keeps, save: bool):
        sv = {"pre": [], "conv": [], "dec": []}
        segs = [(bc.table, bc.type32), bc.vx, label]
        for spec in self._pre:
            r = ex.dense_forward(P, spec, segs, save)
            sv["pre"].append(r)
            segs = [r["out"]]
        h = segs[0]
        for spec, keep in zip(self._convs, keeps):
            h, s = ex.conv_forward(P, spec, bc.csr, h, keep, save)
            sv["conv"].append(s)
        segs = [h]
        for spec in self._dec:
            r = ex.dense_forward(P, spec, segs, save)
            sv["dec"].append(r)
            segs = [r["out"]]
        return segs[0], sv

    def _backward_pass(self, P, bc, sv, g_score: Optional[Tensor], flat: Optional[Tensor], accumulate: bool,
                       inject=None, for_bwd2: bool = False) -> Tensor:
        """First-order backward.  ``flat`` = grad bucket (None: skip parameter gradients).  ``inject`` = per-block
        (ot, ht) cotangents of the second-order sweep; then g_score is None (nothing flows in from the top)."""
        G = None if flat is None else ex.grad_views(self._layout, flat, self._convs)
        g = g_score
        if g is not None:
            for i in range(len(self._dec) - 1, -1, -1):
                spec, saved = self._dec[i], sv["dec"][i]
                gz, (g,) = ex.dense_backward(P, G, spec, saved, g, [(0, saved["segs"][0].shape[1])], accumulate)
                if for_bwd2:
                    saved["b_gz"] = gz
        for i in range(len(self._convs) - 1, -1, -1):
            g = ex.conv_backward(P, G, self._convs[i], bc.csr, sv["conv"][i], g, accumulate,
                                 None if inject is None else inject[i], for_bwd2)
        gz, (g,) = ex.dense_backward(P, G, self._pre[1], sv["pre"][1], g, [(0, sv["pre"][1]["segs"][0].shape[1])], accumulate)
        if for_bwd2:
            sv["pre"][1]["b_gz"] = gz
        lo = self._label_lo
        gz, (g_label,) = ex.dense_backward(P, G, self._pre[0], sv["pre"][0], g, [(lo, lo + self.configuration.NUM_CLASSES)],
                                           accumulate)
        if for_bwd2:
            sv["pre"][0]["b_gz"] = gz
        return g_label

    def _backward2_pass(self, P, bc, sv, Lt: Tensor, flat2: Tensor, want_g_score: bool) -> Optional[Tensor]:
        """Second-order sweep: Lt = cotangent on g_label.  Accumulates the direct parameter cotangents into flat2,
        then runs the injected first-order sweep (also into flat2).  Returns the cotangent on g_score."""
        G2 = ex.grad_views(self._layout, flat2, self._convs)
        lo, k = self._label_lo, self.configuration.NUM_CLASSES
        W0 = P["mlp_encoder.0.weight"]
        # layer pre[0] backward was: gz0 = g_a * [x_a>0] ; g_label = gz0 @ W0[:, lo:lo+k]
        s0, s1 = sv["pre"][0], sv["pre"][1]
        lib.dense_wgrad(s0["b_gz"], [Lt], dW=G2["mlp_encoder.0.weight"][:, lo:lo + k], accumulate=True)
        t = lib.dense_fwd([Lt], W0, cols=(lo, lo + k))["out"]                   # cot(gz0) = Lt @ W0w^T
        t = lib.ln_act_bwd(t, s0["out"], ACT_RELU)[0]                            # cot(g_a) = cot(gz0) * mask_a
        # layer pre[1] backward was: gz1 = g_b * [x_b>0] ; g_a = gz1 @ W1
        lib.dense_wgrad(s1["b_gz"], [t], dW=G2["mlp_encoder.2.weight"], accumulate=True)
        t = lib.dense_fwd([t], P["mlp_encoder.2.weight"])["out"]                # cot(gz1) = cot(g_a) @ W1^T
        t = lib.ln_act_bwd(t, s1["out"], ACT_RELU)[0]                            # cot(g_b)
        inject = []
        for spec, saved in zip(self._convs, sv["conv"]):
            t, inj = ex.conv_backward2(P, G2, spec, bc.csr, saved, t)
            inject.append(inj)
        gt = None
        for spec, saved in zip(self._dec, sv["dec"]):
            # backward was: gz = g_y * act'(y) ; g_in = gz @ W     (t = cot(g_in))
            lib.dense_wgrad(saved["b_gz"], [t], dW=G2[spec.lin + ".weight"], accumulate=True)
            t = lib.dense_fwd([t], P[spec.lin + ".weight"])["out"]               # cot(gz)
            if spec.act != ACT_NONE:
                t = lib.ln_act_bwd(t, saved["out"], spec.act)[0]                 # cot(g_y)
        gt = t if want_g_score else None
        self._backward_pass(P, bc, sv, None, flat2, True, inject=inject)
        return gt


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: VoxelGNNDiscriminator, bc, keeps, need, label, *params):
        P = dict(zip(model._names, params))
        score, sv = model._forward_pass(P, bc, label, keeps, save=need)
        if getattr(model, "debug_keep_saved", False):
            model.debug_saved = sv
        ctx.model, ctx.bc, ctx.sv, ctx.nparams = model, bc, sv if need else None, len(params)
        ctx.save_for_backward(label, *params)
        return score

    @staticmethod
    def backward(ctx, g_score):
        if ctx.sv is None:
            raise RuntimeError("discriminator backward called but the forward ran without grad")
        label, *params = ctx.saved_tensors
        outs = _DiscBwdFn.apply(ctx.model, ctx.bc, ctx.sv, torch.is_grad_enabled(), g_score.contiguous(), label, *params)
        return (None, None, None, None) + tuple(outs)


class _DiscBwdFn(torch.autograd.Function):
    """The discriminator's first-order backward as a differentiable op: (g_score, label, params) ->
    (g_label, param grads).  Its own backward is the second-order sweep; parameter gradients are not
    differentiable a second time (the reference never asks for that: only_inputs=True, trainer.py:311)."""

    @staticmethod
    def forward(ctx, model: VoxelGNNDiscriminator, bc, sv, second_order: bool, g_score, label, *params):
        P = dict(zip(model._names, params))
        flat = torch.empty(model._layout.total, dtype=torch.float32, device=g_score.device)
        g_label = model._backward_pass(P, bc, sv, g_score, flat, False, for_bwd2=second_order)
        ctx.model, ctx.bc, ctx.sv, ctx.P = model, bc, sv, P
        grads = tuple(model._layout.view(flat, n) for n in model._names)
        ctx.mark_non_differentiable(*grads)
        return (g_label,) + grads

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, Lt, *unused):
        model = ctx.model
        flat2 = torch.zeros(model._layout.total, dtype=torch.float32, device=Lt.device)
        gt = model._backward2_pass(ctx.P, ctx.bc, ctx.sv, Lt.contiguous(), flat2, ctx.needs_input_grad[4])
        # the cotangent on `label` (third-order coupling into the generator) is not needed by WGAN-GP: the
        # interpolate is built from detached samples (trainer.py:298-301)
        return (None, None, None, None, gt, None) + tuple(model._layout.view(flat2, n) for n in model._names)


# ------------------------------------------------------------------------------------------------
# native discriminator passes: bg_disc_forward / bg_disc_backward / bg_disc_backward2
# ------------------------------------------------------------------------------------------------
class _DiscNativeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: VoxelGNNDiscriminator, bc, keeps, need, ticket, label, *tensors):
        L, st = lib.load(), model._native
        params = _param_list(model)
        dev, n, e = label.device, bc.n, bc.csr.num_edges
        ws = lib.u8_buffer(L.bg_disc_fwd_ws(C.byref(st.md), n, e), dev)
        red = lib.workspace(lib.RED_BYTES, dev)
        score = torch.empty(n, 1, dtype=torch.float32, device=dev)
        kp = None if not keeps or keeps[0] is None else lib.ptr_array(keeps)
        lib._check(L.bg_disc_forward(C.byref(st.md), st.ptrs(params), C.byref(bc.csr.c_struct()), C.byref(_batch_in(bc)),
                                     label.data_ptr(), kp, int(model.training), ticket[0], ticket[1], ws.data_ptr(), ws.numel(),
                                     red.data_ptr(), red.numel() * 4, score.data_ptr(), lib._stream()))
        lib.pass_launches(2 + 4 * len(model._convs) + 4)
        if getattr(model, "debug_keep_saved", False):
            model.debug_saved = _saved_views(model, ws, n, score, False)
        ctx.need, ctx.bucket_mode, ctx.lane = need, ticket[2], st.lane
        # an undefined output gradient stays None (see backward): autograd would otherwise hand this node a zero-filled g_score
        ctx.set_materialize_grads(False)
        ctx.n_inputs = 6 + len(tensors)
        if need:
            ctx.model, ctx.bc, ctx.ws, ctx.training = model, bc, ws, model.training
            ctx.save_for_backward(label, score, *tensors)
        return score

    @staticmethod
    def backward(ctx, g_score):
        if g_score is None:  # nothing flows into the score (e.g. the gradient-penalty pass inside d_loss.backward())
            return (None,) * ctx.n_inputs
        if not ctx.need:
            raise RuntimeError("discriminator backward called but the forward ran without grad")
        label, score, *tensors = ctx.saved_tensors
        if ctx.lane is not None:  # runs on the lane's stream (autograd: the forward's stream); g_score came from another one
            g_score.record_stream(torch.cuda.current_stream())
        # NOTE (BG_GRADS=bucket): a backward that runs under create_graph=True delivers INPUT gradients only - that is the
        # WGAN-GP ``autograd.grad(score, inputs=[interpolated], create_graph=True)`` of trainer.py:306-312.  It cannot be told
        # apart here from ``loss.backward(create_graph=True)`` (``ctx.needs_input_grad`` is fixed at forward time and says True
        # for the parameters' anchor in both cases; a custom Function does not see the engine's pruning), so the latter is
        # documented as unsupported in this mode (module docstring) instead of being detected: use
        # BG_GRADS=autograd for it.
        # `score` enters detached: the first-order backward reads it as a saved value only.  Passed as this node's own output
        # it would (under create_graph=True) put an edge from the differentiable backward op back to this forward node, and
        # d_loss.backward() of the WGAN-GP critic loss would then run a second, full first-order backward of the gradient-
        # penalty pass with an all-zero g_score (one whole sweep + its weight gradients on the step's critical path,
        # profiles/r02b_summary.md)
        outs = _DiscNativeBwdFn.apply(ctx.model, ctx.bc, ctx.ws, ctx.training,
                                      (torch.is_grad_enabled(), ctx.bucket_mode, ctx.lane, ctx.needs_input_grad[5]), score.detach(),
                                      g_score.contiguous(), label, *tensors)
        return (None, None, None, None, None) + tuple(outs)


class _DiscNativeBwdFn(torch.autograd.Function):
    """First-order backward of the discriminator as a differentiable op (its backward = the second-order sweep)."""

    @staticmethod
    def forward(ctx, model: VoxelGNNDiscriminator, bc, ws, training, flags, score, g_score, label, *tensors):
        L, st = lib.load(), model._native
        params = _param_list(model)
        second_order, bucket_mode, lane, want_g_label = flags  # want_g_label: the label input itself asks for a gradient
        dev, n, e = label.device, bc.n, bc.csr.num_edges
        accumulate = bucket_mode
        if bucket_mode:  # under create_graph only the input gradient is wanted (autograd.grad(..., only_inputs=True))
            if second_order or getattr(st, "skip_param_grads", False):
                flat = None
            elif lane is None:
                flat = st.bind_grads(params)
            else:
                flat, accumulate = st.lane_bucket(lane, dev)
        else:
            flat = torch.empty(st.layout.total, dtype=torch.float32, device=dev)
        saved = lib.u8_buffer(L.bg_disc_bwd_saved_ws(C.byref(st.md), n, e), dev) if second_order else None
        tmp = lib.u8_buffer(L.bg_disc_tmp_ws(C.byref(st.md), n, e), dev)
        red = lib.workspace(lib.RED_BYTES, dev)
        # D(real) / D(fake) of a critic update score constant labels: no backward-input product of the first layer for them
        g_label = torch.empty_like(label) if (want_g_label or second_order) else None
        lib._check(L.bg_disc_backward(C.byref(st.md), st.ptrs(params), C.byref(bc.csr.c_struct()), C.byref(_batch_in(bc)),
                                      label.data_ptr(), ws.data_ptr(), score.data_ptr(), g_score.data_ptr(), int(training),
                                      lib._p(flat), st.goff, int(accumulate), lib._p(saved), 0 if saved is None else saved.numel(),
                                      tmp.data_ptr(), tmp.numel(), red.data_ptr(), red.numel() * 4, lib._p(g_label),
                                      lib._stream()))
        lib.pass_launches((3 * 4 + 6 * len(model._convs) + 3 * 2) if flat is not None else (2 * 4 + 5 * len(model._convs) + 4))
        ctx.model, ctx.bc, ctx.ws, ctx.saved, ctx.training, ctx.bucket_mode = model, bc, ws, saved, training, bucket_mode
        ctx.save_for_backward(label, score)
        if bucket_mode:
            return (g_label, None)
        grads = tuple(st.layout.view(flat, nm) for nm in model._names)
        ctx.mark_non_differentiable(*grads)
        return (g_label,) + grads

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, Lt, *unused):
        if ctx.saved is None:
            raise RuntimeError("second-order backward requested but the first backward ran without create_graph=True")
        L, model, bc = lib.load(), ctx.model, ctx.bc
        st = model._native
        params = _param_list(model)
        label, score = ctx.saved_tensors
        dev, n, e = label.device, bc.n, bc.csr.num_edges
        flat2 = st.bind_grads(params) if ctx.bucket_mode else torch.zeros(st.layout.total, dtype=torch.float32, device=dev)
        tmp = lib.u8_buffer(2 * L.bg_disc_tmp_ws(C.byref(st.md), n, e), dev)
        red = lib.workspace(lib.RED_BYTES, dev)
        want_gt = ctx.needs_input_grad[6]
        gt = torch.empty_like(score) if want_gt else None
        lib._check(L.bg_disc_backward2(C.byref(st.md), st.ptrs(params), C.byref(bc.csr.c_struct()), C.byref(_batch_in(bc)),
                                       label.data_ptr(), ctx.ws.data_ptr(), score.data_ptr(), ctx.saved.data_ptr(),
                                       Lt.contiguous().data_ptr(), int(ctx.training), flat2.data_ptr(), st.goff, tmp.data_ptr(),
                                       tmp.numel(), red.data_ptr(), red.numel() * 4, lib._p(gt), lib._stream()))
        lib.pass_launches(6 + 8 * len(model._convs) + 8 + 6 * len(model._convs) + 6)
        # the cotangent on `label` (third-order coupling into the generator) is not needed by WGAN-GP: the interpolate is
        # built from detached samples (trainer.py:298-301)
        if ctx.bucket_mode:
            return (None, None, None, None, None, None, gt, None, None)
        return (None, None, None, None, None, None, gt, None) + tuple(st.layout.view(flat2, nm) for nm in model._names)
