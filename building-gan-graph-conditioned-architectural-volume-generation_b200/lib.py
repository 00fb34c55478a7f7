"""ctypes binding of ``libbgb200.so`` (C ABI declared in ``include/bg_b200.h``).

Host side of the drop-in boundary: torch owns every buffer (device memory, streams); this module
passes raw pointers + the current CUDA stream to the library and raises ``RuntimeError`` with
``bg_last_error()`` on any non-zero return.  There is NO fallback of any kind: if the shared
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``) every entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

# The overlapped training step (step.Lanes) keeps ~8 streams busy (4 lanes + their weight-gradient side streams); with the
# driver's default of 8 hardware queues streams alias and serialise (measured 19.1 vs 15.5 ms per step).  Only effective if set
# before the CUDA context exists, i.e. when this package is imported before the first CUDA call; never overrides the user's value.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbgb200.so")
MAX_SEG = 6
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
SUPPORTED_WIDTHS = (1, 2, 4, 8, 16, 32, 64, 128)

c_f32p = C.c_void_p
c_i32p = C.c_void_p


class BgGraph(C.Structure):
    _fields_ = [("rowptr", C.c_void_p), ("col", C.c_void_p), ("cscptr", C.c_void_p), ("cscrow", C.c_void_p),
                ("perm", C.c_void_p), ("graph_ptr", C.c_void_p), ("N", C.c_int64), ("E", C.c_int64),
                ("B", C.c_int32), ("max_deg", C.c_int32)]


class BgSeg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("gather", C.c_void_p), ("width", C.c_int32), ("ld", C.c_int32)]


class BgDense(C.Structure):
    _fields_ = [("N", C.c_int64), ("nseg", C.c_int32), ("seg", BgSeg * MAX_SEG), ("W", C.c_void_p),
                ("w_so", C.c_int64), ("w_sk", C.c_int64), ("Cout", C.c_int32), ("bias", C.c_void_p),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("act", C.c_int32), ("att_src", C.c_void_p),
                ("att_dst", C.c_void_p), ("out", C.c_void_p), ("ld_out", C.c_int64), ("xhat", C.c_void_p),
                ("rstd", C.c_void_p), ("s", C.c_void_p), ("d", C.c_void_p), ("gate", C.c_void_p), ("ld_gate", C.c_int64),
                ("gate_slope", C.c_float)]


class BgModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("local_dim", "voxel_dim", "num_classes", "z_dim", "le_dim", "le_layers", "g_hidden",
                                         "g_mlp_layers", "g_repeat", "d_hidden", "d_repeat")]


class BgBatchIn(C.Structure):
    _fields_ = [("table", C.c_void_p), ("type32", C.c_void_p), ("vx", C.c_void_p)]


MAX_PEERS = 8


class BgPeers(C.Structure):
    _fields_ = [("grad", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS), ("rank", C.c_int32), ("world", C.c_int32)]


class BgWgrad(C.Structure):
    _fields_ = [("N", C.c_int64), ("gz", C.c_void_p), ("ld_gz", C.c_int64), ("Cout", C.c_int32), ("nseg", C.c_int32),
                ("seg", BgSeg * MAX_SEG), ("dW", C.c_void_p), ("ld_dw", C.c_int64), ("dbias", C.c_void_p),
                ("accumulate", C.c_int32),
                ("workspace", C.c_void_p), ("ws_bytes", C.c_size_t)]


SMALL_MAX_LAYERS, SMALL_MAX_ROWS = 8, 8


class BgSmallLayer(C.Structure):
    _fields_ = [("W", C.c_void_p), ("bias", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("cin", C.c_int32), ("cout", C.c_int32), ("act", C.c_int32),
                ("out", C.c_void_p), ("xhat", C.c_void_p), ("rstd", C.c_void_p),
                ("dW", C.c_void_p), ("dbias", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p)]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
_P, _I64, _I32, _F, _SZ, _U64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t, C.c_uint64
_MD, _BI, _GR = C.POINTER(BgModelDesc), C.POINTER(BgBatchIn), C.POINTER(BgGraph)
SIGNATURES = {
    "bg_version": (C.c_int, []),
    "bg_last_error": (C.c_char_p, []),
    "bg_csr_build_host": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "bg_type_table": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "bg_type_scatter_sum": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _SZ, _P]),
    "bg_type_scatter_sum_ws": (_SZ, [_I64, _I32, _I32]),
    "bg_dense_fwd": (C.c_int, [C.POINTER(BgDense), _P]),
    "bg_set_dense_tc": (C.c_int, [_I32]),
    "bg_set_rowdense": (C.c_int, [_I32]),
    "bg_set_dense_mma": (C.c_int, [_I32]),
    "bg_small_mlp_fwd": (C.c_int, [C.POINTER(BgSmallLayer), _I32, _P, _I32, _P]),
    "bg_small_mlp_bwd": (C.c_int, [C.POINTER(BgSmallLayer), _I32, _P, _I32, _P, _P, _I32, _P]),
    "bg_dense_wgrad_ws": (_SZ, [_I64, _I32, _I32]),
    "bg_dense_wgrad": (C.c_int, [C.POINTER(BgWgrad), _P]),
    "bg_wgrad_multi_ws": (_SZ, [_I64, _I32, _P, _P]),
    "bg_wgrad_multi": (C.c_int, [C.POINTER(BgWgrad), _I32, _P, _SZ, _P]),
    "bg_ln_act_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _I32, _P, _SZ, _P]),
    "bg_ln_act_bwd_ws": (_SZ, [_I64, _I32]),
    "bg_tune": (C.c_int, [_I32, _I32]),
    "bg_gat_fwd": (C.c_int, [C.POINTER(BgGraph), _P, _P, _P, _P, _P, _P, _P, _I32, _F, _P]),
    "bg_gat_fwd_tma": (C.c_int, [C.POINTER(BgGraph), _P, _P, _P, _P, _P, _P, _P, _I32, _F, _P]),
    "bg_set_gat_tma": (C.c_int, [_I32]),
    "bg_gat_fwd_gn_ws": (_SZ, [_I64, _I32]),
    "bg_gat_fwd_gn": (C.c_int, [C.POINTER(BgGraph), _P, _P, _P, _P, _P, _P, _P, _I32, _F, _P, _F, _P, _P, _SZ, _P]),
    "bg_graphnorm_apply": (C.c_int, [_P, _P, _P, _P, _P, _P, _F, C.c_uint64, C.c_uint64, _I64, _I32, _P, _P]),
    "bg_gat_bwd": (C.c_int, [C.POINTER(BgGraph), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _F, _P]),
    "bg_gat_bwd2": (C.c_int, [C.POINTER(BgGraph)] + [_P] * 15 + [_I32, _F, _P]),
    "bg_gat_bwd_gn": (C.c_int, [C.POINTER(BgGraph)] + [_P] * 7 + [_F] + [_P] * 13 + [_I32, _F, _P]),
    "bg_graphnorm_bwd_moments": (C.c_int, [_P, _P, _P, _P, _P, _P, _F, _I64, _I32, _P, _I32, _P, _P, _SZ, _P]),
    "bg_gcn_norm": (C.c_int, [C.POINTER(BgGraph), _P, _P]),
    "bg_spmm": (C.c_int, [C.POINTER(BgGraph), _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "bg_gatv2_fwd": (C.c_int, [C.POINTER(BgGraph)] + [_P] * 8 + [_I32, _F, _P]),
    "bg_gatv2_bwd": (C.c_int, [C.POINTER(BgGraph)] + [_P] * 12 + [_I32, _F, _P]),
    "bg_gatv2_bwd2": (C.c_int, [C.POINTER(BgGraph)] + [_P] * 14 + [_I32, _F, _P]),
    "bg_graphnorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _F, _U64, _U64, _I64, _I32, _F, _P, _P, _P, _SZ, _P]),
    "bg_graphnorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _F, _I64, _I32, _P, _P, _I32, _P, _P, _SZ, _P]),
    "bg_graphnorm_bwd2": (C.c_int, [_P] * 8 + [_F, _I64, _I32, _P, _P, _P, _I32, _P, _SZ, _P]),
    "bg_graphnorm_ws": (_SZ, [_I64, _I32]),
    "bg_gumbel_st_fwd": (C.c_int, [_P, _P, _U64, _U64, _I64, _I32, _P, _P, _P, _P]),
    "bg_gumbel_st_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    "bg_segment_softmax": (C.c_int, [_P, _P, _I64, _P, _P]),
    "bg_segment_pool": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "bg_segment_confusion": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    "bg_gen_num_params": (C.c_int32, [_MD]),
    "bg_disc_num_params": (C.c_int32, [_MD]),
    "bg_gen_fwd_ws": (_SZ, [_MD, _I64, _I64]),
    "bg_gen_bwd_ws": (_SZ, [_MD, _I64, _I64]),
    "bg_gen_forward": (C.c_int, [_MD, _P, _GR, _BI, _P, _P, _P, _I32, _U64, _U64, _P, _SZ, _P, _SZ, _P, _P, _P, _P]),
    "bg_gen_backward": (C.c_int, [_MD, _P, _GR, _BI, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _I32, _P, _SZ, _P, _SZ, _P]),
    "bg_disc_fwd_ws": (_SZ, [_MD, _I64, _I64]),
    "bg_disc_bwd_saved_ws": (_SZ, [_MD, _I64, _I64]),
    "bg_disc_tmp_ws": (_SZ, [_MD, _I64, _I64]),
    "bg_disc_forward": (C.c_int, [_MD, _P, _GR, _BI, _P, _P, _I32, _U64, _U64, _P, _SZ, _P, _SZ, _P, _P]),
    "bg_disc_backward": (C.c_int, [_MD, _P, _GR, _BI, _P, _P, _P, _P, _I32, _P, _P, _I32, _P, _SZ, _P, _SZ, _P, _SZ, _P, _P]),
    "bg_disc_backward2": (C.c_int, [_MD, _P, _GR, _BI, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _SZ, _P, _SZ, _P, _P]),
    "bg_gen_ws_offsets": (C.c_int32, [_MD, _I64, _P, _I32]),
    "bg_disc_ws_offsets": (C.c_int32, [_MD, _I64, _P, _I32]),
    "bg_axpy": (C.c_int, [_P, _P, _F, _I64, _P]),
    "bg_fill": (C.c_int, [_P, _F, _I64, _P]),
    "bg_gp_mix": (C.c_int, [_P, _P, _I32, _P, _I64, _I32, _P, _P]),
    "bg_critic_loss_ws": (_SZ, [_I64]),
    "bg_critic_loss_fwd": (C.c_int, [_P, _P, _P, _I64, _I32, _F, _P, _P, _SZ, _P, _P]),
    "bg_critic_loss_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _P, _P]),
    "bg_set_pdl": (C.c_int, [_I32]),
    "bg_set_rng_base": (C.c_int, [_P]),
    "bg_adam_flat": (C.c_int, [_P, _P, _P, _P, _I64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _I64, _P, _P]),
    "bg_p2p_allreduce_adam": (C.c_int, [C.POINTER(BgPeers), _P, _P, _P, _P, _P, _P, _I64, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, _I64, _P, _P]),
}

_lib = None

# ---- launch accounting (bench.py reports `gpu_launches`) and optional per-op CUDA-event timing ----
LAUNCHES = 0                      # kernels of libbgb200 launched so far by this process
_prof: Optional[list] = None      # when a list: (op name, tensor shapes, start event, end event) per call


def profile_begin() -> None:
    global _prof
    _prof = []


def profile_end() -> Dict[str, Dict[str, float]]:
    """Per-op totals {name: {calls, ms}} from CUDA events recorded on the launching stream."""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out: Dict[str, Dict[str, float]] = {}
    for name, shapes, e0, e1 in rec:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "by_shape": {}})
        ms = e0.elapsed_time(e1)
        d["calls"] += 1
        d["ms"] += ms
        b = d["by_shape"].setdefault(shapes, [0, 0.0])
        b[0] += 1
        b[1] += ms
    return out


def _op(name: str, launches: int):
    def deco(fn):
        def wrapped(*a, **k):
            global LAUNCHES
            LAUNCHES += launches
            if _prof is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            shapes = tuple(tuple(t.shape) for t in a if isinstance(t, Tensor))
            e0.record()
            r = fn(*a, **k)
            e1.record()
            _prof.append((name, shapes, e0, e1))
            return r
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def load() -> C.CDLL:
    """Load libbgb200.so and bind every symbol of the header.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a kernel library has not been built. Run "
                "`python -c \"import __graft_entry__ as g; g.build()\"` (or `make -C <pkg>/csrc`). "
                "There is no CPU/torch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


DENSE_MODES = {"ffma": 0, "tcgen05": 1, "3xtf32": 1, "bf16": 2}


def set_dense_tc(mode) -> int:
    """Dense-layer mode of the 128/64-wide layers, returns the previous one (0/1/2).  False / 0 / "ffma": FP32 FFMA kernels only
    (strict parity mode, env BG_DENSE_TC=0); True / 1 / "tcgen05": tcgen05 3xTF32, fp32-accurate (default); 2 / "bf16": tcgen05
    with bf16 operands and fp32 accumulation (reduced precision, env BG_DENSE_TC=bf16; tolerance stated in
    tests/test_models_gpu.py::test_bf16_dense_mode)."""
    if isinstance(mode, str):
        mode = DENSE_MODES[mode.lower()]
    return load().bg_set_dense_tc(int(mode))


def set_rowdense(on) -> int:
    """Row-per-thread kernel for the very narrow dense layers of the latency-bound regime: 0 off, 1 on (default, env
    BG_ROWDENSE), 2 = also for every layer up to K 128 / Cout 64; returns the previous setting."""
    return load().bg_set_rowdense(int(on))


def set_dense_mma(on: bool) -> bool:
    """Warp-MMA 3xTF32 kernel for the small single-segment dense layers on/off (default on, env BG_DENSE_MMA)."""
    return bool(load().bg_set_dense_mma(int(bool(on))))


def last_error() -> str:
    return load().bg_last_error().decode()


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libbgb200 error {rc}: {last_error()}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """Raw cudaStream_t of torch's current stream on the current device (every launch goes there).  The private fast path
    skips ~6 us of Python per call (it is called ~100 times per training step); falls back to the public API."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: libbgb200 kernels need CUDA tensors (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    return t


def _cf32(t: Tensor, name: str) -> Tensor:
    _f32(t, name)
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


# ------------------------------------------------------------------------------------------------
# workspace: one persistent zero-initialised buffer per (device, stream); the first 4096 bytes hold
# the self-resetting ticket counters of the "last CTA folds" reductions.
# ------------------------------------------------------------------------------------------------
_workspaces: Dict[Tuple[int, int], Tensor] = {}


def workspace(nbytes: int, device: torch.device) -> Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    need = (nbytes + 3) // 4
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 4 << 20), dtype=torch.float32, device=device)
        _workspaces[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------
# H1: host graph build
# ------------------------------------------------------------------------------------------------
def csr_build_host(edge_index: Tensor, num_nodes: int):
    lib = load()
    assert edge_index.dtype == torch.int64 and edge_index.device.type == "cpu" and edge_index.dim() == 2
    ei = edge_index.contiguous()
    e_in = int(ei.shape[1])
    cap = e_in + num_nodes
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32)
    cscptr = torch.empty(num_nodes + 1, dtype=torch.int32)
    col = torch.empty(cap, dtype=torch.int32)
    cscrow = torch.empty(cap, dtype=torch.int32)
    perm = torch.empty(cap, dtype=torch.int32)
    e_out, max_deg = C.c_int64(0), C.c_int32(0)
    _check(lib.bg_csr_build_host(ei.data_ptr(), e_in, num_nodes, rowptr.data_ptr(), col.data_ptr(), cscptr.data_ptr(),
                                 cscrow.data_ptr(), perm.data_ptr(), C.byref(e_out), C.byref(max_deg)))
    e = int(e_out.value)
    arrays = dict(rowptr=rowptr, col=col[:e].clone() if e < cap else col, cscptr=cscptr,
                  cscrow=cscrow[:e].clone() if e < cap else cscrow, perm=perm[:e].clone() if e < cap else perm)
    return arrays, e, int(max_deg.value)


def make_bg_graph(csr) -> BgGraph:
    g = BgGraph()
    g.rowptr, g.col = csr.rowptr.data_ptr(), csr.col.data_ptr()
    g.cscptr, g.cscrow, g.perm = csr.cscptr.data_ptr(), csr.cscrow.data_ptr(), csr.perm.data_ptr()
    g.graph_ptr = csr.graph_ptr.data_ptr()
    g.N, g.E, g.B, g.max_deg = csr.num_nodes, csr.num_edges, csr.num_graphs, csr.max_deg
    return g


# ------------------------------------------------------------------------------------------------
# H2
# ------------------------------------------------------------------------------------------------
@_op("type_table", 1)
def type_table(local_x: Tensor, local_type: Tensor, num_types: int) -> Tensor:
    lib = load()
    x = _cf32(local_x, "local_x")
    t = local_type.contiguous()
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    table = torch.empty(num_types, x.shape[1], dtype=torch.float32, device=x.device)
    _check(lib.bg_type_table(x.data_ptr(), t.data_ptr(), x.shape[0], x.shape[1], num_types, table.data_ptr(), _stream()))
    return table


@_op("type_scatter_sum", 1)
def type_scatter_sum(g: Tensor, type32: Tensor, num_types: int, width: Optional[int] = None) -> Tensor:
    """out[t] = sum of rows of g[:, :width] whose type is t (backward of table[type])."""
    lib = load()
    _f32(g, "g")
    assert g.stride(1) == 1 and type32.dtype == torch.int32
    n, c = g.shape[0], (width or g.shape[1])
    out = torch.empty(num_types, c, dtype=torch.float32, device=g.device)
    nb = lib.bg_type_scatter_sum_ws(n, c, num_types)
    ws = workspace(nb, g.device)
    _check(lib.bg_type_scatter_sum(g.data_ptr(), g.stride(0), type32.data_ptr(), n, c, num_types, out.data_ptr(),
                                   ws.data_ptr(), ws.numel() * 4, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# dense
# ------------------------------------------------------------------------------------------------
Seg = Union[None, Tensor, Tuple[Tensor, Optional[Tensor]]]


def _fill_segs(arr, segs: Sequence[Seg]) -> Tuple[int, int]:
    if not 1 <= len(segs) <= MAX_SEG:
        raise RuntimeError(f"dense input needs 1..{MAX_SEG} segments, got {len(segs)}")
    n_rows, k = -1, 0
    for i, sg in enumerate(segs):
        if sg is None:  # column of ones
            arr[i].ptr, arr[i].gather, arr[i].width, arr[i].ld = None, None, 1, 0
            k += 1
            continue
        t, gather = sg if isinstance(sg, tuple) else (sg, None)
        _f32(t, f"segment {i}")
        if t.dim() != 2 or t.stride(1) != 1:
            raise RuntimeError(f"segment {i}: expected a 2-D tensor with unit column stride")
        arr[i].ptr, arr[i].width, arr[i].ld = t.data_ptr(), t.shape[1], t.stride(0)
        if gather is not None:
            assert gather.dtype == torch.int32 and gather.is_contiguous()
            arr[i].gather = gather.data_ptr()
            rows = gather.shape[0]
        else:
            arr[i].gather = None
            rows = t.shape[0]
        if n_rows not in (-1, rows):
            raise RuntimeError(f"segment {i}: row count {rows} != {n_rows}")
        n_rows = rows
        k += t.shape[1]
    return n_rows, k


@_op("dense_fwd", 1)
def _small_chain(x: Tensor, layers):
    """layers: list of dicts {W, bias?, ln? = (gamma, beta), act?}; returns (ctypes array, per-layer tensor dicts)."""
    _cf32(x, "x")
    if not 1 <= len(layers) <= SMALL_MAX_LAYERS:
        raise RuntimeError(f"small_mlp: {len(layers)} layers (1..{SMALL_MAX_LAYERS})")
    arr = (BgSmallLayer * len(layers))()
    for a, l in zip(arr, layers):
        W = _cf32(l["W"], "W")
        a.W, a.cout, a.cin = W.data_ptr(), W.shape[0], W.shape[1]
        a.bias = _p(l.get("bias"))
        if l.get("ln") is not None:
            a.gamma, a.beta = l["ln"][0].data_ptr(), l["ln"][1].data_ptr()
        a.act = l.get("act", ACT_NONE)
    return arr


def small_mlp_fwd(x: Tensor, layers):
    """bg_small_mlp_fwd: the whole chain of Linear (+LayerNorm) (+activation) layers on <= 8 rows in one launch.
    Returns a list of {"out", "xhat"?, "rstd"?} per layer."""
    arr = _small_chain(x, layers)
    rows, dev, res = x.shape[0], x.device, []
    for a, l in zip(arr, layers):
        r = {"out": torch.empty(rows, a.cout, dtype=torch.float32, device=dev)}
        a.out = r["out"].data_ptr()
        if l.get("ln") is not None:
            r["xhat"] = torch.empty(rows, a.cout, dtype=torch.float32, device=dev)
            r["rstd"] = torch.empty(rows, dtype=torch.float32, device=dev)
            a.xhat, a.rstd = r["xhat"].data_ptr(), r["rstd"].data_ptr()
        res.append(r)
    _check(load().bg_small_mlp_fwd(arr, len(layers), x.data_ptr(), rows, _stream()))
    return res


def small_mlp_bwd(x: Tensor, layers, saved, gout: Tensor, need_gin: bool = True, grads=None):
    """bg_small_mlp_bwd.  ``saved`` = small_mlp_fwd's result.  ``grads`` (optional): per-layer dicts of existing
    {"dW", "dbias", "dgamma", "dbeta"} tensors to ACCUMULATE into; otherwise fresh tensors are written.
    Returns (gin or None, list of per-layer gradient dicts)."""
    arr = _small_chain(x, layers)
    _cf32(gout, "gout")
    rows, dev = x.shape[0], x.device
    out = []
    for i, (a, l, sv) in enumerate(zip(arr, layers, saved)):
        a.out = sv["out"].data_ptr()
        if l.get("ln") is not None:
            a.xhat, a.rstd = sv["xhat"].data_ptr(), sv["rstd"].data_ptr()
        g = dict(grads[i]) if grads is not None else {}
        if grads is None:
            g["dW"] = torch.empty(a.cout, a.cin, dtype=torch.float32, device=dev)
            if l.get("bias") is not None:
                g["dbias"] = torch.empty(a.cout, dtype=torch.float32, device=dev)
            if l.get("ln") is not None:
                g["dgamma"] = torch.empty(a.cout, dtype=torch.float32, device=dev)
                g["dbeta"] = torch.empty(a.cout, dtype=torch.float32, device=dev)
        a.dW, a.dbias, a.dgamma, a.dbeta = _p(g.get("dW")), _p(g.get("dbias")), _p(g.get("dgamma")), _p(g.get("dbeta"))
        out.append(g)
    gin = torch.empty(rows, arr[0].cin, dtype=torch.float32, device=dev) if need_gin else None
    _check(load().bg_small_mlp_bwd(arr, len(layers), x.data_ptr(), rows, gout.data_ptr(), _p(gin), int(grads is not None), _stream()))
    return gin, out


def dense_fwd(segs: Sequence[Seg], W: Tensor, bias: Optional[Tensor] = None, ln: Optional[Tuple[Tensor, Tensor]] = None,
              act: int = ACT_NONE, att: Optional[Tuple[Tensor, Tensor]] = None, transposed: bool = False,
              save_ln: bool = False, out: Optional[Tensor] = None, cols: Optional[Tuple[int, int]] = None,
              gate: Optional[Tensor] = None, gate_slope: float = 0.0):
    """out = act(LN(X @ Wop^T + bias)).  ``transposed=False``: Wop = W ([Cout,K]).  ``transposed=True``:
    Wop = W^T, i.e. out = X @ W (the backward-input product); ``cols=(a,b)`` then restricts the output to
    columns a..b of W (only those input gradients are needed).  Returns a dict of the produced tensors."""
    lib = load()
    a = BgDense()
    n, k = _fill_segs(a.seg, segs)
    _cf32(W, "W")
    if transposed:
        lo, hi = cols if cols is not None else (0, W.shape[1])
        if W.shape[0] != k:
            raise RuntimeError(f"dense_fwd(transposed): X has {k} columns but W has {W.shape[0]} rows")
        cout, wptr, w_so, w_sk = hi - lo, W.data_ptr() + 4 * lo, 1, W.stride(0)
    else:
        lo, hi = cols if cols is not None else (0, W.shape[1])  # cols: use only input columns lo..hi of W
        if hi - lo != k:
            raise RuntimeError(f"dense_fwd: X has {k} columns but W expects {hi - lo}")
        cout, wptr, w_so, w_sk = W.shape[0], W.data_ptr() + 4 * lo, W.stride(0), 1
    dev = W.device
    if out is None:
        out = torch.empty(n, cout, dtype=torch.float32, device=dev)
    res = {"out": out}
    a.N, a.nseg, a.W, a.w_so, a.w_sk, a.Cout = n, len(segs), wptr, w_so, w_sk, cout
    a.bias = _p(bias)
    a.act = act
    if ln is not None:
        a.ln_gamma, a.ln_beta = ln[0].data_ptr(), ln[1].data_ptr()
        if save_ln:
            res["xhat"] = torch.empty(n, cout, dtype=torch.float32, device=dev)
            res["rstd"] = torch.empty(n, dtype=torch.float32, device=dev)
            a.xhat, a.rstd = res["xhat"].data_ptr(), res["rstd"].data_ptr()
    if att is not None:
        a.att_src, a.att_dst = att[0].data_ptr(), att[1].data_ptr()
        res["s"] = torch.empty(n, dtype=torch.float32, device=dev)
        res["d"] = torch.empty(n, dtype=torch.float32, device=dev)
        a.s, a.d = res["s"].data_ptr(), res["d"].data_ptr()
    a.out, a.ld_out = out.data_ptr(), out.stride(0)
    if gate is not None:  # out *= gate > 0 ? 1 : gate_slope  (activation backward of the layer below, fused)
        _f32(gate, "gate")
        assert gate.shape == out.shape and gate.stride(1) == 1
        a.gate, a.ld_gate, a.gate_slope = gate.data_ptr(), gate.stride(0), gate_slope
    _check(lib.bg_dense_fwd(C.byref(a), _stream()))
    return res


@_op("dense_wgrad", 2)
def dense_wgrad(gz: Tensor, segs: Sequence[Seg], dW: Optional[Tensor] = None, accumulate: bool = False,
                dbias: Optional[Tensor] = None) -> Tensor:
    """dW[o,k] = sum_n gz[n,o] X[n,k]; a ``None`` segment is a column of ones (=> bias gradient column).
    ``dbias`` routes the last column (put the ones segment last) to a separate [Cout] tensor."""
    lib = load()
    a = BgWgrad()
    _f32(gz, "gz")
    assert gz.dim() == 2 and gz.stride(1) == 1
    n, k = _fill_segs(a.seg, segs)
    if n == -1:
        n = gz.shape[0]
    cout = gz.shape[1]
    if dW is None:
        dW = torch.empty(cout, k - (1 if dbias is not None else 0), dtype=torch.float32, device=gz.device)
        accumulate = False
    assert dW.stride(1) == 1 or dW.shape[1] == 1
    nb = lib.bg_dense_wgrad_ws(n, cout, k)
    ws = workspace(nb, gz.device)
    a.N, a.gz, a.ld_gz, a.Cout, a.nseg = n, gz.data_ptr(), gz.stride(0), cout, len(segs)
    a.dW, a.ld_dw, a.accumulate, a.dbias = dW.data_ptr(), dW.stride(0), int(accumulate), _p(dbias)
    a.workspace, a.ws_bytes = ws.data_ptr(), ws.numel() * 4
    _check(lib.bg_dense_wgrad(C.byref(a), _stream()))
    return dW


@_op("ln_act_bwd", 1)
def ln_act_bwd(gout: Tensor, out: Tensor, act: int, xhat: Optional[Tensor] = None, rstd: Optional[Tensor] = None,
               gamma: Optional[Tensor] = None, dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None,
               accumulate: bool = False):
    """Pre-activation gradient gz of [LayerNorm +] activation; LayerNorm path also returns dgamma, dbeta."""
    lib = load()
    _cf32(gout, "gout"), _cf32(out, "out")
    n, c = out.shape
    gz = torch.empty_like(out)
    if xhat is None:
        _check(lib.bg_ln_act_bwd(gout.data_ptr(), out.data_ptr(), None, None, None, n, c, act, gz.data_ptr(), None, None, 0,
                                 None, 0, _stream()))
        return gz, None, None
    if dgamma is None:
        dgamma = torch.empty(c, dtype=torch.float32, device=out.device)
        dbeta = torch.empty(c, dtype=torch.float32, device=out.device)
        accumulate = False
    nb = lib.bg_ln_act_bwd_ws(n, c)
    ws = workspace(nb, out.device)
    _check(lib.bg_ln_act_bwd(gout.data_ptr(), out.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), n, c, act,
                             gz.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), int(accumulate), ws.data_ptr(),
                             ws.numel() * 4, _stream()))
    return gz, dgamma, dbeta


# ------------------------------------------------------------------------------------------------
# GAT aggregation
# ------------------------------------------------------------------------------------------------
@_op("gat_fwd", 1)
def gat_fwd(csr, h: Tensor, s: Tensor, d: Tensor, bias: Optional[Tensor], slope: float = 0.2):
    lib = load()
    _cf32(h, "h")
    n, c = h.shape
    out = torch.empty_like(h)
    m = torch.empty(n, dtype=torch.float32, device=h.device)
    z = torch.empty(n, dtype=torch.float32, device=h.device)
    _check(lib.bg_gat_fwd(C.byref(csr.c_struct()), h.data_ptr(), s.data_ptr(), d.data_ptr(), _p(bias), out.data_ptr(),
                          m.data_ptr(), z.data_ptr(), c, slope, _stream()))
    return out, m, z


@_op("gat_fwd_tma", 1)
def gat_fwd_tma(csr, h: Tensor, s: Tensor, d: Tensor, bias: Optional[Tensor], slope: float = 0.2):
    """bg_gat_fwd with the neighbour rows gathered by TMA (tile::gather4); C in {64, 128}, max in-degree <= 8."""
    lib = load()
    _cf32(h, "h")
    n, c = h.shape
    out = torch.empty_like(h)
    m = torch.empty(n, dtype=torch.float32, device=h.device)
    z = torch.empty(n, dtype=torch.float32, device=h.device)
    _check(lib.bg_gat_fwd_tma(C.byref(csr.c_struct()), h.data_ptr(), s.data_ptr(), d.data_ptr(), _p(bias), out.data_ptr(),
                              m.data_ptr(), z.data_ptr(), c, slope, _stream()))
    return out, m, z


def set_gat_tma(on: bool) -> bool:
    """Route bg_gat_fwd through the TMA-gather kernel for HBM-sized eligible graphs (env BG_GAT_TMA); returns the previous setting."""
    return bool(load().bg_set_gat_tma(int(bool(on))))


@_op("gat_fwd_gn", 2)
def gat_fwd_gn(csr, h: Tensor, s: Tensor, d: Tensor, bias: Optional[Tensor], gn_w: Tensor, gn_beta: Tensor, gn_alpha: Tensor,
               keep: Optional[Tensor] = None, keep_prob: float = 1.0, seed: int = 0, offset: int = 0, slope: float = 0.2,
               eps: float = 1e-5, apply: bool = True):
    """GATConv aggregation with the following GraphNorm's statistics fused into its epilogue, then the elementwise
    GraphNorm + ReLU + dropout pass.  Returns (o, m, z, x1, stats) like gat_fwd + graphnorm_fwd."""
    lib = load()
    _cf32(h, "h")
    n, c = h.shape
    dev = h.device
    out = torch.empty_like(h)
    m = torch.empty(n, dtype=torch.float32, device=dev)
    z = torch.empty(n, dtype=torch.float32, device=dev)
    stats = torch.empty(3 * c, dtype=torch.float32, device=dev)
    ws = workspace(lib.bg_gat_fwd_gn_ws(n, c), dev)
    _check(lib.bg_gat_fwd_gn(C.byref(csr.c_struct()), h.data_ptr(), s.data_ptr(), d.data_ptr(), _p(bias), out.data_ptr(),
                             m.data_ptr(), z.data_ptr(), c, slope, gn_alpha.data_ptr(), eps, stats.data_ptr(), ws.data_ptr(),
                             ws.numel() * 4, _stream()))
    if not apply:  # statistics only (profiling)
        return out, m, z, None, stats
    x1 = torch.empty_like(h)
    _check(lib.bg_graphnorm_apply(out.data_ptr(), gn_w.data_ptr(), gn_beta.data_ptr(), gn_alpha.data_ptr(), stats.data_ptr(),
                                  _p(keep), keep_prob, seed, offset, n, c, x1.data_ptr(), _stream()))
    return out, m, z, x1, stats


@_op("gat_bwd", 2)
def gat_bwd(csr, gout: Tensor, h: Tensor, s: Tensor, d: Tensor, m: Tensor, z: Tensor, a_src: Tensor, a_dst: Tensor,
            slope: float = 0.2):
    """Returns gh_tot[N,C], gsd[N,2], and the per-edge scratch (P, DU)."""
    lib = load()
    _cf32(gout, "gout")
    n, c = h.shape
    dev = h.device
    P = torch.empty(csr.num_edges, dtype=torch.float32, device=dev)
    DU = torch.empty(csr.num_edges, dtype=torch.float32, device=dev)
    gh = torch.empty_like(h)
    gsd = torch.empty(n, 2, dtype=torch.float32, device=dev)
    _check(lib.bg_gat_bwd(C.byref(csr.c_struct()), gout.data_ptr(), h.data_ptr(), s.data_ptr(), d.data_ptr(), m.data_ptr(),
                          z.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(), P.data_ptr(), DU.data_ptr(), gh.data_ptr(),
                          gsd.data_ptr(), c, slope, _stream()))
    return gh, gsd, P, DU


@_op("gat_bwd_gn", 3)
def gat_bwd_gn(csr, gx1: Tensor, o: Tensor, x1: Tensor, gn_w: Tensor, gn_alpha: Tensor, stats: Tensor, keep_scale: float,
               h: Tensor, s: Tensor, d: Tensor, m: Tensor, z: Tensor, a_src: Tensor, a_dst: Tensor,
               inj_o: Optional[Tensor] = None, slope: float = 0.2):
    """GraphNorm backward (moments launch + elementwise half fused into the aggregation backward) followed by the GAT
    backward.  Returns (go, gn dparams[3,C], gh_tot, gsd)."""
    lib = load()
    _cf32(gx1, "gx1")
    n, c = h.shape
    dev = h.device
    dpar = torch.empty(3, c, dtype=torch.float32, device=dev)
    bst = torch.empty(2 * c, dtype=torch.float32, device=dev)
    ws = workspace(lib.bg_graphnorm_ws(n, c), dev)
    _check(lib.bg_graphnorm_bwd_moments(gx1.data_ptr(), o.data_ptr(), x1.data_ptr(), gn_w.data_ptr(), gn_alpha.data_ptr(),
                                        stats.data_ptr(), keep_scale, n, c, dpar.data_ptr(), 0, bst.data_ptr(), ws.data_ptr(),
                                        ws.numel() * 4, _stream()))
    P = torch.empty(csr.num_edges, dtype=torch.float32, device=dev)
    DU = torch.empty_like(P)
    go, gh = torch.empty_like(h), torch.empty_like(h)
    gsd = torch.empty(n, 2, dtype=torch.float32, device=dev)
    _check(lib.bg_gat_bwd_gn(C.byref(csr.c_struct()), gx1.data_ptr(), o.data_ptr(), x1.data_ptr(), gn_w.data_ptr(),
                             gn_alpha.data_ptr(), stats.data_ptr(), bst.data_ptr(), keep_scale, _p(inj_o), h.data_ptr(),
                             s.data_ptr(), d.data_ptr(), m.data_ptr(), z.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(),
                             P.data_ptr(), DU.data_ptr(), go.data_ptr(), gh.data_ptr(), gsd.data_ptr(), c, slope, _stream()))
    return go, dpar, gh, gsd


@_op("gat_bwd2", 2)
def gat_bwd2(csr, Ht: Tensor, St: Tensor, Dt: Tensor, gout: Tensor, h: Tensor, s: Tensor, d: Tensor, m: Tensor, z: Tensor,
             a_src: Tensor, a_dst: Tensor, slope: float = 0.2):
    """Cotangents (Ht,St,Dt) on (gh,gs,gd) -> gt[N,C] (on gout), ht_tot[N,C] (on h, incl. s/d paths), sdt[N,2]."""
    lib = load()
    _cf32(Ht, "Ht")
    n, c = h.shape
    dev = h.device
    scratch = torch.empty(4 * csr.num_edges, dtype=torch.float32, device=dev)
    gt = torch.empty_like(h)
    ht = torch.empty_like(h)
    sdt = torch.empty(n, 2, dtype=torch.float32, device=dev)
    _check(lib.bg_gat_bwd2(C.byref(csr.c_struct()), Ht.data_ptr(), St.data_ptr(), Dt.data_ptr(), gout.data_ptr(), h.data_ptr(),
                           s.data_ptr(), d.data_ptr(), m.data_ptr(), z.data_ptr(), a_src.data_ptr(), a_dst.data_ptr(),
                           scratch.data_ptr(), gt.data_ptr(), ht.data_ptr(), sdt.data_ptr(), c, slope, _stream()))
    return gt, ht, sdt


# ------------------------------------------------------------------------------------------------
# non-default conv types (GCNConv / GraphConv / GATv2Conv aggregation)
# ------------------------------------------------------------------------------------------------
@_op("gcn_norm", 1)
def gcn_norm(csr) -> Tensor:
    """Per-edge symmetric normalisation weights of GCNConv in CSR order."""
    lib = load()
    w = torch.empty(csr.num_edges, dtype=torch.float32, device=csr.device)
    _check(lib.bg_gcn_norm(C.byref(csr.c_struct()), w.data_ptr(), _stream()))
    return w


@_op("spmm", 1)
def spmm(csr, x: Tensor, w: Optional[Tensor] = None, bias: Optional[Tensor] = None, transpose: bool = False,
         self_loops: bool = True) -> Tensor:
    """Weighted neighbour sum over the in-edges (transpose=False) or out-edges (True) of every node."""
    lib = load()
    _cf32(x, "x")
    out = torch.empty_like(x)
    _check(lib.bg_spmm(C.byref(csr.c_struct()), _p(w), x.data_ptr(), _p(bias), out.data_ptr(), x.shape[1], int(transpose),
                       int(self_loops), _stream()))
    return out


@_op("gatv2_fwd", 1)
def gatv2_fwd(csr, xl: Tensor, xr: Tensor, att: Tensor, bias: Optional[Tensor], slope: float = 0.2):
    lib = load()
    _cf32(xl, "xl"), _cf32(xr, "xr")
    n, c = xl.shape
    out = torch.empty_like(xl)
    logit = torch.empty(csr.num_edges, dtype=torch.float32, device=xl.device)
    m = torch.empty(n, dtype=torch.float32, device=xl.device)
    z = torch.empty(n, dtype=torch.float32, device=xl.device)
    _check(lib.bg_gatv2_fwd(C.byref(csr.c_struct()), xl.data_ptr(), xr.data_ptr(), att.data_ptr(), _p(bias), out.data_ptr(),
                            logit.data_ptr(), m.data_ptr(), z.data_ptr(), c, slope, _stream()))
    return out, logit, m, z


@_op("gatv2_bwd", 2)
def gatv2_bwd(csr, gout: Tensor, xl: Tensor, xr: Tensor, att: Tensor, logit: Tensor, m: Tensor, z: Tensor, slope: float = 0.2):
    """Returns gxl, gxr, garow (column sum = attention-vector gradient)."""
    lib = load()
    _cf32(gout, "gout")
    c = xl.shape[1]
    P = torch.empty(csr.num_edges, dtype=torch.float32, device=xl.device)
    DL = torch.empty_like(P)
    gxl, gxr, garow = torch.empty_like(xl), torch.empty_like(xl), torch.empty_like(xl)
    _check(lib.bg_gatv2_bwd(C.byref(csr.c_struct()), gout.data_ptr(), xl.data_ptr(), xr.data_ptr(), att.data_ptr(),
                            logit.data_ptr(), m.data_ptr(), z.data_ptr(), P.data_ptr(), DL.data_ptr(), gxl.data_ptr(),
                            gxr.data_ptr(), garow.data_ptr(), c, slope, _stream()))
    return gxl, gxr, garow


@_op("gatv2_bwd2", 2)
def gatv2_bwd2(csr, Hl: Tensor, Hr: Tensor, gout: Tensor, xl: Tensor, xr: Tensor, att: Tensor, logit: Tensor, m: Tensor,
               z: Tensor, slope: float = 0.2):
    """Cotangents (Hl, Hr) on (gxl, gxr) -> gt (on gout), cxl, cxr, carow (column sum = cotangent on att)."""
    lib = load()
    _cf32(Hl, "Hl"), _cf32(Hr, "Hr")
    c = xl.shape[1]
    scratch = torch.empty(6 * csr.num_edges, dtype=torch.float32, device=xl.device)
    gt, cxl, cxr, carow = torch.empty_like(xl), torch.empty_like(xl), torch.empty_like(xl), torch.empty_like(xl)
    _check(lib.bg_gatv2_bwd2(C.byref(csr.c_struct()), Hl.data_ptr(), Hr.data_ptr(), gout.data_ptr(), xl.data_ptr(), xr.data_ptr(),
                             att.data_ptr(), logit.data_ptr(), m.data_ptr(), z.data_ptr(), scratch.data_ptr(), gt.data_ptr(),
                             cxl.data_ptr(), cxr.data_ptr(), carow.data_ptr(), c, slope, _stream()))
    return gt, cxl, cxr, carow


# ------------------------------------------------------------------------------------------------
# GraphNorm + ReLU + dropout mask
# ------------------------------------------------------------------------------------------------
@_op("graphnorm_fwd", 2)
def graphnorm_fwd(o: Tensor, w: Tensor, beta: Tensor, alpha: Tensor, keep: Optional[Tensor], keep_prob: float = 1.0,
                  seed: int = 0, offset: int = 0, eps: float = 1e-5):
    """GraphNorm + ReLU + dropout.  keep: explicit uint8 mask (then keep_prob = its Bernoulli parameter);
    keep None and keep_prob < 1: Philox mask from (seed, offset); keep None and keep_prob == 1: eval."""
    lib = load()
    _cf32(o, "o")
    n, c = o.shape
    x1 = torch.empty_like(o)
    stats = torch.empty(3 * c, dtype=torch.float32, device=o.device)
    ws = workspace(lib.bg_graphnorm_ws(n, c), o.device)
    if keep is not None:
        assert keep.dtype == torch.uint8 and keep.is_contiguous() and keep.shape == o.shape
    _check(lib.bg_graphnorm_fwd(o.data_ptr(), w.data_ptr(), beta.data_ptr(), alpha.data_ptr(), _p(keep), keep_prob, seed, offset,
                                n, c, eps, x1.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws.numel() * 4, _stream()))
    return x1, stats


@_op("graphnorm_bwd", 2)
def graphnorm_bwd(gx1: Tensor, o: Tensor, x1: Tensor, w: Tensor, alpha: Tensor, stats: Tensor, keep_scale: float,
                  dparams: Optional[Tensor] = None, accumulate: bool = False):
    """Returns go, dparams[3,C] = (dw, dbeta, dalpha), bstats[2C]."""
    lib = load()
    _cf32(gx1, "gx1")
    n, c = o.shape
    go = torch.empty_like(o)
    if dparams is None:
        dparams = torch.empty(3, c, dtype=torch.float32, device=o.device)
        accumulate = False
    bstats = torch.empty(2 * c, dtype=torch.float32, device=o.device)
    ws = workspace(lib.bg_graphnorm_ws(n, c), o.device)
    _check(lib.bg_graphnorm_bwd(gx1.data_ptr(), o.data_ptr(), x1.data_ptr(), w.data_ptr(), alpha.data_ptr(), stats.data_ptr(),
                                keep_scale, n, c, go.data_ptr(), dparams.data_ptr(), int(accumulate), bstats.data_ptr(),
                                ws.data_ptr(), ws.numel() * 4, _stream()))
    return go, dparams, bstats


@_op("graphnorm_bwd2", 2)
def graphnorm_bwd2(Xt: Tensor, gx1: Tensor, o: Tensor, x1: Tensor, w: Tensor, alpha: Tensor, stats: Tensor, bstats: Tensor,
                   keep_scale: float, dparams2: Optional[Tensor] = None, accumulate: bool = False):
    """Cotangent Xt on go -> gx1t (on gx1), ot (on o), dparams2[3,C] (on w, beta, alpha)."""
    lib = load()
    _cf32(Xt, "Xt")
    n, c = o.shape
    gx1t = torch.empty_like(o)
    ot = torch.empty_like(o)
    if dparams2 is None:
        dparams2 = torch.empty(3, c, dtype=torch.float32, device=o.device)
        accumulate = False
    ws = workspace(lib.bg_graphnorm_ws(n, c), o.device)
    _check(lib.bg_graphnorm_bwd2(Xt.data_ptr(), gx1.data_ptr(), o.data_ptr(), x1.data_ptr(), w.data_ptr(), alpha.data_ptr(),
                                 stats.data_ptr(), bstats.data_ptr(), keep_scale, n, c, gx1t.data_ptr(), ot.data_ptr(),
                                 dparams2.data_ptr(), int(accumulate), ws.data_ptr(), ws.numel() * 4, _stream()))
    return gx1t, ot, dparams2


# ------------------------------------------------------------------------------------------------
# Gumbel straight-through, segment primitives, utilities
# ------------------------------------------------------------------------------------------------
@_op("gumbel_st_fwd", 1)
def gumbel_st_fwd(logits: Tensor, noise: Optional[Tensor], seed: int = 0, offset: int = 0):
    lib = load()
    _cf32(logits, "logits")
    if noise is not None:
        _cf32(noise, "noise")
    n, k = logits.shape
    soft, hard = torch.empty_like(logits), torch.empty_like(logits)
    amax = torch.empty(n, dtype=torch.int32, device=logits.device)
    _check(lib.bg_gumbel_st_fwd(logits.data_ptr(), _p(noise), seed, offset, n, k, soft.data_ptr(), hard.data_ptr(),
                                amax.data_ptr(), _stream()))
    return soft, hard, amax


@_op("gumbel_st_bwd", 1)
def gumbel_st_bwd(g_hard: Optional[Tensor], g_soft: Optional[Tensor], soft: Tensor) -> Tensor:
    lib = load()
    n, k = soft.shape
    gl = torch.empty_like(soft)
    _check(lib.bg_gumbel_st_bwd(_p(g_hard), _p(g_soft), soft.data_ptr(), n, k, gl.data_ptr(), _stream()))
    return gl


@_op("segment_softmax", 1)
def segment_softmax(v: Tensor, seg_ptr: Tensor) -> Tensor:
    lib = load()
    _cf32(v, "v")
    assert seg_ptr.dtype == torch.int32
    out = torch.empty_like(v)
    _check(lib.bg_segment_softmax(v.data_ptr(), seg_ptr.data_ptr(), seg_ptr.numel() - 1, out.data_ptr(), _stream()))
    return out


@_op("segment_confusion", 1)
def segment_confusion(score: Tensor, target: Tensor, seg_ptr: Tensor) -> Tensor:
    """cm[S,K,K] int32: per-segment confusion counts of argmax(score) against int64 targets."""
    lib = load()
    _cf32(score, "score")
    assert target.dtype == torch.int64 and target.is_contiguous() and seg_ptr.dtype == torch.int32
    s, k = seg_ptr.numel() - 1, score.shape[1]
    cm = torch.empty(s, k, k, dtype=torch.int32, device=score.device)
    _check(lib.bg_segment_confusion(score.data_ptr(), target.data_ptr(), seg_ptr.data_ptr(), s, k, cm.data_ptr(), _stream()))
    return cm


@_op("segment_pool", 1)
def segment_pool(x: Tensor, seg_ptr: Tensor, mode: str = "mean") -> Tensor:
    lib = load()
    _cf32(x, "x")
    assert seg_ptr.dtype == torch.int32
    s = seg_ptr.numel() - 1
    out = torch.empty(s, x.shape[1], dtype=torch.float32, device=x.device)
    _check(lib.bg_segment_pool(x.data_ptr(), seg_ptr.data_ptr(), s, x.shape[1], {"mean": 0, "max": 1, "sum": 2}[mode],
                               out.data_ptr(), _stream()))
    return out


@_op("axpy", 1)
def axpy_(y: Tensor, x: Tensor, a: float = 1.0) -> Tensor:
    lib = load()
    _cf32(y, "y"), _cf32(x, "x")
    assert y.numel() == x.numel()
    _check(lib.bg_axpy(y.data_ptr(), x.data_ptr(), a, y.numel(), _stream()))
    return y


@_op("fill", 1)
def fill_(y: Tensor, v: float) -> Tensor:
    lib = load()
    _cf32(y, "y")
    _check(lib.bg_fill(y.data_ptr(), v, y.numel(), _stream()))
    return y


@_op("gp_mix", 1)
def gp_mix(e: Tensor, onehot: Tensor, soft: Tensor) -> Tensor:
    """mixed = e * onehot + (1 - e) * soft (trainer.py:298-301); onehot int64 (as the reference holds it) or fp32."""
    lib = load()
    _cf32(e, "e"), _cf32(soft, "soft")
    n, k = soft.shape
    assert e.numel() == n and tuple(onehot.shape) == (n, k) and onehot.is_contiguous() and onehot.dtype in (torch.int64, torch.float32)
    mixed = torch.empty_like(soft)
    _check(lib.bg_gp_mix(e.data_ptr(), onehot.data_ptr(), int(onehot.dtype == torch.int64), soft.data_ptr(), n, k, mixed.data_ptr(),
                         _stream()))
    return mixed


@_op("critic_loss_fwd", 1)
def critic_loss_fwd(d_fake: Tensor, d_real: Tensor, grad: Tensor, lam: float):
    """(out4 = [loss, mean D(fake), mean D(real), gp], coef[N]) - bg_critic_loss_fwd."""
    lib = load()
    _cf32(d_fake, "d_fake"), _cf32(d_real, "d_real"), _cf32(grad, "grad")
    n, k = grad.shape
    assert d_fake.numel() == n and d_real.numel() == n
    coef = torch.empty(n, dtype=torch.float32, device=grad.device)
    out4 = torch.empty(4, dtype=torch.float32, device=grad.device)
    ws = workspace(lib.bg_critic_loss_ws(n), grad.device)
    _check(lib.bg_critic_loss_fwd(d_fake.data_ptr(), d_real.data_ptr(), grad.data_ptr(), n, k, float(lam), coef.data_ptr(),
                                  ws.data_ptr(), ws.numel() * 4, out4.data_ptr(), _stream()))
    return out4, coef


@_op("critic_loss_bwd", 1)
def critic_loss_bwd(g_loss: Tensor, coef: Tensor, grad: Tensor, want_fake: bool, want_real: bool, want_grad: bool):
    lib = load()
    n, k = grad.shape
    g_fake = torch.empty(n, 1, dtype=torch.float32, device=grad.device) if want_fake else None
    g_real = torch.empty(n, 1, dtype=torch.float32, device=grad.device) if want_real else None
    g_grad = torch.empty_like(grad) if want_grad else None
    _check(lib.bg_critic_loss_bwd(g_loss.data_ptr(), coef.data_ptr(), grad.data_ptr(), n, k, _p(g_fake), _p(g_real), _p(g_grad),
                                  _stream()))
    return g_fake, g_real, g_grad


def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch on/off (process-wide); returns the previous setting."""
    return bool(load().bg_set_pdl(int(bool(on))))


def set_rng_base(base: Optional[Tensor]) -> None:
    """Device counter added to every in-kernel Philox offset (None = off); see bg_set_rng_base."""
    if base is not None:
        assert base.dtype == torch.int64 and base.is_cuda and base.numel() == 1
    _check(load().bg_set_rng_base(0 if base is None else base.data_ptr()))


@_op("adam_flat", 1)
def adam_flat_(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, step: int, step_dev: Optional[Tensor] = None) -> None:
    """torch.optim.Adam.step over index-aligned flat buffers (bg_adam_flat, H15)."""
    lib = load()
    for t, nm in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _cf32(t, nm)
    assert p.numel() == g.numel() == m.numel() == v.numel()
    sd = 0
    if step_dev is not None:
        assert step_dev.dtype == torch.int64 and step_dev.is_cuda and step_dev.numel() == 1
        sd = step_dev.data_ptr()
    _check(lib.bg_adam_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                            weight_decay, int(step), sd, _stream()))


@_op("p2p_allreduce_adam", 1)
def p2p_allreduce_adam_(peers: BgPeers, epoch: Tensor, ticket: Tensor, n: int, p: Optional[Tensor] = None, m: Optional[Tensor] = None,
                        v: Optional[Tensor] = None, gavg: Optional[Tensor] = None, lr: float = 0.0, beta1: float = 0.9,
                        beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0, step: int = 1,
                        step_dev: Optional[Tensor] = None) -> None:
    """bg_p2p_allreduce_adam: one-shot peer-memory all-reduce (average) of the ranks' flat gradient buckets fused with the Adam
    step over the local flat buffers; with p = m = v = None only the averaged gradient is written to ``gavg``."""
    lib = load()
    assert epoch.dtype == torch.int32 and ticket.dtype == torch.int32 and epoch.is_cuda and ticket.is_cuda
    for t in (p, m, v, gavg):
        if t is not None:
            _cf32(t, "flat buffer")
            assert t.numel() == n
    _check(lib.bg_p2p_allreduce_adam(C.byref(peers), epoch.data_ptr(), ticket.data_ptr(), _p(p), _p(m), _p(v), _p(gavg), n, lr, beta1, beta2,
                                     eps, weight_decay, int(step), _p(step_dev), _stream()))


# ------------------------------------------------------------------------------------------------
# whole-pass executors (csrc/bg_passes.cu): one C call per generator / discriminator pass
# ------------------------------------------------------------------------------------------------
RED_BYTES = 72 << 20


def pass_launches(n: int) -> None:
    """Launch accounting for the native passes (they bypass the per-op wrappers)."""
    global LAUNCHES
    LAUNCHES += n


def ptr_array(tensors) -> "C.Array":
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def u8_buffer(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
