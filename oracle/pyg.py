"""ORACLE (test infrastructure, see oracle/__init__.py): pure-torch restatement of the
torch-geometric 2.6.1 pieces the reference calls.

torch_geometric is a third-party dependency pinned at ``requirements.txt:11`` of the
reference (``torch-geometric==2.6.1``); it is neither vendored under /root/reference nor
installable here, so this file restates its *published algorithm* for exactly the call
sites the reference has:

* ``tgnn.GATConv`` / ``GCNConv`` / ``GraphConv`` / ``GATv2Conv``  - models.py:22-29,166-173
* ``tgnn.norm.GraphNorm``                                       - models.py:73,83,193,203
* ``tgnn.Sequential``                                            - models.py:90,210
* ``Data`` / ``Batch.from_data_list`` / ``Batch.__getitem__``    - data.py:118-163, trainer.py:364,421

**parity unpinned** for this file (no wheel, no golden vectors in the reference); validated by
dense-matrix definitions, hand-computed known answers and fp64 gradcheck in
tests/test_oracle_pyg.py.  No torch_scatter/pyg_lib is pinned, so ``scatter`` follows PyG's
pure-torch branch (scatter_add_ / scatter_reduce_('amax', include_self=False)).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence as Seq, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn


# --------------------------------------------------------------------------------------
# utils.scatter / utils.softmax / self-loop helpers
# --------------------------------------------------------------------------------------
def scatter(src: Tensor, index: Tensor, dim_size: int, reduce: str = "sum") -> Tensor:
    """PyG ``utils.scatter`` along dim 0, pure-torch branch."""
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return src.new_zeros(shape).scatter_add_(0, idx, src)
    if reduce == "mean":
        count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.size(0)))
        count = count.clamp(min=1)
        out = src.new_zeros(shape).scatter_add_(0, idx, src)
        return out / count.view(-1, *([1] * (src.dim() - 1)))
    if reduce == "max":
        return src.new_zeros(shape).scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def segment_softmax(src: Tensor, index: Tensor, num_segments: int) -> Tensor:
    """PyG ``utils.softmax(src, index, num_nodes=N)``: max on detached values, +1e-16 in the sum."""
    src_max = scatter(src.detach(), index, num_segments, reduce="max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, num_segments, reduce="sum") + 1e-16
    return out / out_sum.index_select(0, index)


def remove_self_loops(edge_index: Tensor) -> Tensor:
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask]


def add_self_loops(edge_index: Tensor, num_nodes: int) -> Tensor:
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)


def gat_edges(edge_index: Tensor, num_nodes: int) -> Tensor:
    """The edge list GATConv actually aggregates over: self loops stripped, then one self
    loop per node APPENDED AT THE END (order matters for CPU summation order)."""
    return add_self_loops(remove_self_loops(edge_index), num_nodes)


def glorot_(t: Tensor) -> Tensor:
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-bound, bound)


# --------------------------------------------------------------------------------------
# conv layers
# --------------------------------------------------------------------------------------
class PygLinear(nn.Module):
    """torch_geometric.nn.dense.Linear(in, out, bias, weight_initializer=..., bias_initializer=None) of PyG 2.6.1:
    ``Linear.__init__`` ends with ``reset_parameters()``: weight by ``glorot`` or (weight_initializer=None)
    ``inits.kaiming_uniform(fan=in, a=sqrt 5)`` = U(+-1/sqrt(in)); bias by ``inits.uniform(in, bias)`` = U(+-1/sqrt(in))
    (NOT zeros).  The convs' own ``reset_parameters()`` call it a second time, so every Linear of a conv consumes the RNG
    twice - restated here so a seeded construction matches PyG draw for draw."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, weight_initializer: Optional[str] = "glorot"):
        super().__init__()
        self.in_channels, self.weight_initializer = in_channels, weight_initializer
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        bound = 1.0 / math.sqrt(self.in_channels)
        with torch.no_grad():
            if self.weight_initializer == "glorot":
                glorot_(self.weight)
            else:  # None / 'kaiming_uniform': sqrt(6 / ((1 + a^2) fan)) with a = sqrt(5)
                self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight, self.bias)


def gat_core(h: Tensor, s: Tensor, d: Tensor, edges: Tensor, negative_slope: float = 0.2) -> Tensor:
    """Attention aggregation of GATConv (heads=1) on an edge list that already has self loops.

    edges[0]=source j, edges[1]=target i (flow source_to_target).  out_i = sum_e p_e h_j.
    """
    src, dst = edges[0], edges[1]
    n = h.size(0)
    logit = F.leaky_relu(s.index_select(0, src) + d.index_select(0, dst), negative_slope)
    p = segment_softmax(logit, dst, n)
    msg = p.unsqueeze(-1) * h.index_select(0, src)
    return scatter(msg, dst, n, reduce="sum")


class GATConv(nn.Module):
    """GATConv(in, out) with all defaults (heads=1, concat, slope 0.2, dropout 0,
    add_self_loops, bias, no edge_dim, no residual).  Parameter names follow PyG >= 2.5:
    ``att_src[1,1,C]``, ``att_dst[1,1,C]``, ``bias[C]``, ``lin.weight[C,Cin]``."""

    def __init__(self, in_channels: int, out_channels: int, negative_slope: float = 0.2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.negative_slope = negative_slope
        self.lin = PygLinear(in_channels, out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        # PyG's constructor ends with reset_parameters(), which re-draws lin.weight (already
        # drawn once by Linear.__init__) before att_src / att_dst: keep the same RNG consumption.
        self.lin.reset_parameters()
        glorot_(self.att_src)
        glorot_(self.att_dst)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        h = self.lin(x)
        s = (h * self.att_src.view(-1)).sum(-1)
        d = (h * self.att_dst.view(-1)).sum(-1)
        edges = gat_edges(edge_index, x.size(0))
        return gat_core(h, s, d, edges, self.negative_slope) + self.bias


class GATv2Conv(nn.Module):
    """GATv2Conv(in, out) defaults: heads=1, share_weights=False, lin_l/lin_r with bias,
    att[1,1,C], add_self_loops, output bias."""

    def __init__(self, in_channels: int, out_channels: int, negative_slope: float = 0.2):
        super().__init__()
        self.negative_slope = negative_slope
        self.lin_l = PygLinear(in_channels, out_channels, bias=True)
        self.lin_r = PygLinear(in_channels, out_channels, bias=True)
        self.att = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin_l.reset_parameters()  # GATv2Conv.reset_parameters(): lin_l, lin_r, att, zeros(bias)
        self.lin_r.reset_parameters()
        glorot_(self.att)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        n = x.size(0)
        xl, xr = self.lin_l(x), self.lin_r(x)
        edges = gat_edges(edge_index, n)
        src, dst = edges[0], edges[1]
        e = F.leaky_relu(xl.index_select(0, src) + xr.index_select(0, dst), self.negative_slope)
        logit = (e * self.att.view(-1)).sum(-1)
        p = segment_softmax(logit, dst, n)
        out = scatter(p.unsqueeze(-1) * xl.index_select(0, src), dst, n, reduce="sum")
        return out + self.bias


class GCNConv(nn.Module):
    """GCNConv(in, out) defaults: add_remaining_self_loops(fill 1), symmetric normalisation,
    lin without bias (glorot), output bias (zeros)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin = PygLinear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin.reset_parameters()  # GCNConv.reset_parameters() draws lin a second time

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        n = x.size(0)
        # add_remaining_self_loops: existing loops are dropped from the list and re-added at the
        # end with weight 1 (edge_weight=None => all ones), so this equals gat_edges().
        edges = gat_edges(edge_index, n)
        src, dst = edges[0], edges[1]
        w = x.new_ones(edges.size(1))
        deg = scatter(w, dst, n, reduce="sum")
        dinv = deg.pow(-0.5)
        dinv = dinv.masked_fill(dinv == float("inf"), 0.0)
        norm = dinv.index_select(0, src) * w * dinv.index_select(0, dst)
        h = self.lin(x)
        out = scatter(norm.unsqueeze(-1) * h.index_select(0, src), dst, n, reduce="sum")
        return out + self.bias


class GraphConv(nn.Module):
    """GraphConv(in, out), aggr='add': lin_rel(sum_{j in N(i)} x_j) + lin_root(x_i);
    lin_rel has bias, lin_root has none; NO self loops are added."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        # PyG's GraphConv builds its Linears with the default initialisers (kaiming-uniform weight, uniform bias) and
        # GraphConv.reset_parameters() draws both again
        self.lin_rel = PygLinear(in_channels, out_channels, bias=True, weight_initializer=None)
        self.lin_root = PygLinear(in_channels, out_channels, bias=False, weight_initializer=None)
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        n = x.size(0)
        agg = scatter(x.index_select(0, edge_index[0]), edge_index[1], n, reduce="sum")
        return self.lin_rel(agg) + self.lin_root(x)


class GraphNorm(nn.Module):
    """tgnn.norm.GraphNorm(C, eps=1e-5).  The reference calls it WITHOUT ``batch``
    (models.py:73: bare module inside tgnn.Sequential) => one segment over all nodes."""

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def forward(self, x: Tensor, batch: Optional[Tensor] = None, batch_size: Optional[int] = None) -> Tensor:
        if batch is None:
            batch = x.new_zeros(x.size(0), dtype=torch.long)
            batch_size = 1
        if batch_size is None:
            batch_size = int(batch.max()) + 1
        mean = scatter(x, batch, batch_size, reduce="mean")
        out = x - mean.index_select(0, batch) * self.mean_scale
        var = scatter(out.pow(2), batch, batch_size, reduce="mean")
        std = (var + self.eps).sqrt().index_select(0, batch)
        return self.weight * out / std + self.bias


class Sequential(nn.Module):
    """tgnn.Sequential(input_args, modules): entries given as ``(module, "a, b -> c")`` are
    called with the named variables; bare modules are called with the previous output only.
    Children are registered as ``module_{i}``."""

    def __init__(self, input_args: str, modules: Seq[Union[nn.Module, Tuple[nn.Module, str]]]):
        super().__init__()
        self._input_args = [a.strip() for a in input_args.split(",")]
        self._calls: List[Tuple[str, Optional[List[str]], Optional[List[str]]]] = []
        for i, entry in enumerate(modules):
            name = f"module_{i}"
            if isinstance(entry, (tuple, list)):
                module, desc = entry
                lhs, rhs = desc.split("->")
                ins = [a.strip() for a in lhs.split(",")]
                outs = [a.strip() for a in rhs.split(",")]
            else:
                module, ins, outs = entry, None, None
            setattr(self, name, module)
            self._calls.append((name, ins, outs))

    def forward(self, *args, **kwargs):
        env: Dict[str, Any] = dict(zip(self._input_args, args))
        env.update(kwargs)
        last_names = [self._input_args[0]]
        for name, ins, outs in self._calls:
            module = getattr(self, name)
            if ins is None:
                ins, outs = last_names, last_names
            res = module(*[env[k] for k in ins])
            if len(outs) == 1:
                env[outs[0]] = res
            else:
                for k, v in zip(outs, res):
                    env[k] = v
            last_names = outs
        return env[last_names[0]] if len(last_names) == 1 else tuple(env[k] for k in last_names)


# --------------------------------------------------------------------------------------
# Data / Batch
# --------------------------------------------------------------------------------------
class Data:
    """Minimal torch_geometric.data.Data: an attribute bag; ``num_nodes`` inferred from ``x``."""

    def __init__(self, **kwargs):
        self.__dict__["_store"] = dict(kwargs)

    def __getattr__(self, key):
        store = self.__dict__["_store"]
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self.__dict__["_store"][key] = value

    def keys(self) -> List[str]:
        return list(self.__dict__["_store"].keys())

    @property
    def num_nodes(self) -> int:
        return int(self._store["x"].size(0))

    def cat_dim(self, key: str) -> int:
        return -1 if ("index" in key or key == "face") else 0

    def inc(self, key: str) -> int:
        return self.num_nodes if ("index" in key or key == "face") else 0

    def to(self, device):
        store = self.__dict__["_store"]
        for k, v in store.items():
            if isinstance(v, Tensor):
                store[k] = v.to(device)
        return self


class Batch(Data):
    """Minimal torch_geometric.data.Batch (from_data_list / get_example / num_graphs)."""

    @classmethod
    def from_data_list(cls, data_list: Seq[Data]) -> "Batch":
        data_list = list(data_list)
        out = cls()
        slices: Dict[str, List[int]] = {}
        incs: Dict[str, List[int]] = {}
        counts = [d.num_nodes for d in data_list]
        for key in data_list[0].keys():
            values = [getattr(d, key) for d in data_list]
            if isinstance(values[0], Tensor):
                dim = data_list[0].cat_dim(key)
                offs, run, cum, shifted = [0], 0, [], []
                for d, v in zip(data_list, values):
                    step = d.inc(key)
                    cum.append(run)
                    shifted.append(v + run if step and run else v)
                    run += step
                    offs.append(offs[-1] + v.size(dim))
                setattr(out, key, torch.cat(shifted, dim=dim))
                slices[key], incs[key] = offs, cum
            else:
                setattr(out, key, values)  # non-tensor attrs become a list per graph
                slices[key], incs[key] = list(range(len(values) + 1)), [0] * len(values)
        out.batch = torch.repeat_interleave(torch.arange(len(data_list)), torch.tensor(counts))
        ptr = torch.zeros(len(data_list) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor(counts), 0)
        out.ptr = ptr
        out.__dict__["_slices"], out.__dict__["_incs"] = slices, incs
        out.__dict__["_num_graphs"] = len(data_list)
        return out

    @property
    def num_graphs(self) -> int:
        return self.__dict__["_num_graphs"]

    def __getitem__(self, idx: int) -> Data:
        slices, incs = self.__dict__["_slices"], self.__dict__["_incs"]
        d = Data()
        for key, offs in slices.items():
            v = self._store[key]
            if isinstance(v, Tensor):
                dim = self.cat_dim(key)
                piece = v.narrow(dim if dim >= 0 else v.dim() + dim, offs[idx], offs[idx + 1] - offs[idx])
                if incs[key][idx]:
                    piece = piece - incs[key][idx]
                setattr(d, key, piece)
            else:
                setattr(d, key, v[idx])
        return d
