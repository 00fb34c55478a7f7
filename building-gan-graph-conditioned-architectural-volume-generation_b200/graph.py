"""Graph-batch objects and collation (hot-path row H1).

Replaces ``GraphDataset.collate_fn`` -> ``Batch.from_data_list`` x2 (reference data.py:156-163).
``Data`` / ``Batch`` are duck types of the torch_geometric classes for exactly the surface the
reference touches: attribute access, ``.num_nodes``, ``.num_graphs``, ``batch[gi]``,
``.to(device)``, list-of-lists ``data_number`` (models.py:122-146,230-242;
trainer.py:298,319,348-349,363-373,389,420-426,461-464).

New relative to the reference: the voxel batch additionally carries a ``VoxelCSR`` - the
destination-sorted CSR over the edges GATConv aggregates (input self loops stripped, one self
loop per node appended LAST, per-row order = COO order, i.e. the reference's CPU summation
order), its transpose (CSC) with the edge permutation the backward kernels need, and the graph
pointer.  It is built ONCE per batch on the host by ``bg_csr_build_host`` (include/bg_b200.h)
inside the DataLoader worker and uploaded with the batch.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import lib


def _is_index_key(key: str) -> bool:
    return "index" in key or key == "face"


class Data:
    """Attribute bag for one graph (stand-in for torch_geometric.data.Data)."""

    def __init__(self, **fields: Any):
        object.__setattr__(self, "_fields", dict(fields))

    def __getattr__(self, name: str) -> Any:
        try:
            return object.__getattribute__(self, "_fields")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        self._fields[name] = value

    def __contains__(self, name: str) -> bool:
        return name in self._fields

    def keys(self) -> List[str]:
        return list(self._fields)

    @property
    def num_nodes(self) -> int:
        return int(self._fields["x"].shape[0])

    def _moved(self, fn) -> None:
        for k, v in self._fields.items():
            if isinstance(v, Tensor):
                self._fields[k] = fn(v)

    def to(self, device, non_blocking: bool = False):
        self._moved(lambda t: t.to(device, non_blocking=non_blocking))
        return self

    def pin_memory(self):
        self._moved(lambda t: t.pin_memory())
        return self


class VoxelCSR:
    """Borrowed-pointer view handed to every aggregation kernel as ``BgGraph`` (include/bg_b200.h).

    rowptr[N+1], col[E']   in-edges of each destination row (source ids), self loop last
    cscptr[N+1], cscrow[E'], perm[E']   out-edges of each source (destination ids) and, for each,
                                        the position of the same edge in the CSR arrays
    graph_ptr[B+1]         node range of each building
    All int32.  ``max_deg`` = max in-degree incl. the self loop.
    """

    FIELDS = ("rowptr", "col", "cscptr", "cscrow", "perm", "graph_ptr")

    def __init__(self, num_nodes: int, num_edges: int, num_graphs: int, max_deg: int, num_input_self_loops: int = 0,
                 **arrays: Tensor):
        self.num_nodes, self.num_edges, self.num_graphs, self.max_deg = num_nodes, num_edges, num_graphs, max_deg
        self.num_input_self_loops = num_input_self_loops  # stripped by the build (GraphConv, which adds none, refuses them)
        for f in self.FIELDS:
            setattr(self, f, arrays[f])
        self._c = None

    @classmethod
    def build(cls, edge_index: Tensor, num_nodes: int, graph_ptr: Optional[Tensor] = None) -> "VoxelCSR":
        """Host-side build (bg_csr_build_host) from an int64 COO ``edge_index[2,E]`` on the CPU."""
        ei = edge_index.detach().to("cpu", torch.int64).contiguous()
        if graph_ptr is None:
            graph_ptr = torch.tensor([0, num_nodes])
        gp = graph_ptr.detach().to("cpu", torch.int32).contiguous()
        arrays, e_out, max_deg = lib.csr_build_host(ei, num_nodes)
        return cls(num_nodes, e_out, gp.numel() - 1, max_deg, int(ei.shape[1]) + num_nodes - e_out, graph_ptr=gp, **arrays)

    @property
    def device(self) -> torch.device:
        return self.rowptr.device

    def to(self, device, non_blocking: bool = False) -> "VoxelCSR":
        out = VoxelCSR(self.num_nodes, self.num_edges, self.num_graphs, self.max_deg, self.num_input_self_loops,
                       **{f: getattr(self, f).to(device, non_blocking=non_blocking) for f in self.FIELDS})
        return out

    def pin_memory(self) -> "VoxelCSR":
        return VoxelCSR(self.num_nodes, self.num_edges, self.num_graphs, self.max_deg, self.num_input_self_loops,
                        **{f: getattr(self, f).pin_memory() for f in self.FIELDS})

    def c_struct(self):
        """ctypes ``BgGraph`` (cached; pointers stay valid while this object is alive)."""
        if self._c is None:
            self._c = lib.make_bg_graph(self)
        return self._c


class Batch(Data):
    """Several graphs concatenated into one (stand-in for torch_geometric.data.Batch)."""

    @classmethod
    def from_data_list(cls, graphs: Sequence[Data], with_csr: bool = False) -> "Batch":
        graphs = list(graphs)
        sizes = [g.num_nodes for g in graphs]
        starts = [0]
        for s in sizes:
            starts.append(starts[-1] + s)
        out = cls()
        cuts: Dict[str, List[int]] = {}
        for key in graphs[0].keys():
            vals = [getattr(g, key) for g in graphs]
            if not isinstance(vals[0], Tensor):
                out._fields[key] = vals
                continue
            if _is_index_key(key):
                out._fields[key] = torch.cat([v + o if o else v for v, o in zip(vals, starts)], dim=-1)
                lens = [v.shape[-1] for v in vals]
            else:
                out._fields[key] = torch.cat(vals, dim=0)
                lens = [v.shape[0] for v in vals]
            acc = [0]
            for n in lens:
                acc.append(acc[-1] + n)
            cuts[key] = acc
        ptr = torch.tensor(starts, dtype=torch.long)
        out._fields["ptr"] = ptr
        out._fields["batch"] = torch.repeat_interleave(torch.arange(len(graphs)), torch.tensor(sizes))
        object.__setattr__(out, "_cuts", cuts)
        object.__setattr__(out, "_starts", starts)
        object.__setattr__(out, "_pack", None)
        if with_csr:
            out._fields["bg_csr"] = VoxelCSR.build(out._fields["edge_index"], starts[-1], ptr)
        return out

    @property
    def num_graphs(self) -> int:
        return len(object.__getattribute__(self, "_starts")) - 1

    def __getitem__(self, gi: int) -> Data:
        cuts = object.__getattribute__(self, "_cuts")
        starts = object.__getattribute__(self, "_starts")
        one = Data()
        for key, v in self._fields.items():
            if key in ("ptr", "batch", "bg_csr", "_bg_cache"):
                continue
            if not isinstance(v, Tensor):
                one._fields[key] = v[gi]
            elif _is_index_key(key):
                one._fields[key] = v[..., cuts[key][gi]:cuts[key][gi + 1]] - starts[gi]
            else:
                one._fields[key] = v[cuts[key][gi]:cuts[key][gi + 1]]
        return one

    def to(self, device, non_blocking: bool = False):
        """Moves every tensor field (and the CSR).  A batch packed by ``pin_memory()`` moves with ONE host->device copy
        of its pinned buffer (the reference issues ~25 separate copies per batch, trainer.py:461-462)."""
        pack = self.__dict__.get("_pack")
        if pack is not None and torch.device(device).type == "cuda":
            buf, index = pack
            dbuf = buf.to(device, non_blocking=True)
            for key, off, shape, dtype in index:
                n = 1
                for d in shape:
                    n *= d
                view = dbuf[off: off + n * dtype.itemsize].view(dtype).view(shape)
                if key.startswith("bg_csr."):
                    continue
                self._fields[key] = view
            if "bg_csr" in self._fields:
                old = self._fields["bg_csr"]
                arrays = {}
                for key, off, shape, dtype in index:
                    if key.startswith("bg_csr."):
                        n = 1
                        for d in shape:
                            n *= d
                        arrays[key[7:]] = dbuf[off: off + n * dtype.itemsize].view(dtype).view(shape)
                self._fields["bg_csr"] = VoxelCSR(old.num_nodes, old.num_edges, old.num_graphs, old.max_deg,
                                                  old.num_input_self_loops, **arrays)
            object.__setattr__(self, "_pack", None)
            self._fields.pop("_bg_cache", None)
            return self
        super().to(device, non_blocking=non_blocking)
        if "bg_csr" in self._fields:
            self._fields["bg_csr"] = self._fields["bg_csr"].to(device, non_blocking=non_blocking)
        self._fields.pop("_bg_cache", None)
        return self

    def pin_memory(self):
        """Packs every tensor field and the CSR arrays into ONE page-locked buffer (fields become views of it), so that
        ``.to('cuda')`` is a single asynchronous H2D copy.  ``DataLoader(pin_memory=True)`` calls this on the batch."""
        items = [(k, v) for k, v in self._fields.items() if isinstance(v, Tensor)]
        csr = self._fields.get("bg_csr")
        if csr is not None:
            items += [("bg_csr." + f, getattr(csr, f)) for f in VoxelCSR.FIELDS]
        index, off = [], 0
        for k, v in items:
            off = (off + 255) // 256 * 256
            index.append((k, off, tuple(v.shape), v.dtype))
            off += v.numel() * v.element_size()
        buf = torch.empty(max(off, 256), dtype=torch.uint8)
        try:
            buf = buf.pin_memory()
        except RuntimeError:
            pass  # no CUDA context (CPU-only process): the packed layout still gives the single-copy .to()
        for (k, o, shape, dtype), (_, v) in zip(index, items):
            n = v.numel() * v.element_size()
            view = buf[o: o + n].view(dtype).view(shape)
            view.copy_(v)
            if k.startswith("bg_csr."):
                setattr(csr, k[7:], view)
            else:
                self._fields[k] = view
        object.__setattr__(self, "_pack", (buf, index))
        return self


def collate_fn(pairs: Sequence[Tuple[Data, Data]]) -> Tuple[Batch, Batch]:
    """Drop-in for ``GraphDataset.collate_fn`` (data.py:156-163); the voxel batch also gets its CSR."""
    local_graphs, voxel_graphs = zip(*pairs)
    return Batch.from_data_list(local_graphs), Batch.from_data_list(voxel_graphs, with_csr=True)


def csr_of(voxel_graph) -> VoxelCSR:
    """The batch's CSR on the device of ``voxel_graph.x``; built lazily (one D2H + host build + H2D)
    when the batch came from a loader that did not attach it (e.g. real torch_geometric)."""
    csr = getattr(voxel_graph, "bg_csr", None) if not isinstance(voxel_graph, Data) else voxel_graph._fields.get("bg_csr")
    dev = voxel_graph.x.device
    if csr is None:
        ptr = getattr(voxel_graph, "ptr", None)
        csr = VoxelCSR.build(voxel_graph.edge_index, int(voxel_graph.x.shape[0]), ptr)
        try:
            voxel_graph.bg_csr = csr
        except Exception:
            pass
    if csr.device != dev:
        csr = csr.to(dev)
        try:
            voxel_graph.bg_csr = csr
        except Exception:
            pass
    return csr
