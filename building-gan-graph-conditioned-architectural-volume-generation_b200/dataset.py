"""Packed, memory-mappable dataset of processed building graphs.

The reference keeps every processed building as two pickled ``Data`` objects on disk (``data.py:111-148``: ~12 000
``torch.load`` calls into RAM at start-up, ``weights_only=False`` required on torch >= 2.6) and stores the voxel adjacency
as a dense N x N matrix during preprocessing (``data.py:326-335``).  This module stores the same per-graph tensors of a
whole dataset in ONE file of ragged arrays,

    magic "BGPACK01" | u64 header length | JSON header | 64-byte aligned raw arrays

(node-level fields concatenated over the graphs with a ``node_ptr``; ``edge_index`` as int32 ``[E, 2]`` with an
``edge_ptr``; int64 fields narrowed to int32 / uint8 on disk and widened back on read), plus - for the voxel side - the
per-graph destination-sorted CSR / CSC / perm arrays ``bg_csr_build_host`` produces.  ``collate`` then builds the batch CSR
by SHIFTING the per-graph arrays (rowptr by the edge offset, col by the node offset): no sort at collation time, and the
result is bit-identical to ``graph.collate_fn`` on the same graphs (tests/test_dataset.py).

Reading is ``numpy.memmap``: opening a pack costs nothing, workers share the page cache, a batch touches only its rows.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from .graph import Batch, Data, VoxelCSR, _is_index_key

MAGIC = b"BGPACK01"
_ALIGN = 64
_CSR_FIELDS = ("rowptr", "col", "cscptr", "cscrow", "perm")


def _narrow(t: Tensor) -> np.ndarray:
    """int64 -> the narrowest of uint8 / int32 that holds the values (widened back to int64 on read)."""
    a = t.detach().cpu().numpy()
    if a.dtype == np.int64 and a.size:
        lo, hi = int(a.min()), int(a.max())
        if 0 <= lo and hi <= 255:
            return a.astype(np.uint8)
        if -(2 ** 31) <= lo and hi < 2 ** 31:
            return a.astype(np.int32)
    return a


class _Writer:
    def __init__(self):
        self.arrays: List[Tuple[str, np.ndarray]] = []

    def add(self, name: str, a: np.ndarray) -> Dict:
        a = np.ascontiguousarray(a)
        self.arrays.append((name, a))
        return {"name": name, "dtype": a.dtype.str, "shape": list(a.shape)}


def write_pack(path: str, pairs: Sequence[Tuple[Data, Data]]) -> None:
    """Write ``pairs`` = [(local Data, voxel Data), ...] (what ``GraphDataset.__getitem__`` / ``synth.building_pair`` return)."""
    pairs = list(pairs)
    w = _Writer()
    header: Dict = {"version": 1, "num_graphs": len(pairs), "sides": {}}
    for side, graphs in (("local", [p[0] for p in pairs]), ("voxel", [p[1] for p in pairs])):
        info: Dict = {"fields": {}, "lists": {}, "order": list(graphs[0].keys())}
        node_ptr = np.zeros(len(graphs) + 1, dtype=np.int64)
        node_ptr[1:] = np.cumsum([g.num_nodes for g in graphs])
        info["node_ptr"] = w.add(f"{side}.node_ptr", node_ptr)
        for key in graphs[0].keys():
            vals = [getattr(g, key) for g in graphs]
            if not isinstance(vals[0], Tensor):
                # per-node Python lists (data_number): one value per graph when constant inside the graph
                if all(isinstance(v, list) and len(set(v)) <= 1 for v in vals):
                    info["lists"][key] = {"per_graph": [v[0] if v else None for v in vals]}
                else:
                    info["lists"][key] = {"per_node": vals}
                continue
            if _is_index_key(key):
                ptr = np.zeros(len(graphs) + 1, dtype=np.int64)
                ptr[1:] = np.cumsum([v.shape[-1] for v in vals])
                cat = torch.cat([v.t() for v in vals], dim=0)  # [E_total, 2], graph-local indices
                meta = w.add(f"{side}.{key}", _narrow(cat))
                meta.update(index=True, orig_dtype=str(vals[0].dtype).replace("torch.", ""), ptr=w.add(f"{side}.{key}.ptr", ptr))
            else:
                assert all(v.shape[0] == g.num_nodes for v, g in zip(vals, graphs)), f"{side}.{key}: not a node-level field"
                meta = w.add(f"{side}.{key}", _narrow(torch.cat(vals, dim=0)))
                meta.update(index=False, orig_dtype=str(vals[0].dtype).replace("torch.", ""))
            info["fields"][key] = meta
        if side == "voxel":  # per-graph CSR / CSC / perm (bg_csr_build_host on each graph alone)
            csrs = [VoxelCSR.build(g.edge_index, g.num_nodes) for g in graphs]
            eptr = np.zeros(len(graphs) + 1, dtype=np.int64)
            eptr[1:] = np.cumsum([c.num_edges for c in csrs])
            info["csr"] = {"edge_ptr": w.add("voxel.csr.edge_ptr", eptr),
                           "max_deg": [int(c.max_deg) for c in csrs],
                           "self_loops": [int(c.num_input_self_loops) for c in csrs]}
            for f in _CSR_FIELDS:
                info["csr"][f] = w.add(f"voxel.csr.{f}", np.concatenate([getattr(c, f).numpy() for c in csrs]))
        header["sides"][side] = info
    # lay the arrays out after the header
    offsets, off = {}, 0
    for name, a in w.arrays:
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        offsets[name] = off
        off += a.nbytes
    header["offsets"] = offsets
    blob = json.dumps(header).encode()
    base = (len(MAGIC) + 8 + len(blob) + _ALIGN - 1) // _ALIGN * _ALIGN
    with open(path, "wb") as fh:
        fh.write(MAGIC + struct.pack("<Q", len(blob)) + blob)
        fh.write(b"\0" * (base - fh.tell()))
        for name, a in w.arrays:
            fh.write(b"\0" * (base + offsets[name] - fh.tell()))
            fh.write(a.tobytes())


class PackedDataset(torch.utils.data.Dataset):
    """Read side.  ``ds[i]`` -> (local Data, voxel Data) like the reference's ``GraphDataset.__getitem__``;
    ``ds.collate(indices)`` / ``ds.collate_fn(list of ds[i])`` -> (local Batch, voxel Batch with its CSR)."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as fh:
            if fh.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a BGPACK01 file")
            (hlen,) = struct.unpack("<Q", fh.read(8))
            self.header = json.loads(fh.read(hlen).decode())
        self._base = (len(MAGIC) + 8 + hlen + _ALIGN - 1) // _ALIGN * _ALIGN
        self._mm = None  # opened lazily, per process: see __getstate__
        self._cache: Dict[str, np.ndarray] = {}

    def _map(self) -> np.ndarray:
        if self._mm is None:
            self._mm = np.memmap(self.path, dtype=np.uint8, mode="r")
        return self._mm

    def __getstate__(self):
        """Pickled (DataLoader workers started with spawn / forkserver) WITHOUT the mapping and the array views: a worker
        re-opens the file on first use instead of receiving a full copy of it."""
        state = dict(self.__dict__)
        state["_mm"], state["_cache"] = None, {}
        return state

    def __len__(self) -> int:
        return int(self.header["num_graphs"])

    def _arr(self, meta: Dict) -> np.ndarray:
        name = meta["name"]
        a = self._cache.get(name)
        if a is None:
            dt = np.dtype(meta["dtype"])
            n = int(np.prod(meta["shape"])) if meta["shape"] else 1
            start = self._base + self.header["offsets"][name]
            a = self._map()[start:start + n * dt.itemsize].view(dt).reshape(meta["shape"])
            self._cache[name] = a
        return a

    @staticmethod
    def _tensor(a: np.ndarray, orig_dtype: str) -> Tensor:
        t = torch.from_numpy(np.array(a))  # copies the rows out of the mapping (a batch owns its memory)
        want = getattr(torch, orig_dtype)
        return t if t.dtype == want else t.to(want)

    def _side(self, side: str, i: int) -> Data:
        info = self.header["sides"][side]
        nptr = self._arr(info["node_ptr"])
        a, b = int(nptr[i]), int(nptr[i + 1])
        fields: Dict = {}
        for key in info["order"]:  # the attribute order of the stored Data objects is kept
            if key in info["lists"]:
                spec = info["lists"][key]
                fields[key] = [spec["per_graph"][i]] * (b - a) if "per_graph" in spec else spec["per_node"][i]
                continue
            meta = info["fields"][key]
            if meta["index"]:
                ptr = self._arr(meta["ptr"])
                rows = self._arr(meta)[int(ptr[i]):int(ptr[i + 1])]
                fields[key] = self._tensor(rows, meta["orig_dtype"]).t().contiguous()
            else:
                fields[key] = self._tensor(self._arr(meta)[a:b], meta["orig_dtype"])
        return Data(**fields)

    def __getitem__(self, i: int) -> Tuple[Data, Data]:
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._side("local", i), self._side("voxel", i)

    # -- collation -------------------------------------------------------------------------------
    def batch_csr(self, indices: Sequence[int]) -> VoxelCSR:
        """The batch's CSR / CSC / perm from the stored per-graph arrays: rowptr / cscptr / perm shifted by the batch's
        running edge count, col / cscrow by its running node count.  Equals ``VoxelCSR.build`` on the collated batch."""
        info = self.header["sides"]["voxel"]
        nptr, eptr = self._arr(info["node_ptr"]), self._arr(info["csr"]["edge_ptr"])
        arrs = {f: self._arr(info["csr"][f]) for f in _CSR_FIELDS}
        parts: Dict[str, List[np.ndarray]] = {f: [] for f in _CSR_FIELDS}
        noff = eoff = 0
        gptr = [0]
        # rowptr / cscptr are stored per graph WITH their trailing entry: graph i occupies [nptr[i] + i, nptr[i+1] + i + 1)
        for i in indices:
            n0, n1, e0, e1 = int(nptr[i]), int(nptr[i + 1]), int(eptr[i]), int(eptr[i + 1])
            parts["rowptr"].append(arrs["rowptr"][n0 + i:n1 + i] + eoff)
            parts["cscptr"].append(arrs["cscptr"][n0 + i:n1 + i] + eoff)
            parts["col"].append(arrs["col"][e0:e1] + noff)
            parts["cscrow"].append(arrs["cscrow"][e0:e1] + noff)
            parts["perm"].append(arrs["perm"][e0:e1] + eoff)
            noff += n1 - n0
            eoff += e1 - e0
            gptr.append(noff)
        last = np.array([eoff], dtype=np.int32)
        out = {f: torch.from_numpy(np.concatenate(parts[f] + ([last] if f in ("rowptr", "cscptr") else [])).astype(np.int32))
               for f in _CSR_FIELDS}
        md = info["csr"]["max_deg"]
        sl = info["csr"]["self_loops"]
        return VoxelCSR(noff, eoff, len(indices), max(md[i] for i in indices), sum(sl[i] for i in indices),
                        graph_ptr=torch.tensor(gptr, dtype=torch.int32), **out)

    def collate(self, indices: Sequence[int]) -> Tuple[Batch, Batch]:
        indices = [int(i) for i in indices]
        pairs = [self[i] for i in indices]
        lb = Batch.from_data_list([p[0] for p in pairs])
        vb = Batch.from_data_list([p[1] for p in pairs])
        vb._fields["bg_csr"] = self.batch_csr(indices)
        return lb, vb
