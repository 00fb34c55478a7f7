"""Synthetic "6types-like" buildings in the reference's own schemas.

The real dataset (building_gan/data/6types-raw_data-10000.zip) is a Git-LFS pointer, so the
benchmark and the tests run on synthetic buildings that follow

* the RAW JSON schema consumed by ``DataCreatorHelper.process_data`` (data.py:215-391):
  global ``{far, site_area, global_node:[{type, proportion}]}``, local ``{node:[{floor, type,
  type_id, center, neighbors}]}``, voxel ``{voxel_node:[{location, coordinate, dimension, type,
  neighbors}]}``;
* the PROCESSED feature layout of ``LocalGraphData`` / ``VoxelGraphData`` (data.py:16-77) and the
  ``Data`` fields of ``GraphDataset`` (data.py:118-147);
* the dataset statistics recorded in analyze.py:99-110 (type histogram, value ranges) and the
  invariant ``far == sum_{non-void}(dim_y*dim_x)/site_area`` (analyze.py:76-79).

``raw_building`` emits the JSON dicts, ``process_raw`` turns them into feature tensors (a
vectorised equivalent of data.py:215-391 that never materialises the dense N x N adjacency);
``building_fields`` / ``building_fields_fast`` return the (local, voxel) FIELD DICTS of one building -
the keyword arguments of a ``Data`` object, in the reference's attribute order.  ``grid_arrays`` is
the shared generator; the 1e5-voxel graphs of BASELINE config 4 come from ``large_grid_fields``.

Neutral workload code: depends on numpy + torch only, imports neither the product package
(``building_gan_b200``) nor ``oracle/``.  Both benchmark arms and the tests wrap the SAME field dicts in
their own ``Data`` / ``Batch`` classes (``building_gan_b200.synth`` / ``oracle.pyg``), so the reference
arm of ``bench.py`` runs without touching the product.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch



class Configuration:
    """The constants of the reference's ``Configuration`` that preprocessing reads (config.py:9-45); any object with
    these attributes (the reference's own, ``building_gan_b200.Configuration``, ``oracle.config.Configuration``) works."""
    NUM_CLASSES = 7
    VOID, VOID_OLD = 6, -1
    NORMALIZATION_FACTOR_FLOOR_LEVEL = 10
    NORMALIZATION_FACTOR_DIMENSION = 11
    NORMALIZATION_FACTOR_LOCATION = 11
    NORMALIZATION_FACTOR_COORDINATE = 42
    NORMALIZATION_FACTOR_SITE = 1600


# analyze.py:100 - void, lobby, restroom, stairs, elevator, office, mechanical (percent)
_HIST = {6: 33.65, 0: 13.10, 1: 6.35, 2: 2.75, 3: 4.95, 4: 38.09, 5: 1.12}
_NONVOID = np.array([0, 1, 2, 3, 4, 5])
_NONVOID_P = np.array([_HIST[t] for t in _NONVOID]) / sum(_HIST[t] for t in _NONVOID)


def _axis_widths(rng: np.random.Generator, cells: int, lo: int, hi: int, cap: int) -> np.ndarray:
    """``cells`` integer widths in [lo, hi] whose sum stays <= cap (coordinates <= 42, analyze.py:107)."""
    w = rng.integers(lo, hi + 1, size=cells)
    while w.sum() > cap:
        k = int(np.argmax(w))
        if w[k] <= lo:
            break
        w[k] -= 1
    return w


def grid_arrays(building_id: int, floors: int = 0, ny: int = 0, nx: int = 0, shuffle: bool = False) -> Dict[str, np.ndarray]:
    """One irregular F x Y x X voxel grid (non-uniform cell widths per axis, 6-neighbour faces)."""
    rng = np.random.default_rng(777 + int(building_id))
    F = floors or int(rng.integers(2, 12))
    Y = ny or int(rng.integers(4, 13))
    X = nx or int(rng.integers(4, 13))
    cap = 42 if max(F, Y, X) <= 12 else 10 ** 9
    dz = _axis_widths(rng, F, 3, 4, cap)
    dy = _axis_widths(rng, Y, 3, 11, cap)
    dx = _axis_widths(rng, X, 3, 11, cap)
    cz, cy, cx = (np.concatenate([[0], np.cumsum(w)[:-1]]) for w in (dz, dy, dx))
    f, y, x = np.meshgrid(np.arange(F), np.arange(Y), np.arange(X), indexing="ij")
    f, y, x = f.ravel(), y.ravel(), x.ravel()
    n = f.size
    # occupied footprint: a rectangle that steps back on upper floors; the rest is void
    y0, x0 = int(rng.integers(0, max(1, Y // 4) + 1)), int(rng.integers(0, max(1, X // 4) + 1))
    y1 = Y - int(rng.integers(0, max(1, Y // 4) + 1))
    x1 = X - int(rng.integers(0, max(1, X // 4) + 1))
    setback = (f * rng.uniform(0.0, 0.25)).astype(np.int64)
    inside = (y >= y0 + setback) & (y < y1 - setback) & (x >= x0) & (x < x1)
    # one program per vertical column block, re-drawn every few floors
    col_block = (y // 2) * ((X + 1) // 2) + (x // 2) + (f // 3) * 1009
    _, blk = np.unique(col_block, return_inverse=True)
    blk_type = rng.choice(_NONVOID, size=blk.max() + 1, p=_NONVOID_P)
    vtype = np.where(inside, blk_type[blk], -1)
    order = rng.permutation(n) if shuffle else np.arange(n)
    return dict(
        F=np.int64(F), Y=np.int64(Y), X=np.int64(X),
        location=np.stack([f, y, x], 1)[order],
        coordinate=np.stack([cz[f], cy[y], cx[x]], 1)[order].astype(np.float64),
        dimension=np.stack([dz[f], dy[y], dx[x]], 1)[order].astype(np.float64),
        type=vtype[order].astype(np.int64),
        plan=np.array([dy.sum(), dx.sum()], dtype=np.int64),
    )


def voxel_count(building_id: int) -> int:
    """Number of voxels of ``grid_arrays(building_id)`` without building it (the first three draws of its generator)."""
    rng = np.random.default_rng(777 + int(building_id))
    return int(rng.integers(2, 12)) * int(rng.integers(4, 13)) * int(rng.integers(4, 13))


def _neighbour_pairs(loc: np.ndarray, F: int, Y: int, X: int) -> Tuple[np.ndarray, np.ndarray]:
    """Directed face-adjacency (src, dst) pairs in node-id space, sorted src-major then dst - the
    order ``adjacency.nonzero().t()`` produces (data.py:326-335)."""
    n = loc.shape[0]
    lin = (loc[:, 0] * Y + loc[:, 1]) * X + loc[:, 2]
    where = np.full(F * Y * X, -1, dtype=np.int64)
    where[lin] = np.arange(n)
    srcs, dsts = [], []
    for axis, size in ((0, F), (1, Y), (2, X)):
        for step in (-1, 1):
            ok = (loc[:, axis] + step >= 0) & (loc[:, axis] + step < size)
            nb = loc[ok].copy()
            nb[:, axis] += step
            tgt = where[(nb[:, 0] * Y + nb[:, 1]) * X + nb[:, 2]]
            keep = tgt >= 0
            srcs.append(np.nonzero(ok)[0][keep])
            dsts.append(tgt[keep])
    src, dst = np.concatenate(srcs), np.concatenate(dsts)
    key = np.unique(src * n + dst)
    return key // n, key % n


def raw_building(building_id: int, **grid_kw) -> Tuple[dict, dict, dict]:
    """(global, local, voxel) raw-JSON dicts for one synthetic building."""
    g = grid_arrays(building_id, **grid_kw)
    F, Y, X = int(g["F"]), int(g["Y"]), int(g["X"])
    loc, typ = g["location"], g["type"]
    src, dst = _neighbour_pairs(loc, F, Y, X)
    nbrs: List[List[List[int]]] = [[] for _ in range(loc.shape[0])]
    for s, d in zip(src.tolist(), dst.tolist()):
        nbrs[s].append(loc[d].tolist())
    voxel = {
        "voxel_node": [
            {
                "location": loc[i].tolist(),
                "coordinate": g["coordinate"][i].tolist(),
                "dimension": g["dimension"][i].tolist(),
                "type": int(typ[i]),
                "neighbors": nbrs[i],
            }
            for i in range(loc.shape[0])
        ]
    }
    glob = _global_record(g)
    return glob, {"node": _local_nodes(g)}, voxel


def _global_record(g: Dict[str, np.ndarray]) -> dict:
    typ = g["type"]
    site_area = int(min(1600, max(324, int(g["plan"][0]) * int(g["plan"][1]))))
    solid = typ >= 0
    area = g["dimension"][:, 1] * g["dimension"][:, 2]
    gfa = float(area[solid].sum())
    area_by_type = np.array([area[typ == t].sum() for t in range(6)])
    prop = area_by_type / max(area_by_type.sum(), 1.0)
    return {
        "far": gfa / site_area,
        "site_area": site_area,
        "global_node": [{"type": t, "proportion": float(prop[t])} for t in range(6) if prop[t] > 0],
    }


def _local_nodes(g: Dict[str, np.ndarray]) -> List[dict]:
    """Program graph: one node per (floor, type); same-floor nodes are chained, same-type nodes of consecutive
    floors are linked.  models.py never reads local edges (only trainer.py:106 plots them)."""
    F = int(g["F"])
    loc, typ = g["location"], g["type"]
    mid = g["coordinate"] + g["dimension"] / 2
    nodes, index = [], {}
    for f in range(F):
        on_floor = loc[:, 0] == f
        for t in range(6):
            sel = on_floor & (typ == t)
            if sel.any():
                index[(f, t, 0)] = len(nodes)
                nodes.append({"floor": f, "type": t, "type_id": 0, "center": mid[sel].mean(0).tolist(), "neighbors": []})
    keys = list(index.keys())
    by_floor: Dict[int, List[int]] = {}
    for k in keys:
        by_floor.setdefault(k[0], []).append(k[1])
    for ka in keys:
        nxt = [t for t in by_floor[ka[0]] if t > ka[1]]
        links = []
        if nxt:
            links.append((ka[0], min(nxt), 0))
        if (ka[0] + 1, ka[1], 0) in index:
            links.append((ka[0] + 1, ka[1], 0))
        for kb in links:
            nodes[index[ka]]["neighbors"].append(list(kb))
            nodes[index[kb]]["neighbors"].append(list(ka))
    return nodes


def process_raw(glob: dict, local: dict, voxel: dict, cfg=Configuration, data_number: str = "000000"):
    """Raw JSON -> (local fields, voxel fields) with the reference's processed layout
    (data.py:215-391 + LocalGraphData/VoxelGraphData data.py:16-77 + Data fields data.py:118-147)."""
    K = cfg.NUM_CLASSES
    far = torch.tensor([glob["far"]])
    site = torch.tensor([glob["site_area"]])
    site_n = site / cfg.NORMALIZATION_FACTOR_SITE
    ratio = [0] * K
    for gn in glob["global_node"]:
        ratio[gn["type"]] = gn["proportion"]
    ratio = torch.tensor(ratio)

    ln = local["node"]
    m = len(ln)
    l_floor = torch.tensor([v["floor"] for v in ln])
    l_type = torch.tensor([v["type"] for v in ln])
    l_tid = torch.tensor([v["type_id"] for v in ln])
    l_center = torch.tensor([v["center"] for v in ln])
    l_onehot = torch.nn.functional.one_hot(l_type, num_classes=K)
    l_ratio = l_onehot * ratio
    lookup = {(v["floor"], v["type"], v["type_id"]): i for i, v in enumerate(ln)}
    pairs = sorted({(lookup[(v["floor"], v["type"], v["type_id"])], lookup[tuple(nb)]) for v in ln for nb in v["neighbors"]})
    l_edges = torch.tensor(pairs, dtype=torch.long).reshape(-1, 2).t().contiguous()
    l_x = torch.cat(
        [l_onehot, l_ratio, torch.zeros(m, 1) + far, (l_floor / cfg.NORMALIZATION_FACTOR_FLOOR_LEVEL).unsqueeze(1),
         site_n.repeat(m).unsqueeze(1)], dim=1)
    local_fields = dict(
        x=l_x, edge_index=l_edges, node_cluster=l_type.clone(), node_ratio=l_ratio, types_onehot=l_onehot,
        center=l_center, type=l_type, type_id=l_tid, floor=l_floor, data_number=[data_number] * m,
        site_area=site.repeat(m),
    )

    vn = voxel["voxel_node"]
    n = len(vn)
    v_loc = torch.tensor([v["location"] for v in vn])
    v_coord = torch.tensor([v["coordinate"] for v in vn])
    v_dim = torch.tensor([v["dimension"] for v in vn])
    v_type = torch.tensor([cfg.VOID if v["type"] == cfg.VOID_OLD else v["type"] for v in vn])
    v_onehot = torch.nn.functional.one_hot(v_type, num_classes=K)
    feats = torch.tensor(
        [[*(c / cfg.NORMALIZATION_FACTOR_COORDINATE for c in v["coordinate"]),
          *(d / cfg.NORMALIZATION_FACTOR_DIMENSION for d in v["dimension"]),
          *(q / cfg.NORMALIZATION_FACTOR_LOCATION for q in v["location"])] for v in vn])
    vlookup = {tuple(v["location"]): i for i, v in enumerate(vn)}
    vpairs = sorted({(i, vlookup[tuple(nb)]) for i, v in enumerate(vn) for nb in v["neighbors"]})
    v_edges = torch.tensor(vpairs, dtype=torch.long).reshape(-1, 2).t().contiguous()
    v_floor = v_loc[:, 0].clone()
    counts = torch.bincount(v_type, minlength=K) / n
    v_x = torch.cat(
        [feats, torch.zeros(n, 1) + far, (v_floor / cfg.NORMALIZATION_FACTOR_FLOOR_LEVEL).unsqueeze(1),
         site_n.repeat(n).unsqueeze(1)], dim=1)
    voxel_fields = dict(
        x=v_x, edge_index=v_edges, voxel_level=v_floor, type=v_type, types_onehot=v_onehot, coordinate=v_coord,
        dimension=v_dim, location=v_loc, node_ratio=(v_onehot * counts).max(dim=1)[0].unsqueeze(1),
        data_number=[data_number] * n, site_area=site.repeat(n),
    )
    return local_fields, voxel_fields


def building_fields(building_id: int, cfg=Configuration, **grid_kw) -> Tuple[dict, dict]:
    """(local fields, voxel fields) of one synthetic building through the raw-JSON schema - what ``GraphDataset[i]``
    holds (data.py:118-147)."""
    return process_raw(*raw_building(building_id, **grid_kw), cfg=cfg, data_number=f"{building_id:06d}")


def building_fields_fast(building_id: int, cfg=Configuration, **grid_kw) -> Tuple[dict, dict]:
    """Same field dicts as ``building_fields`` - bit for bit, tests/test_collate.py - but built straight
    from arrays: the JSON detour of ``raw_building`` costs O(N) Python objects (0.1-0.2 s per building)."""
    g = grid_arrays(building_id, **grid_kw)
    F, Y, X = int(g["F"]), int(g["Y"]), int(g["X"])
    K = cfg.NUM_CLASSES
    number = f"{building_id:06d}"
    glob = _global_record(g)
    far = torch.tensor([glob["far"]])
    site = torch.tensor([glob["site_area"]])
    site_n = site / cfg.NORMALIZATION_FACTOR_SITE
    ratio = [0] * K
    for gn in glob["global_node"]:
        ratio[gn["type"]] = gn["proportion"]
    ratio = torch.tensor(ratio)

    src, dst = _neighbour_pairs(g["location"], F, Y, X)
    n = g["location"].shape[0]
    typ = torch.from_numpy(np.where(g["type"] < 0, cfg.VOID, g["type"]))
    onehot = torch.nn.functional.one_hot(typ, num_classes=K)
    loc = torch.from_numpy(g["location"])
    feats = torch.from_numpy(np.concatenate(
        [g["coordinate"] / cfg.NORMALIZATION_FACTOR_COORDINATE, g["dimension"] / cfg.NORMALIZATION_FACTOR_DIMENSION,
         g["location"] / cfg.NORMALIZATION_FACTOR_LOCATION], 1)).float()
    floor = loc[:, 0].clone()
    counts = torch.bincount(typ, minlength=K) / n
    vx = torch.cat([feats, torch.zeros(n, 1) + far, (floor / cfg.NORMALIZATION_FACTOR_FLOOR_LEVEL).unsqueeze(1),
                    site_n.repeat(n).unsqueeze(1)], 1)
    voxel = dict(
        x=vx, edge_index=torch.from_numpy(np.stack([src, dst])), voxel_level=floor, type=typ, types_onehot=onehot,
        coordinate=torch.from_numpy(g["coordinate"]).float(), dimension=torch.from_numpy(g["dimension"]).float(),
        location=loc, node_ratio=(onehot * counts).max(dim=1)[0].unsqueeze(1), data_number=[number] * n,
        site_area=site.repeat(n))

    ln = _local_nodes(g)
    m = len(ln)
    l_floor = torch.tensor([v["floor"] for v in ln])
    l_type = torch.tensor([v["type"] for v in ln])
    l_onehot = torch.nn.functional.one_hot(l_type, num_classes=K)
    l_ratio = l_onehot * ratio
    lookup = {(v["floor"], v["type"], v["type_id"]): i for i, v in enumerate(ln)}
    pairs = sorted({(lookup[(v["floor"], v["type"], v["type_id"])], lookup[tuple(nb)]) for v in ln for nb in v["neighbors"]})
    lx = torch.cat([l_onehot, l_ratio, torch.zeros(m, 1) + far, (l_floor / cfg.NORMALIZATION_FACTOR_FLOOR_LEVEL).unsqueeze(1),
                    site_n.repeat(m).unsqueeze(1)], 1)
    local = dict(
        x=lx, edge_index=torch.tensor(pairs, dtype=torch.long).reshape(-1, 2).t().contiguous(), node_cluster=l_type.clone(),
        node_ratio=l_ratio, types_onehot=l_onehot, center=torch.tensor([v["center"] for v in ln]), type=l_type,
        type_id=torch.tensor([v["type_id"] for v in ln]), floor=l_floor, data_number=[number] * m, site_area=site.repeat(m))
    return local, voxel


def large_grid_fields(building_id: int, floors: int = 10, ny: int = 100, nx: int = 100, cfg=Configuration) -> Tuple[dict, dict]:
    """BASELINE config 4: one F x Y x X irregular grid (default 1e5 voxels, 576 000 directed edges)."""
    return building_fields_fast(building_id, cfg=cfg, floors=floors, ny=ny, nx=nx)
