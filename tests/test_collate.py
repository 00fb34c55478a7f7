"""H1 collation: the host-built CSR / CSC / perm (bg_csr_build_host) against the oracle's COO edge list
(PyG remove_self_loops + add_self_loops) - integer work, bit-exact.  CPU only (no CUDA context)."""
import numpy as np
import pytest
import torch

from building_gan_b200 import graph, lib, synth
from oracle import pyg
from util import small_batch


def _check_csr(edge_index, n, csr):
    edges = pyg.gat_edges(edge_index, n)          # the oracle's edge list, in the order it scatters
    src, dst = edges[0].numpy(), edges[1].numpy()
    assert csr.num_edges == edges.shape[1]
    rowptr, col = csr.rowptr.numpy(), csr.col.numpy()
    order = np.argsort(dst, kind="stable")        # stable => COO order inside every destination row
    assert np.array_equal(col, src[order].astype(np.int32))
    assert np.array_equal(rowptr, np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=n))]).astype(np.int32))
    assert np.all(col[rowptr[1:] - 1] == np.arange(n)), "self loop must be the last entry of every row"
    # CSC + perm: perm maps each out-edge to the same edge in the CSR arrays
    cscptr, cscrow, perm = csr.cscptr.numpy(), csr.cscrow.numpy(), csr.perm.numpy()
    row_of = np.repeat(np.arange(n), np.diff(rowptr))
    assert np.array_equal(row_of[perm], cscrow)
    src_of_csc = np.repeat(np.arange(n), np.diff(cscptr))
    assert np.array_equal(col[perm], src_of_csc.astype(np.int32))
    assert sorted(perm.tolist()) == list(range(csr.num_edges))
    assert csr.max_deg == int(np.diff(rowptr).max())


@pytest.mark.parametrize("shuffle", [False, True])
def test_batched_csr_matches_oracle_coo(shuffle):
    _, vb = small_batch(ids=(31, 32, 33, 34), shuffle=shuffle)
    _check_csr(vb.edge_index, vb.num_nodes, vb.bg_csr)
    assert vb.bg_csr.max_deg <= 7                 # 6 face neighbours + self loop
    assert np.array_equal(vb.bg_csr.graph_ptr.numpy(), vb.ptr.numpy().astype(np.int32))


def test_csr_edge_cases():
    # input self loops are stripped, duplicates kept, isolated nodes get only their self loop, unsorted COO
    ei = torch.tensor([[3, 0, 0, 2, 2, 1, 0], [0, 0, 3, 2, 0, 0, 3]])
    csr = graph.VoxelCSR.build(ei, 5)
    _check_csr(ei, 5, csr)
    assert csr.col.tolist()[: csr.rowptr[1]] == [3, 2, 1, 0]      # row 0: COO order 3,2,1 then self loop
    empty = graph.VoxelCSR.build(torch.zeros(2, 0, dtype=torch.long), 4)
    assert empty.num_edges == 4 and empty.col.tolist() == [0, 1, 2, 3] and empty.max_deg == 1


def test_csr_rejects_bad_indices():
    with pytest.raises(RuntimeError, match="out of range"):
        graph.VoxelCSR.build(torch.tensor([[0, 9], [1, 0]]), 3)


def test_large_grid_shape():
    _, v = synth.large_grid_pair(0, floors=4, ny=20, nx=20)
    n = 4 * 20 * 20
    assert v.x.shape == (n, 12)
    assert v.edge_index.shape[1] == 2 * (3 * 400 + 4 * 19 * 20 * 2)
    csr = graph.VoxelCSR.build(v.edge_index, n)
    _check_csr(v.edge_index, n, csr)


def test_far_invariant_of_synthetic_buildings():
    """analyze.py:76-79: FAR == sum over non-void voxels of dim_y*dim_x / site_area."""
    for i in (1, 2, 3):
        g, _, v = synth.raw_building(i)
        gfa = sum(n["dimension"][1] * n["dimension"][2] for n in v["voxel_node"] if n["type"] >= 0)
        assert abs(g["far"] - gfa / g["site_area"]) < 1e-9
        assert abs(sum(x["proportion"] for x in g["global_node"]) - 1.0) < 1e-9


def test_csr_random_ragged_graphs_property():
    """Property test of bg_csr_build_host on random COO lists (ragged degrees, duplicate edges, input self loops, isolated
    nodes, a single node, hubs with hundreds of in-edges): every structural invariant of _check_csr holds and building the same
    list twice gives identical arrays (the counting sort is stable and has no data-dependent tie breaks)."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(n=st.integers(1, 40), e=st.integers(0, 400), hub=st.booleans(), seed=st.integers(0, 2**31 - 1))
    def run(n, e, hub, seed):
        rng = np.random.default_rng(seed)
        src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
        if hub and e:
            dst[rng.random(e) < 0.7] = rng.integers(0, n)  # most edges end in one node
        ei = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
        a, b = graph.VoxelCSR.build(ei, n), graph.VoxelCSR.build(ei.clone(), n)
        _check_csr(ei, n, a)
        for name in ("rowptr", "col", "cscptr", "cscrow", "perm"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        assert a.num_edges == int((src != dst).sum()) + n  # input self loops stripped, one appended per node

    run()


def test_collate_single_graph_and_order_independence():
    """A batch of one building collates to that building's own arrays (offsets 0), and the CSR of a batch equals the per-graph
    CSRs shifted by the cumulative node counts (graphs are disjoint components: data.py:160-161)."""
    pairs = [synth.building_pair(i) for i in (51, 52, 53)]
    _, one = graph.collate_fn(pairs[:1])
    assert one.ptr.tolist() == [0, one.num_nodes] and torch.equal(one.edge_index, pairs[0][1].edge_index)
    _, vb = graph.collate_fn(pairs)
    off, rp, col = 0, [0], []
    for _, v in pairs:
        c = graph.VoxelCSR.build(v.edge_index, v.num_nodes)
        rp += (c.rowptr[1:] + rp[-1]).tolist()
        col += (c.col + off).tolist()
        off += v.num_nodes
    assert vb.bg_csr.rowptr.tolist() == rp and vb.bg_csr.col.tolist() == col
