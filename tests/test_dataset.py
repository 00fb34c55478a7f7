"""the packed, memory-mappable dataset file (building_gan_b200/dataset.py).

Round trip of every per-graph tensor, and the shifted-per-graph-CSR collation against ``graph.collate_fn`` (= the
reference's ``Batch.from_data_list`` + ``bg_csr_build_host`` on the concatenated batch): bit-exact, CPU only."""
import os

import pytest
import torch

from building_gan_b200 import dataset, graph, synth


@pytest.fixture(scope="module")
def pack(tmp_path_factory):
    ids = list(range(40, 52))
    pairs = [synth.building_pair(i, shuffle=(i % 2 == 0)) for i in ids]
    path = os.path.join(tmp_path_factory.mktemp("pack"), "buildings.bgpack")
    dataset.write_pack(path, pairs)
    return pairs, dataset.PackedDataset(path)


def _same_data(a, b):
    assert a.keys() == b.keys()
    for k in a.keys():
        va, vb = getattr(a, k), getattr(b, k)
        if isinstance(va, torch.Tensor):
            assert va.dtype == vb.dtype and va.shape == vb.shape and torch.equal(va, vb), k
        else:
            assert va == vb, k


def test_round_trip_of_every_graph(pack):
    pairs, ds = pack
    assert len(ds) == len(pairs)
    for i, (l, v) in enumerate(pairs):
        l2, v2 = ds[i]
        _same_data(l, l2)
        _same_data(v, v2)
    with pytest.raises(IndexError):
        ds[len(pairs)]


@pytest.mark.parametrize("indices", [[0, 1, 2, 3], [7], [11, 3, 5, 0, 9], list(range(12))])
def test_collate_equals_reference_collation(pack, indices):
    pairs, ds = pack
    lb, vb = ds.collate(indices)
    rlb, rvb = graph.collate_fn([pairs[i] for i in indices])
    for got, ref in ((lb, rlb), (vb, rvb)):
        for k, v in ref._fields.items():
            if k == "bg_csr":
                continue
            g = got._fields[k]
            if isinstance(v, torch.Tensor):
                assert g.dtype == v.dtype and torch.equal(g, v), k
            else:
                assert g == v, k
        assert got.num_graphs == ref.num_graphs
        for gi in range(ref.num_graphs):
            _same_data(got[gi], ref[gi])
    a, b = vb.bg_csr, rvb.bg_csr
    assert (a.num_nodes, a.num_edges, a.num_graphs, a.max_deg, a.num_input_self_loops) == \
           (b.num_nodes, b.num_edges, b.num_graphs, b.max_deg, b.num_input_self_loops)
    for f in graph.VoxelCSR.FIELDS:
        assert getattr(a, f).dtype == torch.int32 and torch.equal(getattr(a, f), getattr(b, f)), f


def test_not_a_pack_is_refused(tmp_path):
    p = tmp_path / "x.bgpack"
    p.write_bytes(b"not a pack at all")
    with pytest.raises(ValueError, match="BGPACK01"):
        dataset.PackedDataset(str(p))


def test_on_disk_size_is_ragged_not_dense(pack):
    """The reference materialises a dense N x N adjacency per building while preprocessing (data.py:326); the pack stores
    O(N + E) integers per building."""
    pairs, ds = pack
    n2 = sum(4 * v.num_nodes ** 2 for _, v in pairs)
    assert os.path.getsize(ds.path) < n2


def test_pickling_does_not_copy_the_mapping(pack):
    """DataLoader workers started with spawn pickle the dataset: the memory map must not travel (it would be copied in
    full); the worker re-opens the file lazily."""
    import pickle
    pairs, ds = pack
    _ = ds[0]  # opens the mapping and fills the view cache
    blob = pickle.dumps(ds)
    assert len(blob) < 64 * 1024 + len(pickle.dumps(ds.header)), len(blob)
    ds2 = pickle.loads(blob)
    assert ds2._mm is None and ds2._cache == {}
    _same_data(ds2[1][1], pairs[1][1])
