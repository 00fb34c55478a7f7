"""Self-tests of ``oracle/pyg.py`` - the restatement of the torch-geometric 2.6.1 ops the reference calls
(models.py:22-29,72-75,90,166-173,192-195,210; data.py:160-161; trainer.py:364,421).

torch_geometric is not installable in the build container and the reference ships no golden vectors for it
(SURVEY section 8c), so these are the pins the PyG semantics have:

(i)   dense-adjacency DEFINITIONS of every conv, written without scatter / index_select (row-by-row softmax over the
      in-neighbours incl. the self loop; ``D^-1/2 (A+I) D^-1/2``; ``A X W_rel^T + X W_root^T``),
(ii)  hand-computed 3-node known answers (numbers derived on paper with scalar arithmetic, literals below),
(iii) fp64 ``gradcheck`` / ``gradgradcheck`` (the WGAN-GP path differentiates the discriminator's convs twice),
(iv)  the ``tgnn.Sequential`` call convention (bare modules get the previous output only => ``GraphNorm(x)`` with
      ``batch=None``), child naming, and the ``Batch.from_data_list`` / ``Batch.__getitem__`` round trip,
(v)   initialisers and RNG consumption of every conv (PyG draws each conv's Linears twice).
"""
from __future__ import annotations

import math

import pytest
import torch
from torch import nn

from oracle import pyg

torch.set_default_dtype(torch.float32)


def _graph(n: int, e: int, seed: int, self_loops: int = 2):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n, (e,), generator=g)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    # unique directed edges (the reference's edge lists come from nonzero() of an adjacency: no multi-edges)
    key = torch.unique(src * n + dst)
    src, dst = key // n, key % n
    loops = torch.randint(0, n, (self_loops,), generator=g)  # input self loops: GATConv/GCNConv must strip them
    return torch.stack([torch.cat([src, loops]), torch.cat([dst, loops])])


def _dense_adj(edge_index, n, with_self=True):
    """A[i, j] = 1 iff there is an edge j -> i (self loops stripped, then exactly one per node if ``with_self``)."""
    A = torch.zeros(n, n, dtype=torch.float64)
    for j, i in edge_index.t().tolist():
        if i != j:
            A[i, j] = 1.0
    if with_self:
        A += torch.eye(n, dtype=torch.float64)
    return A


def _dbl(m: nn.Module) -> nn.Module:
    return m.double()


# ------------------------------------------------------------------------------------------------------------
# (i) dense definitions
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,e,cin,c", [(9, 30, 5, 4), (17, 40, 3, 8), (6, 0, 2, 1)])
def test_gatconv_equals_dense_row_softmax(n, e, cin, c):
    torch.manual_seed(n)
    ei = _graph(n, max(e, 1), seed=n + e)
    conv = _dbl(pyg.GATConv(cin, c))
    with torch.no_grad():
        conv.bias.normal_()
    x = torch.randn(n, cin, dtype=torch.float64)
    A = _dense_adj(ei, n)
    h = x @ conv.lin.weight.t()
    s, d = h @ conv.att_src.view(-1), h @ conv.att_dst.view(-1)
    want = torch.zeros(n, c, dtype=torch.float64)
    for i in range(n):
        nb = [j for j in range(n) if A[i, j] > 0]
        logit = torch.stack([nn.functional.leaky_relu(s[j] + d[i], 0.2) for j in nb])
        p = torch.softmax(logit, 0)
        want[i] = sum(p[k] * h[j] for k, j in enumerate(nb)) + conv.bias
    got = conv(x, ei)
    assert torch.allclose(got, want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("n,e,cin,c", [(9, 30, 5, 4), (12, 25, 4, 2)])
def test_gcnconv_equals_dense_sym_norm(n, e, cin, c):
    torch.manual_seed(1)
    ei = _graph(n, e, seed=3)
    conv = _dbl(pyg.GCNConv(cin, c))
    with torch.no_grad():
        conv.bias.normal_()
    x = torch.randn(n, cin, dtype=torch.float64)
    A = _dense_adj(ei, n)                                 # A + I
    dinv = A.sum(1).pow(-0.5)                             # in-degree incl. the self loop
    want = (dinv[:, None] * A * dinv[None, :]) @ (x @ conv.lin.weight.t()) + conv.bias
    assert torch.allclose(conv(x, ei), want, rtol=0, atol=1e-12)


def test_graphconv_equals_dense_neighbour_sum_without_self_loops():
    torch.manual_seed(2)
    n, cin, c = 10, 3, 5
    ei = _graph(n, 30, seed=5, self_loops=0)
    conv = _dbl(pyg.GraphConv(cin, c))
    x = torch.randn(n, cin, dtype=torch.float64)
    A = _dense_adj(ei, n, with_self=False)
    want = (A @ x) @ conv.lin_rel.weight.t() + conv.lin_rel.bias + x @ conv.lin_root.weight.t()
    assert torch.allclose(conv(x, ei), want, rtol=0, atol=1e-12)


def test_gatv2conv_equals_dense_definition():
    torch.manual_seed(4)
    n, cin, c = 8, 4, 3
    ei = _graph(n, 24, seed=8)
    conv = _dbl(pyg.GATv2Conv(cin, c))
    x = torch.randn(n, cin, dtype=torch.float64)
    A = _dense_adj(ei, n)
    xl = x @ conv.lin_l.weight.t() + conv.lin_l.bias
    xr = x @ conv.lin_r.weight.t() + conv.lin_r.bias
    want = torch.zeros(n, c, dtype=torch.float64)
    for i in range(n):
        nb = [j for j in range(n) if A[i, j] > 0]
        logit = torch.stack([(nn.functional.leaky_relu(xl[j] + xr[i], 0.2) * conv.att.view(-1)).sum() for j in nb])
        p = torch.softmax(logit, 0)
        want[i] = sum(p[k] * xl[j] for k, j in enumerate(nb)) + conv.bias
    assert torch.allclose(conv(x, ei), want, rtol=0, atol=1e-12)


def test_graphnorm_equals_plain_formula_and_segments():
    torch.manual_seed(5)
    n, c = 11, 6
    gn = _dbl(pyg.GraphNorm(c))
    with torch.no_grad():
        gn.weight.normal_(), gn.bias.normal_(), gn.mean_scale.normal_()
    x = torch.randn(n, c, dtype=torch.float64)
    o = x - x.mean(0) * gn.mean_scale
    want = gn.weight * o / (o.pow(2).mean(0) + 1e-5).sqrt() + gn.bias
    assert torch.allclose(gn(x), want, rtol=0, atol=1e-13)          # batch=None: ONE segment over all rows
    batch = torch.tensor([0] * 4 + [1] * 7)
    got = gn(x, batch)
    for lo, hi in ((0, 4), (4, 11)):
        seg = x[lo:hi]
        o = seg - seg.mean(0) * gn.mean_scale
        assert torch.allclose(got[lo:hi], gn.weight * o / (o.pow(2).mean(0) + 1e-5).sqrt() + gn.bias, rtol=0, atol=1e-13)


# ------------------------------------------------------------------------------------------------------------
# (ii) hand-computed 3-node known answers
# ------------------------------------------------------------------------------------------------------------
def test_gatconv_three_node_known_answer():
    """x = [1,2,3], W = [[2]], att_src = 0.5, att_dst = -1, bias = 0.1; edges 0->1, 2->1, 1->0 plus a stray input self
    loop 2->2 (must be replaced by exactly one).  h = [2,4,6], s = [1,2,3], d = [-2,-4,-6].
    node 0: logits LReLU(s1+d0)=0, LReLU(s0+d0)=-0.2     -> p = [0.5498339973, 0.4501660027] -> 0.549834*4 + 0.450166*2 + 0.1
    node 1: logits -0.6 (0->1), -0.2 (2->1), -0.4 (self) -> p = [0.2693074992, 0.4017595785, 0.3289329223]
    node 2: only its self loop -> p = 1 -> 6 + 0.1."""
    conv = _dbl(pyg.GATConv(1, 1))
    with torch.no_grad():
        conv.lin.weight.fill_(2.0), conv.att_src.fill_(0.5), conv.att_dst.fill_(-1.0), conv.bias.fill_(0.1)
    x = torch.tensor([[1.0], [2.0], [3.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 2, 1, 2], [1, 1, 0, 2]])
    got = conv(x, ei).view(-1)
    want = torch.tensor([3.199667994624956, 4.3649041587112345, 6.1], dtype=torch.float64)
    assert torch.allclose(got, want, rtol=0, atol=1e-13)
    # the edge list the aggregation runs over: input self loop stripped, one self loop per node APPENDED LAST
    assert pyg.gat_edges(ei, 3).tolist() == [[0, 2, 1, 0, 1, 2], [1, 1, 0, 0, 1, 2]]


def test_graphnorm_three_node_known_answer():
    """x = [[1,2],[3,6],[5,10]], weight = [2,1], bias = [0.5,-1], mean_scale = [1,0.5], batch=None.
    channel 0: mean 3, o = [-2,0,2], var = 8/3   -> y = 2 o / sqrt(8/3 + 1e-5) + 0.5
    channel 1: mean 6, o = x - 3 = [-1,3,7], var = 59/3 (NOT re-centred: mean_scale = 0.5) -> y = o / sqrt(59/3 + 1e-5) - 1."""
    gn = _dbl(pyg.GraphNorm(2))
    with torch.no_grad():
        gn.weight.copy_(torch.tensor([2.0, 1.0])), gn.bias.copy_(torch.tensor([0.5, -1.0]))
        gn.mean_scale.copy_(torch.tensor([1.0, 0.5]))
    x = torch.tensor([[1.0, 2.0], [3.0, 6.0], [5.0, 10.0]], dtype=torch.float64)
    want = torch.tensor([[-1.9494851500028276, -1.2254937510719361], [0.5, -0.32351874678419146],
                         [2.9494851500028276, 0.5784562575035532]], dtype=torch.float64)
    assert torch.allclose(gn(x), want, rtol=0, atol=1e-13)


def test_segment_softmax_conventions():
    """PyG utils.softmax: max over the segment on detached values, +1e-16 in the denominator; utils.scatter 'max' leaves
    empty segments at 0 (include_self=False on a zero tensor)."""
    src = torch.tensor([1.0, 3.0, -2.0], dtype=torch.float64, requires_grad=True)
    idx = torch.tensor([0, 0, 2])
    p = pyg.segment_softmax(src, idx, 4)
    e = math.exp(-2.0)
    assert torch.allclose(p, torch.tensor([e / (1 + e + 1e-16), 1 / (1 + e + 1e-16), 1 / (1 + 1e-16)], dtype=torch.float64), atol=1e-15)
    assert pyg.scatter(src.detach(), idx, 4, "max").tolist() == [3.0, 0.0, -2.0, 0.0]
    assert pyg.scatter(src.detach(), idx, 4, "mean").tolist() == [2.0, 0.0, -2.0, 0.0]
    (g,) = torch.autograd.grad(p[0], src)   # gradient flows through the exp and the sum, not through the max
    pd = p.detach()
    assert torch.allclose(g, torch.stack([pd[0] * (1 - pd[0]), -pd[0] * pd[1], torch.zeros((), dtype=torch.float64)]), atol=1e-15)


# ------------------------------------------------------------------------------------------------------------
# (iii) derivatives, first and second order
# ------------------------------------------------------------------------------------------------------------
def _conv_fn(conv, ei):
    names = [k for k, _ in conv.named_parameters()]

    def fn(x, *params):
        return torch.func.functional_call(conv, dict(zip(names, params)), (x, ei))

    return fn, [p.detach().clone().requires_grad_(True) for _, p in conv.named_parameters()]


@pytest.mark.parametrize("kind", ["GATConv", "GATv2Conv", "GCNConv", "GraphConv"])
def test_conv_gradcheck_and_gradgradcheck(kind):
    torch.manual_seed(11)
    n, cin, c = 6, 3, 2
    ei = _graph(n, 14, seed=21, self_loops=0 if kind == "GraphConv" else 1)
    conv = _dbl(getattr(pyg, kind)(cin, c))
    fn, params = _conv_fn(conv, ei)
    x = torch.randn(n, cin, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(fn, (x, *params), eps=1e-6, atol=1e-5)
    # second order w.r.t. the input only is what WGAN-GP needs; check the full thing anyway (LeakyReLU kinks are measure zero)
    assert torch.autograd.gradgradcheck(fn, (x, *params), eps=1e-6, atol=1e-4)


def test_graphnorm_gradcheck_and_gradgradcheck():
    torch.manual_seed(12)
    gn = _dbl(pyg.GraphNorm(3))
    with torch.no_grad():
        gn.weight.normal_(), gn.bias.normal_(), gn.mean_scale.normal_()
    names = [k for k, _ in gn.named_parameters()]
    params = [p.detach().clone().requires_grad_(True) for _, p in gn.named_parameters()]
    x = torch.randn(7, 3, dtype=torch.float64, requires_grad=True)
    fn = lambda x, *ps: torch.func.functional_call(gn, dict(zip(names, ps)), (x,))
    assert torch.autograd.gradcheck(fn, (x, *params), eps=1e-6, atol=1e-5)
    assert torch.autograd.gradgradcheck(fn, (x, *params), eps=1e-6, atol=1e-4)


# ------------------------------------------------------------------------------------------------------------
# (iv) Sequential / Batch conventions
# ------------------------------------------------------------------------------------------------------------
class _Spy(nn.Module):
    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, *args, **kwargs):
        self.calls.append((len(args), sorted(kwargs)))
        return args[0] + 1


def test_sequential_call_convention_and_child_names():
    """models.py:68-90: ``tgnn.Sequential("x, edge_index", [(conv, "x, edge_index -> x"), GraphNorm, ReLU, Dropout, ...])``
    called as ``self.encoder(x=x, edge_index=...)`` (models.py:144,242): signed entries get the named variables, bare
    modules ONLY the previous output - so GraphNorm never sees ``batch``."""
    conv, bare1, bare2, conv2 = _Spy(), _Spy(), _Spy(), _Spy()
    seq = pyg.Sequential("x, edge_index", [(conv, "x, edge_index -> x"), bare1, bare2, (conv2, "x, edge_index -> x")])
    assert [k for k, _ in seq.named_children()] == ["module_0", "module_1", "module_2", "module_3"]
    x, ei = torch.zeros(3, 2), torch.zeros(2, 4, dtype=torch.long)
    out = seq(x=x, edge_index=ei)
    assert torch.equal(out, x + 4)
    assert conv.calls == [(2, [])] and conv2.calls == [(2, [])] and bare1.calls == [(1, [])] and bare2.calls == [(1, [])]
    # the reference's block: GraphNorm inside Sequential == GraphNorm over all rows as one graph
    torch.manual_seed(0)
    gn = pyg.GraphNorm(2)
    blk = pyg.Sequential("x, edge_index", [(pyg.GATConv(2, 2), "x, edge_index -> x"), gn, nn.ReLU(True)])
    xx, ee = torch.randn(5, 2), torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]])
    o = blk.module_0(xx, ee)
    assert torch.allclose(blk(x=xx, edge_index=ee), torch.relu(gn(o, None)))
    assert [k for k, _ in blk.module_0.named_parameters()] == ["att_src", "att_dst", "bias", "lin.weight"]
    assert [k for k, _ in gn.named_parameters()] == ["weight", "bias", "mean_scale"]


def test_batch_from_data_list_round_trip():
    """data.py:160-161 / trainer.py:364,421: tensors concatenate on dim 0 except ``*index*`` keys (dim -1, incremented by
    the running node count); ``batch``/``ptr`` are added; list attributes become a list per graph; ``batch[i]`` undoes it."""
    gs = []
    for k, n in enumerate((3, 5, 2)):
        gs.append(pyg.Data(x=torch.full((n, 2), float(k)), edge_index=torch.tensor([[0, n - 1], [n - 1, 0]]),
                           type=torch.arange(n), data_number=[f"{k:06d}"] * n, site_area=torch.full((n,), 100 + k)))
    b = pyg.Batch.from_data_list(gs)
    assert b.num_graphs == 3 and b.num_nodes == 10
    assert b.ptr.tolist() == [0, 3, 8, 10] and b.batch.tolist() == [0] * 3 + [1] * 5 + [2] * 2
    assert b.edge_index.tolist() == [[0, 2, 3, 7, 8, 9], [2, 0, 7, 3, 9, 8]]
    assert b.data_number == [g.data_number for g in gs]
    for i, g in enumerate(gs):
        back = b[i]
        assert back.num_nodes == g.num_nodes and back.data_number == g.data_number
        for key in ("x", "edge_index", "type", "site_area"):
            assert torch.equal(getattr(back, key), getattr(g, key)), key


# ------------------------------------------------------------------------------------------------------------
# (v) initialisers and RNG consumption (PyG 2.6.1: every conv re-draws its Linears in reset_parameters())
# ------------------------------------------------------------------------------------------------------------
def _draws(shapes_and_bounds, seed):
    torch.manual_seed(seed)
    return [torch.empty(*shape).uniform_(-b, b) for shape, b in shapes_and_bounds]


def test_conv_initialisers_and_draw_order():
    cin, c, seed = 12, 8, 99
    glo = math.sqrt(6.0 / (cin + c))
    fan = 1.0 / math.sqrt(cin)
    att = math.sqrt(6.0 / (1 + c))
    # GATConv: lin (Linear.__init__), lin again (reset_parameters), att_src, att_dst; bias zeros
    torch.manual_seed(seed)
    m = pyg.GATConv(cin, c)
    d = _draws([((c, cin), glo), ((c, cin), glo), ((1, 1, c), att), ((1, 1, c), att)], seed)
    assert torch.equal(m.lin.weight, d[1]) and torch.equal(m.att_src, d[2]) and torch.equal(m.att_dst, d[3])
    assert not m.bias.any()
    # GCNConv: lin twice
    torch.manual_seed(seed)
    m = pyg.GCNConv(cin, c)
    d = _draws([((c, cin), glo), ((c, cin), glo)], seed)
    assert torch.equal(m.lin.weight, d[1]) and not m.bias.any()
    # GraphConv: lin_rel (weight, bias), lin_root, then all three again; kaiming-uniform(a=sqrt 5) == U(+-1/sqrt(in))
    torch.manual_seed(seed)
    m = pyg.GraphConv(cin, c)
    d = _draws([((c, cin), fan), ((c,), fan), ((c, cin), fan)] * 2, seed)
    assert torch.equal(m.lin_rel.weight, d[3]) and torch.equal(m.lin_rel.bias, d[4]) and torch.equal(m.lin_root.weight, d[5])
    assert m.lin_rel.bias.abs().max() <= fan and m.lin_rel.bias.abs().max() > 0
    # GATv2Conv: lin_l (w,b), lin_r (w,b) at construction, lin_l, lin_r again, then att
    torch.manual_seed(seed)
    m = pyg.GATv2Conv(cin, c)
    d = _draws([((c, cin), glo), ((c,), fan)] * 4 + [((1, 1, c), att)], seed)
    assert torch.equal(m.lin_l.weight, d[4]) and torch.equal(m.lin_l.bias, d[5])
    assert torch.equal(m.lin_r.weight, d[6]) and torch.equal(m.lin_r.bias, d[7]) and torch.equal(m.att, d[8])
    assert [k for k, _ in m.named_parameters()] == ["att", "bias", "lin_l.weight", "lin_l.bias", "lin_r.weight", "lin_r.bias"]


@pytest.mark.parametrize("kind", ["GATConv", "GCNConv", "GraphConv", "GATv2Conv"])
def test_product_parameter_holders_initialise_like_the_oracle(kind):
    """The drop-in modules' parameter holders consume the RNG exactly like the oracle's (= PyG's) constructors."""
    from building_gan_b200 import models as M
    torch.manual_seed(7)
    a = getattr(pyg, kind)(16, 8)
    ra = torch.rand(1)
    torch.manual_seed(7)
    b = getattr(M, kind)(16, 8)
    rb = torch.rand(1)
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    assert list(pa) == list(pb)
    for k in pa:
        assert torch.equal(pa[k], pb[k]), k
    assert torch.equal(ra, rb)
