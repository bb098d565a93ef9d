"""K1/K2 parity: CUDA SpMM (through the C-ABI) vs the CPU oracle."""
import pytest
import torch

from oracle import gist_oracle as O
from tests.util import assert_close, ograph, powerlaw_graph, random_graph

pytestmark = pytest.mark.gpu

DIMS = [1, 3, 7, 16, 32, 41, 64, 100, 128, 256, 602]


def _gist(src, dst, n):
    from gist_b200 import GistGraph
    return GistGraph.from_edges(src, dst, n, device='cuda')


@pytest.mark.parametrize('d', DIMS)
def test_copy_src_sum_matches_oracle(d):
    from gist_b200 import ops
    n, nnz = 1000, 20000
    src, dst = random_graph(n, nnz, seed=d, isolated=17)
    g = _gist(src, dst, n)
    og = ograph(src, dst, n)
    torch.manual_seed(d)
    x = torch.randn(n, d)
    ref = O.copy_src_sum(og, x)
    ref64 = O.copy_src_sum(og, x.double())
    got = ops.copy_src_sum(g, x.cuda())
    assert_close(got, ref, what='fp32 oracle d=%d' % d)
    assert_close(got, ref64, what='fp64 oracle d=%d' % d)
    # zero in-degree rows are exactly zero
    assert (got[n - 17:] == 0).all()


@pytest.mark.parametrize('d', [16, 100, 256, 602])
def test_powerlaw_rows(d):
    from gist_b200 import ops
    n = 3000
    src, dst = powerlaw_graph(n, 40, seed=5)
    g = _gist(src, dst, n)
    og = ograph(src, dst, n)
    assert int(og.in_degrees().max()) > 2000      # a few very heavy rows
    x = torch.randn(n, d)
    assert_close(ops.copy_src_sum(g, x.cuda()), O.copy_src_sum(og, x.double()), what='powerlaw')


@pytest.mark.parametrize('d', [7, 64, 130, 602])
@pytest.mark.parametrize('flags', [0, 2, 4])
def test_fused_epilogue(d, flags):
    """act(t * A(s*X) + addend + bias) and the self-copy output, all vector widths."""
    from gist_b200 import ops
    n, nnz = 700, 9000
    src, dst = random_graph(n, nnz, seed=3)
    g = _gist(src, dst, n)
    og = ograph(src, dst, n)
    torch.manual_seed(0)
    x, s, t = torch.randn(n, d), torch.rand(n) + 0.5, torch.rand(n) + 0.5
    bias, addend = torch.randn(d), torch.randn(n, d)
    ref = torch.relu(O.copy_src_sum(og, x.double() * s.double()[:, None]) * t.double()[:, None]
                     + addend.double() + bias.double())
    out = torch.full((n, 2 * d + 2), float('nan'), device='cuda')
    xs = x.cuda()
    y = out[:, d:2 * d]
    self_out = out[:, :d]
    ops.spmm_raw(g.rowptr, g.col_buffer, n, n, xs, y, src_scale=s.cuda(), dst_scale=t.cuda(),
                 bias=bias.cuda(), addend=addend.cuda(), self_out=self_out, relu=True, flags=flags)
    assert_close(y, ref, what='epilogue')
    assert torch.equal(self_out.cpu(), x)
    assert torch.isnan(out[:, 2 * d:]).all()      # nothing written outside the views


def test_empty_and_tiny():
    from gist_b200 import GistGraph, ops
    # no edges at all
    g = GistGraph.from_edges(torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long), 5,
                             device='cuda')
    y = ops.copy_src_sum(g, torch.ones(5, 8, device='cuda'))
    assert (y == 0).all()
    # hand-worked 4-node case: edges 0->1, 2->1, 2->1 (multi), 3->3 (self loop)
    src = torch.tensor([0, 2, 2, 3])
    dst = torch.tensor([1, 1, 1, 3])
    g = GistGraph.from_edges(src, dst, 4, device='cuda')
    x = torch.tensor([[1., 10.], [2., 20.], [3., 30.], [4., 40.]], device='cuda')
    y = ops.copy_src_sum(g, x).cpu()
    assert torch.equal(y, torch.tensor([[0., 0.], [7., 70.], [0., 0.], [4., 40.]]))
    assert torch.equal(g.in_degrees().cpu(), torch.tensor([0, 3, 0, 1]))
    assert torch.equal(g.out_degrees().cpu(), torch.tensor([1, 0, 2, 1]))
    assert torch.equal(g.inv_in_degree().cpu(), torch.tensor([0., 1 / 3., 0., 1.]))


@pytest.mark.parametrize('d', [5, 32, 602])
def test_backward_matches_autograd_of_oracle(d):
    """K2 (CSC transpose) gradients vs autograd through the oracle, non-symmetric graph."""
    from gist_b200 import ops
    n, nnz = 600, 7000
    src, dst = random_graph(n, nnz, seed=11, isolated=5)
    g = _gist(src, dst, n)
    assert not g.is_symmetric()
    og = ograph(src, dst, n)
    torch.manual_seed(1)
    x = torch.randn(n, d, dtype=torch.double, requires_grad=True)
    w = torch.randn(n, 2 * d, dtype=torch.double)
    zr = torch.cat((x, O.copy_src_sum(og, x) * O.sage_norm(og, torch.double)), 1)
    (zr * w).sum().backward()
    xg = x.detach().float().cuda().requires_grad_(True)
    z = ops.sage_concat(g, xg)
    (z * w.float().cuda()).sum().backward()
    assert_close(z, zr, what='sage_concat fwd')
    assert_close(xg.grad, x.grad, what='sage_concat bwd')

    # scaled SpMM with bias + relu
    s, t = torch.rand(n, dtype=torch.double) + .5, torch.rand(n, dtype=torch.double) + .5
    b = torch.randn(d, dtype=torch.double, requires_grad=True)
    x2 = torch.randn(n, d, dtype=torch.double, requires_grad=True)
    yr = torch.relu(O.copy_src_sum(og, x2 * s[:, None]) * t[:, None] + b)
    wy = torch.randn(n, d, dtype=torch.double)
    (yr * wy).sum().backward()
    x2g = x2.detach().float().cuda().requires_grad_(True)
    bg = b.detach().float().cuda().requires_grad_(True)
    y = ops.gspmm(g, x2g, s.float().cuda(), t.float().cuda(), bg, True)
    (y * wy.float().cuda()).sum().backward()
    assert_close(y, yr, what='gspmm fwd')
    assert_close(x2g.grad, x2.grad, what='gspmm dX')
    assert_close(bg.grad, b.grad, rtol=1e-4, what='gspmm dbias')


def test_update_all_duck_type():
    import gist_b200.function as fn
    n = 300
    src, dst = random_graph(n, 3000, seed=2)
    g = _gist(src, dst, n).local_var()
    x = torch.randn(n, 12)
    g.ndata['h'] = x.cuda()
    g.update_all(fn.copy_src(src='h', out='m'), fn.sum(msg='m', out='h'))
    assert_close(g.ndata.pop('h'), O.copy_src_sum(ograph(src, dst, n), x.double()))


def test_no_cpu_path():
    from gist_b200 import GistGraph, ops
    from gist_b200._lib import GistLibraryError
    g = GistGraph.from_edges(torch.tensor([0]), torch.tensor([1]), 2)
    with pytest.raises(GistLibraryError):
        ops.copy_src_sum(g, torch.ones(2, 4))


def test_full_size_linearity_and_degree_identity():
    """Size-independent properties at a large size the oracle cannot time: A·1 equals
    the in-degree vector exactly, and A(ax+by) = aAx + bAy."""
    from gist_b200 import ops, synth
    ds = synth.make('reddit', seed=0, device='cuda', scale=0.1)
    from gist_b200 import GistGraph
    g = GistGraph.from_edges(ds.src, ds.dst, ds.num_nodes)
    n = ds.num_nodes
    ones = torch.ones(n, 64, device='cuda')
    y = ops.copy_src_sum(g, ones)
    assert torch.equal(y[:, 0], g.in_degrees().float())       # integers < 2^24: exact
    assert (y == y[:, :1]).all()
    x1, x2 = torch.randn(n, 602, device='cuda'), torch.randn(n, 602, device='cuda')
    lhs = ops.copy_src_sum(g, 2 * x1 - 3 * x2)
    rhs = 2 * ops.copy_src_sum(g, x1) - 3 * ops.copy_src_sum(g, x2)
    assert_close(lhs, rhs, rtol=1e-4, what='linearity')


# ------------------------------------------------------------ segment-balanced K1/K2 ----
@pytest.mark.parametrize('d', [1, 7, 32, 100, 256, 602])
@pytest.mark.parametrize('has_scale', [False, True])
@pytest.mark.parametrize('slab', [True, False])
def test_segment_balanced_matches_oracle_and_is_deterministic(d, has_scale, slab, monkeypatch):
    """The scheduled kernels (cluster batches) on a power-law graph: hub rows span dozens of segments
    finished by whichever group arrives last, yet the result is run-to-run identical and within
    tolerance of the fp64 oracle; arrival counters are left zero.  slab=True lets the launch stage
    column slabs of X in shared memory where the layout allows it (here: X rows 16-byte aligned,
    outputs at odd offsets -> scalar stores), slab=False forces the L2-gather segment kernel."""
    from gist_b200 import _lib, ops
    fl = 0 if slab else _lib.SPMM_SLAB_OFF
    monkeypatch.setattr(ops, 'SLAB_ENABLED', True)          # the slab kernel is opt-in (ops.SLAB_ENABLED)
    n = 3000
    src, dst = powerlaw_graph(n, 40, seed=7)
    g = _gist(src, dst, n)
    og = ograph(src, dst, n)
    sch = g.seg_schedule()
    assert sch is not None and sch.seg_len == 64
    deg = og.in_degrees()
    nseg = torch.clamp((deg + 63) // 64, min=1)
    assert int(sch.seg_ptr[n].item()) == int(nseg.sum()) and int(nseg.max()) > 20
    assert torch.equal(sch.seg_ptr.cpu().long()[1:] - sch.seg_ptr.cpu().long()[:-1], nseg)
    torch.manual_seed(d)
    x = torch.randn(n, d)
    s = (torch.rand(n) + 0.5) if has_scale else None
    t = (torch.rand(n) + 0.5) if has_scale else None
    addend, bias = torch.randn(n, d), torch.randn(d)
    xin = x.double() * s.double()[:, None] if has_scale else x.double()
    ref = O.copy_src_sum(og, xin)
    if has_scale:
        ref = ref * t.double()[:, None]
    ref = torch.relu(ref + addend.double() + bias.double())
    cu = lambda v: v.cuda() if v is not None else None      # noqa: E731
    outs = []
    for _ in range(3):
        out = torch.full((n, 2 * d + 1), float('nan'), device='cuda')
        ops.spmm_raw(g.rowptr, g.col_buffer, n, n, x.cuda(), out[:, d:2 * d], src_scale=cu(s), dst_scale=cu(t),
                     bias=bias.cuda(), addend=addend.cuda(), self_out=out[:, :d], relu=True, schedule=sch, flags=fl)
        outs.append(out)
    assert_close(outs[0][:, d:2 * d], ref, what='segment-balanced')
    assert torch.equal(outs[0][:, :d].cpu(), x)
    assert torch.isnan(outs[0][:, 2 * d:]).all()
    assert torch.equal(outs[0][:, :2 * d], outs[1][:, :2 * d]) and torch.equal(outs[1][:, :2 * d], outs[2][:, :2 * d])
    assert (sch.counters(1) == 0).all()
    # same numbers as the row-per-group kernel up to summation order
    plain = torch.empty(n, d, device='cuda')
    ops.spmm_raw(g.rowptr, g.col_buffer, n, n, x.cuda(), plain, src_scale=cu(s), dst_scale=cu(t),
                 bias=bias.cuda(), addend=addend.cuda(), relu=True)
    assert_close(outs[0][:, d:2 * d], plain, what='vs row-per-group kernel')


@pytest.mark.parametrize('d', [16, 30, 32, 41, 256, 602])
@pytest.mark.parametrize('mode', ['fwd', 'transpose', 'background'])
def test_slab_kernel_on_the_training_layouts(d, mode, monkeypatch):
    """spmm_slab_kernel on the buffer layouts of the training step: X with 16-byte-aligned (padded)
    rows, z = [h | agg] in ONE padded buffer (the aggregate half is 16-, 8- or 4-byte aligned depending
    on d), dst_scale = 1/deg, self copy; 'transpose' = the backward launch (src_scale, addend, no self
    copy); 'background' = one CTA per slab.  Against the fp64 oracle, bit-identical across repeats and
    across the foreground / background grids (the summation order does not depend on the grid)."""
    from gist_b200 import ops
    monkeypatch.setattr(ops, 'SLAB_ENABLED', True)
    n = 2600
    src, dst = powerlaw_graph(n, 40, seed=11)
    src, dst = torch.cat([src, dst]), torch.cat([dst, src])           # symmetric, like the bench graphs
    g = _gist(src, dst, n)
    og = ograph(src, dst, n)
    sch = g.seg_schedule()
    torch.manual_seed(d)
    x = ops.pad_rows(torch.randn(n, d, device='cuda'))
    assert x.data_ptr() % 16 == 0 and x.stride(0) % 4 == 0
    inv = g.inv_in_degree()
    norm = O.sage_norm(og, torch.float64)
    if mode == 'transpose':
        add = ops._padded_empty(n, d, 'cuda').normal_()
        ref = O.copy_src_sum(og, x.double().cpu() * norm) + add.double().cpu()      # symmetric: A^T = A
        outs = []
        for _ in range(2):
            y = torch.empty(n, d, device='cuda')
            ops.spmm_raw(g.rowptr, g.col_buffer, n, n, x, y, src_scale=inv, addend=add, schedule=sch)
            outs.append(y)
        assert_close(outs[0], ref, what='slab transpose')
        assert torch.equal(outs[0], outs[1])
        return
    ref = O.copy_src_sum(og, x.double().cpu()) * norm
    outs = []
    for bg in ((0, 0, 1) if mode == 'background' else (0, 0)):
        z = ops._padded_empty(n, 2 * d, 'cuda')
        z.fill_(float('nan'))
        ops.spmm_raw(g.rowptr, g.col_buffer, n, n, x, z[:, d:], dst_scale=inv, self_out=z[:, :d], schedule=sch,
                     flags=bg << 8)
        outs.append(z)
    assert_close(outs[0][:, d:], ref, what='slab forward')
    assert torch.equal(outs[0][:, :d], x)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert (sch.counters(1) == 0).all()


def test_segment_schedule_edge_cases():
    from gist_b200 import GistGraph, ops
    # edgeless graph: one (empty) segment per row, output = epilogue of zero
    g = GistGraph.from_edges(torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long), 9, device='cuda')
    sch = g.seg_schedule()
    assert sch.seg_ptr.tolist() == list(range(10))
    y = torch.empty(9, 5, device='cuda')
    ops.spmm_raw(g.rowptr, g.col_buffer, 9, 9, torch.ones(9, 5, device='cuda'), y, bias=torch.ones(5, device='cuda'),
                 schedule=sch)
    assert (y == 1).all()
    # a row of exactly 64, 65 and 128 edges: 1, 2 and 2 segments
    src = torch.cat([torch.arange(64), torch.arange(65), torch.arange(128)])
    dst = torch.cat([torch.zeros(64), torch.ones(65), torch.full((128,), 2.0)]).long()
    g = GistGraph.from_edges(src, dst, 130, device='cuda')
    sch = g.seg_schedule()
    assert sch.seg_ptr[:4].tolist() == [0, 1, 3, 5]
    x = torch.arange(130, dtype=torch.float32, device='cuda').unsqueeze(1).repeat(1, 3)
    y = ops.copy_src_sum(g, x)
    assert y[0, 0].item() == sum(range(64)) and y[1, 0].item() == sum(range(65)) and y[2, 0].item() == sum(range(128))
    # large graphs keep the row-per-warp kernel
    from gist_b200.graph import GistGraph as GG
    big = GG.from_edges(torch.arange(40000), (torch.arange(40000) + 1) % 40000, 40000, device='cuda')
    assert big.seg_schedule() is None


def test_segment_schedule_rebuilt_in_place_for_reused_batch_buffers():
    """subgraph(out=...) (the pipelined trainer's buffer reuse) refreshes the schedule."""
    import numpy as np
    from gist_b200 import ops
    n = 4000
    src, dst = powerlaw_graph(n, 30, seed=2)
    src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    g = _gist(src, dst, n)
    g.ndata['feat'] = torch.randn(n, 16, device='cuda')
    rng = np.random.RandomState(0)
    ids1 = torch.from_numpy(rng.choice(n, 600, replace=False)).cuda()
    ids2 = torch.from_numpy(rng.choice(n, 600, replace=False)).cuda()
    cap = int((g.rowptr[1:] - g.rowptr[:-1]).max().item()) * 600
    sg = g.subgraph(ids1, col_capacity=cap)
    sch = sg.seg_schedule()
    y1 = ops.copy_src_sum(sg, sg.ndata['feat'])
    fresh2 = g.subgraph(ids2, col_capacity=cap)
    want = ops.copy_src_sum(fresh2, fresh2.ndata['feat'])
    sg2 = g.subgraph(ids2, col_capacity=cap, out=sg)
    assert sg2 is sg and sg.seg_schedule() is sch
    got = ops.copy_src_sum(sg, sg.ndata['feat'])
    assert torch.equal(got, want) and not torch.equal(got, y1)
