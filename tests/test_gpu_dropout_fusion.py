"""Fused dropout / 3xTF32-split path of the SAGE layer (K1 extended epilogue, K4 mask epilogue,
layer-norm / cross-entropy backward low halves) against the same computation composed from the
stand-alone kernels and an fp64 torch reference with the SAME (materialised) mask.

Tolerances: bit-exact where the fused kernel performs the same float operations as the composed
path (mask application, low halves); 1e-5 norm-wise relative for 3xTF32 layer outputs / gradients
against fp64 (the north_star's fp32 parity tolerance); 2e-3 for single-pass TF32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import random_graph

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return ((got.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)).item()


def _graph(n, nnz, seed):
    from gist_b200 import GistGraph
    src, dst = random_graph(n, nnz, seed=seed)
    return GistGraph.from_edges(src, dst, n, device='cuda')


def _mask(n, width, p, stream_id, step=None):
    """The multiplier matrix (0 or 1/(1-p)) the fused kernels apply, materialised."""
    from gist_b200 import ops
    st = ops.dropout_state('cuda')
    desc = st.desc(p, stream_id, step=step)
    return ops.dropout(torch.ones(n, width, device='cuda'), desc)


def test_dropout_mask_statistics_clock_and_seed():
    from gist_b200 import ops
    torch.manual_seed(11)
    st = ops.dropout_state('cuda')
    n, w, p = 3000, 512, 0.2
    m1 = _mask(n, w, p, 5)
    assert set(torch.unique(m1).tolist()) == {0.0, np.float32(1.0 / 0.8).item()}
    keep = (m1 != 0).float().mean().item()
    assert abs(keep - 0.8) < 4 * (0.16 / (n * w)) ** 0.5 + 1e-4          # binomial 4-sigma
    assert abs((m1 != 0).float().mean(0) - 0.8).max().item() < 0.06       # no dead columns / rows
    assert abs((m1 != 0).float().mean(1) - 0.8).max().item() < 0.12
    assert torch.equal(m1, _mask(n, w, p, 5))                              # pure function of its inputs
    assert not torch.equal(m1, _mask(n, w, p, 6))                          # other layer, other mask
    st.tick()
    m2 = _mask(n, w, p, 5)
    assert not torch.equal(m1, m2)                                          # the clock moved
    assert abs(((m1 != 0) & (m2 != 0)).float().mean().item() - 0.64) < 0.01   # independent draws
    torch.manual_seed(12)
    assert not torch.equal(m2, _mask(n, w, p, 5))                          # follows torch.manual_seed
    torch.manual_seed(11)
    assert torch.equal(m2, _mask(n, w, p, 5))
    # column offsets address the same logical matrix
    desc = st.desc(p, 5)
    part = ops.dropout(torch.ones(n, 100, device='cuda'), desc, col0=37)
    assert torch.equal(part, m2[:, 37:137])


@pytest.mark.parametrize('d', [602, 256, 32, 7])
def test_sage_prepare_matches_composed(d):
    """K1's extended epilogue == sage_concat -> mask -> split, bit for bit."""
    from gist_b200 import ops
    n = 700
    g = _graph(n, 9000, seed=d)
    torch.manual_seed(d)
    h = torch.randn(n, d, device='cuda')
    ops.set_matmul_precision('3xtf32')
    try:
        pre = ops.sage_prepare(g, h, 0.3, 9)
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    z_plain = ops.sage_concat(g, h)
    m = _mask(n, 2 * d, 0.3, 9, step=pre.step_saved)
    assert torch.equal(pre.z, z_plain * m)
    assert torch.equal(pre.z_lo, ops.split_tf32(pre.z.contiguous()))
    assert int(pre.step_saved.item()) == int(ops.dropout_state('cuda').step.item())
    # no dropout: plain concat, still split
    ops.set_matmul_precision('3xtf32')
    try:
        pre0 = ops.sage_prepare(g, h, 0.0, 9)
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    assert torch.equal(pre0.z, z_plain) and not pre0.dropped and pre0.step_saved is None


@pytest.mark.parametrize('background', [1, 2, 5])
@pytest.mark.parametrize('n,nnz,d', [(700, 9000, 602), (5000, 200000, 256), (40, 300, 7)])
def test_sage_prepare_background_mode_is_bit_identical(n, nnz, d, background):
    """Background mode (GIST_SPMM_BG_SHIFT: a capped grid walking the work with a grid stride, used
    for the next batch's layer-0 aggregation beside the training branch) computes exactly what the
    one-CTA-per-item launch computes — including hub rows split over the CTA."""
    from gist_b200 import GistGraph, ops
    src, dst = random_graph(n, nnz, seed=n + d)
    torch.manual_seed(d)
    hub_src = torch.randint(0, n, (min(4 * n, 3000),), dtype=src.dtype)      # rows 0 and n-1: hub rows
    src = torch.cat((src, hub_src, hub_src))
    dst = torch.cat((dst, torch.zeros_like(hub_src), torch.full_like(hub_src, n - 1)))
    g = GistGraph.from_edges(src, dst, n, device='cuda')
    h = torch.randn(n, d, device='cuda')
    ops.set_matmul_precision('3xtf32')
    try:
        ref = ops.sage_prepare(g, h, 0.3, 4, balanced=False)
        got = ops.sage_prepare(g, h, 0.3, 4, balanced=False, background=background)
        again = ops.sage_prepare(g, h, 0.3, 4, balanced=False, background=background, out=got)
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    assert again is got
    assert torch.equal(got.z, ref.z) and torch.equal(got.z_lo, ref.z_lo)
    assert torch.equal(got.step_saved, ref.step_saved)


@pytest.mark.parametrize('x3', [False, True])
@pytest.mark.parametrize('shape', [(700, 512, 256), (300, 70, 100), (1000, 1204, 32)])
def test_gemm_dropmask_matches_masked_gemm(shape, x3):
    from gist_b200 import ops
    M, N, K = shape
    torch.manual_seed(M + N)
    dy = torch.randn(M, (K + 3) // 4 * 4, device='cuda')[:, :K]
    W = torch.randn(K, (N + 3) // 4 * 4, device='cuda')[:, :N]          # [out, in]: dz = dy W, B MN-major
    lo = dict(A_lo=ops.split_tf32(dy), B_lo=ops.split_tf32(W)) if x3 else {}
    st = ops.dropout_state('cuda')
    desc = st.desc(0.25, 3)
    got = ops.gemm_dropmask(dy, W, desc, b_mn=True, **lo)
    ref = ops.gemm(dy, W, b_mn=True, flags=2, **lo) * _mask(M, N, 0.25, 3)
    assert torch.equal(got, ref)


def _hub_graph(n, seed):
    """Power-law multigraph whose row 0 is the heaviest (dozens of 64-edge segments): the first
    CTA of the segment-balanced kernel is then NOT the one that finishes row 0."""
    from gist_b200 import GistGraph
    from tests.util import powerlaw_graph
    src, dst = powerlaw_graph(n, 30, seed=seed)
    hub = int(torch.bincount(dst, minlength=n).argmax())
    swap = lambda t: torch.where(t == hub, torch.zeros_like(t), torch.where(t == 0, torch.full_like(t, hub), t))  # noqa: E731
    g = GistGraph.from_edges(swap(src), swap(dst), n, device='cuda')
    assert int(g.rowptr[1]) > 1000
    return g


@pytest.mark.parametrize('precision,tol', [('3xtf32', 1e-5), ('tf32', 2e-3)])
@pytest.mark.parametrize('p', [0.0, 0.4])
@pytest.mark.parametrize('hub', [False, True])
def test_sage_linear_forward_backward(precision, tol, p, hub):
    from gist_b200 import ops
    n, d, out = 600, 96, 40
    g = _hub_graph(n, seed=1) if hub else _graph(n, 8000, seed=1)
    ops.dropout_state('cuda').tick()            # a non-zero clock: the saved step must really be recorded
    torch.manual_seed(3)
    h = torch.randn(n, d, device='cuda', requires_grad=True)
    W = (torch.randn(out, 2 * d, device='cuda') * 0.05).requires_grad_(True)
    b = torch.randn(out, device='cuda', requires_grad=True)
    wy = torch.randn(n, out, device='cuda')
    ops.set_matmul_precision(precision)
    try:
        y = ops.sage_linear(g, h, W, b, p, 4)
        (y * wy).sum().backward()
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    m = _mask(n, 2 * d, p, 4).double() if p else 1.0
    h2, W2, b2 = (t.detach().double().requires_grad_(True) for t in (h, W, b))
    rp, col = g.rowptr.long(), g.col.long()
    rows = torch.repeat_interleave(torch.arange(n, device='cuda'), rp[1:] - rp[:-1])
    agg = torch.zeros(n, d, device='cuda', dtype=torch.float64).index_add_(0, rows, h2[col])
    deg = (rp[1:] - rp[:-1]).double()
    inv = torch.where(deg > 0, 1.0 / deg, torch.zeros_like(deg)).unsqueeze(1)
    z = torch.cat([h2, agg * inv], 1) * m
    y2 = F.linear(z, W2, b2)
    (y2 * wy.double()).sum().backward()
    for got, ref, name in ((y, y2, 'y'), (h.grad, h2.grad, 'dh'), (W.grad, W2.grad, 'dW'), (b.grad, b2.grad, 'db')):
        assert _rel(got, ref) < tol, (name, _rel(got, ref))


def test_ist_sage_layer_train_mode_3xtf32_vs_oracle():
    """Train-mode ISTSAGELayer on the fused path against the CPU oracle with the same mask."""
    from gist_b200 import ISTSAGELayer, ops
    from oracle import gist_oracle as O
    from tests.util import ograph
    n, fin, fout = 500, 64, 32
    src, dst = random_graph(n, 6000, seed=4)
    from gist_b200 import GistGraph
    g, og = GistGraph.from_edges(src, dst, n, device='cuda'), ograph(src, dst, n)
    torch.manual_seed(0)
    layer = ISTSAGELayer(fin, fout, 0.5, True, activation=F.relu).cuda().train()
    x = torch.randn(n, fin)
    ops.set_matmul_precision('3xtf32')
    try:
        out = layer(g, x.cuda())
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    mask = _mask(n, 2 * fin, 0.5, layer._drop_stream).cpu().double()
    ref = O.ist_sage_layer(og, x.double(), layer.linear.weight.detach().double().cpu(),
                           layer.linear.bias.detach().double().cpu(), True, F.relu, mask)
    assert _rel(out.cpu(), ref) < 1e-5


def test_backward_low_halves_match_split():
    """dx_lo of the layer-norm backward and dlogits_lo of the cross-entropy backward are exactly
    split_tf32 of the gradients they accompany, and reach the consumer through the hand-off."""
    from gist_b200 import ops
    torch.manual_seed(5)
    ops.set_matmul_precision('3xtf32')
    try:
        x = torch.randn(777, 256, device='cuda', requires_grad=True)
        y = ops.layer_norm_act(x, 1e-5, relu=True)
        dy = torch.randn_like(y)
        (dx,) = torch.autograd.grad(y, x, dy)
        lo = ops._lo_take(dx)
        assert lo is not None and torch.equal(lo, ops.split_tf32(dx))
        logits = torch.randn(500, 41, device='cuda', requires_grad=True)
        labels = torch.randint(0, 41, (500,), device='cuda')
        mask = torch.rand(500, device='cuda') < 0.7
        loss = ops.masked_cross_entropy(logits, labels, mask)
        (dl,) = torch.autograd.grad(loss, logits)
        lo = ops._lo_take(dl)
        assert lo is not None and torch.equal(lo, ops.split_tf32(dl))
        assert ops._lo_take(dl) is None                     # consumed
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
