"""K4 parity: tcgen05 TF32 GEMM (through the C-ABI) vs an fp64 matmul.  Tolerance: TF32
keeps 10 mantissa bits of each input (truncation, 2^-10 relative per element); for
N(0,1) data the products' relative error norm-wise is ~1e-3 — asserted at 2e-3 of
||A||·||B|| per entry, far from the 1e-5 of the aggregation kernels, and stated in DESIGN.md."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 64, 32), (1, 16, 8), (2586, 256, 1204), (2586, 41, 512), (256, 1204, 2586),
          (300, 70, 100), (129, 65, 36), (4096, 512, 1024), (2280, 1024, 2048), (20000, 512, 64)]


def _err_bound(A, B):
    # per-entry bound: sum_k |a||b| * 2 * 2^-10
    return (A.abs().double() @ B.abs().double().t()) * 2.0 ** -9


@pytest.mark.parametrize('shape', SHAPES)
def test_gemm_tn_tf32(shape):
    from gist_b200 import ops
    M, N, K = shape
    torch.manual_seed(M + N + K)
    Kp = (K + 3) // 4 * 4
    A = torch.randn(M, Kp, device='cuda')[:, :K]
    B = torch.randn(N, Kp, device='cuda')[:, :K]
    bias = torch.randn(N, device='cuda')
    ref = A.double() @ B.double().t()
    got = ops.gemm_tn(A, B)
    bound = _err_bound(A, B)
    assert ((got.double() - ref).abs() <= bound + 1e-6).all(), (got.double() - ref).abs().max().item()
    # typical error is far below the worst-case bound
    rel = (got.double() - ref).norm() / ref.norm()
    assert rel < 2e-3, rel
    got2 = ops.gemm_tn(A, B, bias=bias, relu=True)
    ref2 = torch.relu(ref + bias.double())
    assert ((got2.double() - ref2).abs() <= bound + 1e-6).all()
    # exactness on tf32-representable inputs (small integers): tensor core result must be exact
    Ai = torch.randint(-4, 5, (M, Kp), device='cuda').float()[:, :K]
    Bi = torch.randint(-4, 5, (N, Kp), device='cuda').float()[:, :K]
    assert torch.equal(ops.gemm_tn(Ai, Bi), (Ai.double() @ Bi.double().t()).float())


LAYOUT_SHAPES = [(128, 64, 32), (2586, 512, 256), (256, 1204, 2586), (256, 512, 2586), (41, 512, 2586),
                 (2586, 512, 41), (300, 70, 100), (129, 65, 36), (1000, 300, 520), (2280, 1024, 1056)]


@pytest.mark.parametrize('shape', LAYOUT_SHAPES)
@pytest.mark.parametrize('a_mn', [False, True])
@pytest.mark.parametrize('b_mn', [False, True])
def test_gemm_layouts(shape, a_mn, b_mn):
    """Every storage-layout combination (K-major / MN-major per operand), auto tile width and
    split-K, against fp64; exact on small-integer (tf32-representable) inputs."""
    from gist_b200 import ops
    M, N, K = shape
    torch.manual_seed(M * 7 + N * 3 + K)

    def stored(rows, cols, mn, gen):
        # logical [rows, cols] operand; stored transposed when mn; padded leading dimension
        r, c = (cols, rows) if mn else (rows, cols)
        buf = gen(r, (c + 3) // 4 * 4)[:, :c]
        return buf, (buf.t() if mn else buf)

    rnd = lambda r, c: torch.randn(r, c, device='cuda')                      # noqa: E731
    rint = lambda r, c: torch.randint(-4, 5, (r, c), device='cuda').float()  # noqa: E731
    A_st, A = stored(M, K, a_mn, rnd)
    B_st, B = stored(N, K, b_mn, rnd)
    bias = torch.randn(N, device='cuda')
    ref = A.double() @ B.double().t()
    got = ops.gemm(A_st, B_st, a_mn=a_mn, b_mn=b_mn)
    bound = _err_bound(A, B)
    assert ((got.double() - ref).abs() <= bound + 1e-6).all(), (got.double() - ref).abs().max().item()
    assert (got.double() - ref).norm() / ref.norm() < 2e-3
    got2 = ops.gemm(A_st, B_st, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True)
    assert ((got2.double() - torch.relu(ref + bias.double())).abs() <= bound + 1e-6).all()
    Ai_st, Ai = stored(M, K, a_mn, rint)
    Bi_st, Bi = stored(N, K, b_mn, rint)
    exact = (Ai.double() @ Bi.double().t()).float()
    for fl in (0, 2, 4, 8, 16):      # auto, no split-K, forced 64 / 128 / 256-wide tiles
        assert torch.equal(ops.gemm(Ai_st, Bi_st, a_mn=a_mn, b_mn=b_mn, flags=fl), exact), fl


def test_gemm_splitk_deterministic():
    from gist_b200 import ops, _lib
    torch.manual_seed(1)
    dy = torch.randn(2586, 256, device='cuda')
    z = torch.randn(2586, 1204, device='cuda')
    assert _lib.load().gist_gemm_tf32_workspace_bytes(256, 1204, 2586, 0) > 0    # this shape splits K
    a = ops.gemm(dy, z, a_mn=True, b_mn=True)
    b = ops.gemm(dy, z, a_mn=True, b_mn=True)
    assert torch.equal(a, b)
    ref = dy.double().t() @ z.double()
    assert (a.double() - ref).norm() / ref.norm() < 2e-3


def test_gemm_output_view_and_rejects_unaligned():
    from gist_b200 import ops
    from gist_b200._lib import GistLibraryError
    A = torch.randint(-3, 4, (200, 64), device='cuda').float()
    B = torch.randint(-3, 4, (48, 64), device='cuda').float()
    buf = torch.full((200, 100), float('nan'), device='cuda')
    ops.gemm_tn(A, B, out=buf[:, 3:51])
    assert torch.equal(buf[:, 3:51], A @ B.t())
    assert torch.isnan(buf[:, :3]).all() and torch.isnan(buf[:, 51:]).all()
    with pytest.raises(GistLibraryError):
        ops.gemm_tn(torch.randn(10, 41, device='cuda'), torch.randn(8, 41, device='cuda'))


@pytest.mark.parametrize('shape', [(5, 7), (2586, 256), (100, 1204), (70000, 40)])
def test_transpose(shape):
    from gist_b200 import ops
    x = torch.randn(*shape, device='cuda')
    t = ops.transpose(x)
    assert t.shape == (shape[1], shape[0]) and t.stride(0) % 4 == 0
    assert torch.equal(t, x.t())


def test_linear_tf32_autograd():
    from gist_b200 import ops
    torch.manual_seed(0)
    _check_linear(1000, 1204, 256)
    _check_linear(2586, 512, 41)       # 41-wide logit gradient: padded copy for TMA
    _check_linear(777, 64, 32)


def _check_linear(n, fin, fout):
    from gist_b200 import ops
    z = torch.randn(n, fin, device='cuda', requires_grad=True)
    W = (torch.randn(fout, fin, device='cuda') * 0.03).requires_grad_(True)
    b = torch.randn(fout, device='cuda', requires_grad=True)
    wy = torch.randn(n, fout, device='cuda')
    ops.set_matmul_precision('tf32')
    try:
        y = ops.linear(z, W, b)
        (y * wy).sum().backward()
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    z2, W2, b2 = (t.detach().double().requires_grad_(True) for t in (z, W, b))
    y2 = torch.nn.functional.linear(z2, W2, b2)
    (y2 * wy.double()).sum().backward()
    for got, ref, name in ((y, y2, 'y'), (z.grad, z2.grad, 'dz'), (W.grad, W2.grad, 'dW'), (b.grad, b2.grad, 'db')):
        rel = ((got.double() - ref).norm() / ref.norm()).item()
        assert rel < 2e-3, (name, rel)


# ------------------------------------------------------------------ 3xTF32 ----
# Tolerance: 1e-5 norm-wise relative (the north_star's fp32 parity tolerance); measured ~5e-7,
# i.e. the level of a plain fp32 sgemm.  A per-entry check against sum_k|a||b| guards the tails.
X3_TOL = 1e-5


def test_split_tf32_is_exact_residual():
    """x == trunc_tf32(x) + (x - trunc_tf32(x)) exactly; lo is the TF32 rounding of the residual and
    hi has its 13 low mantissa bits clear."""
    from gist_b200 import ops, _lib
    from gist_b200._lib import check, ptr, stream_ptr
    torch.manual_seed(0)
    for shape in [(7, 5), (300, 1204), (64, 41)]:
        x = torch.randn(*shape, device='cuda') * torch.logspace(-6, 6, shape[1], device='cuda')
        lo = ops.split_tf32(x)
        hi = torch.empty_like(x)
        lo2 = torch.empty_like(x)
        check(_lib.load().gist_split_tf32_f32(ptr(x), x.stride(0), shape[0], shape[1], ptr(hi), hi.stride(0),
                                              ptr(lo2), lo2.stride(0), stream_ptr(x.device)), 'split')
        assert torch.equal(lo, lo2)
        hi_bits = hi.view(torch.int32)
        assert torch.equal(hi_bits, x.view(torch.int32) & ~0x1FFF)
        assert (lo.view(torch.int32) & 0x1FFF).eq(0).all()                 # lo is TF32-representable
        resid = x.double() - hi.double()
        assert ((lo.double() - resid).abs() <= resid.abs() * 2.0 ** -11 + 1e-45).all()
    z = torch.tensor([[float('inf'), float('nan'), 0.0, -0.0]], device='cuda')
    assert torch.equal(ops.split_tf32(z), torch.zeros_like(z))


@pytest.mark.parametrize('shape', LAYOUT_SHAPES + [(4096, 512, 1024)])
@pytest.mark.parametrize('a_mn', [False, True])
@pytest.mark.parametrize('b_mn', [False, True])
def test_gemm_3xtf32_layouts(shape, a_mn, b_mn):
    from gist_b200 import ops
    M, N, K = shape
    torch.manual_seed(M * 5 + N * 11 + K)

    def stored(rows, cols, mn):
        r, c = (cols, rows) if mn else (rows, cols)
        buf = torch.randn(r, (c + 3) // 4 * 4, device='cuda')[:, :c]
        return buf, (buf.t() if mn else buf)

    A_st, A = stored(M, K, a_mn)
    B_st, B = stored(N, K, b_mn)
    bias = torch.randn(N, device='cuda')
    ref = A.double() @ B.double().t()
    absref = A.abs().double() @ B.abs().double().t()
    A_lo, B_lo = ops.split_tf32(A_st), ops.split_tf32(B_st)
    for fl in (0, 2, 4, 8, 16):      # auto, no split-K, forced 64 / 128 / 256-wide tiles
        got = ops.gemm(A_st, B_st, a_mn=a_mn, b_mn=b_mn, A_lo=A_lo, B_lo=B_lo, flags=fl)
        rel = ((got.double() - ref).norm() / ref.norm()).item()
        assert rel < X3_TOL, (fl, rel)
        assert ((got.double() - ref).abs() <= absref * 2.0 ** -18 + 1e-30).all(), fl
    got2 = ops.gemm(A_st, B_st, a_mn=a_mn, b_mn=b_mn, A_lo=A_lo, B_lo=B_lo, bias=bias, relu=True)
    ref2 = torch.relu(ref + bias.double())
    assert ((got2.double() - ref2).norm() / ref2.norm()).item() < X3_TOL


def test_gemm_3xtf32_beats_fp32_sgemm_error_class():
    """Same error class as cuBLAS fp32 on a long contraction with badly scaled columns."""
    from gist_b200 import ops
    torch.manual_seed(3)
    A = torch.randn(512, 8192, device='cuda') * torch.logspace(-3, 3, 8192, device='cuda')
    B = torch.randn(256, 8192, device='cuda')
    ref = A.double() @ B.double().t()
    got = ops.gemm(A, B, A_lo=ops.split_tf32(A), B_lo=ops.split_tf32(B))
    rel3 = ((got.double() - ref).norm() / ref.norm()).item()
    rel1 = ((ops.gemm(A, B).double() - ref).norm() / ref.norm()).item()
    assert rel3 < 2e-6 and rel1 > 50 * rel3, (rel3, rel1)


def test_linear_3xtf32_autograd():
    from gist_b200 import ops
    torch.manual_seed(0)
    for n, fin, fout in [(1000, 1204, 256), (2586, 512, 41), (777, 64, 32)]:
        z = torch.randn(n, fin, device='cuda', requires_grad=True)
        W = (torch.randn(fout, fin, device='cuda') * 0.03).requires_grad_(True)
        b = torch.randn(fout, device='cuda', requires_grad=True)
        wy = torch.randn(n, fout, device='cuda')
        ops.set_matmul_precision('3xtf32')
        try:
            y = ops.linear(z, W, b)
            (y * wy).sum().backward()
        finally:
            ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
        z2, W2, b2 = (t.detach().double().requires_grad_(True) for t in (z, W, b))
        y2 = torch.nn.functional.linear(z2, W2, b2)
        (y2 * wy.double()).sum().backward()
        for got, ref, name in ((y, y2, 'y'), (z.grad, z2.grad, 'dz'), (W.grad, W2.grad, 'dW'),
                               (b.grad, b2.grad, 'db')):
            rel = ((got.double() - ref).norm() / ref.norm()).item()
            assert rel < X3_TOL, (name, rel)


# ---------------------------------------------------------------------------------------------
# Extended epilogue (gist_gemm_ex_f32): in-kernel split-K, row sums on the tensor core, LayerNorm
# ---------------------------------------------------------------------------------------------
def _stored(rows, cols, mn, gen):
    r, c = (cols, rows) if mn else (rows, cols)
    buf = gen(r, (c + 3) // 4 * 4)[:, :c]
    return buf, (buf.t() if mn else buf)


@pytest.mark.parametrize('shape', [(256, 1204, 2586), (41, 512, 2586), (2586, 256, 1204), (2586, 41, 512),
                                   (2590, 32, 1204), (300, 70, 1000), (129, 65, 700)])
@pytest.mark.parametrize('x3', [False, True])
@pytest.mark.parametrize('layout', [(False, False), (True, True), (False, True)])
def test_in_kernel_splitk_equals_two_kernel_splitk(shape, x3, layout, monkeypatch):
    """The last-arriver reduction adds the partial tiles in split order — the SAME order as
    splitk_reduce_kernel — so both forms give bit-identical results; repeated launches reuse the
    tile counters (left at zero); bias / ReLU run in the reducing CTA."""
    import ctypes
    from gist_b200 import _lib, ops
    if not _lib.load().gist_gemm_has_inkernel_splitk():
        pytest.skip('library built without -DGIST_GEMM_INKERNEL_SPLITK (the default)')
    M, N, K = shape
    a_mn, b_mn = layout
    torch.manual_seed(M + 3 * N + K)
    rnd = lambda r, c: torch.randn(r, c, device='cuda')      # noqa: E731
    A_st, A = _stored(M, K, a_mn, rnd)
    B_st, B = _stored(N, K, b_mn, rnd)
    lo = dict(A_lo=ops.split_tf32(A_st), B_lo=ops.split_tf32(B_st)) if x3 else {}
    bias = torch.randn(N, device='cuda')
    tn, sp, kb = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.load().gist_gemm_plan(M, N, K, 0, 1 if x3 else 0, ctypes.byref(tn), ctypes.byref(sp), ctypes.byref(kb))
    monkeypatch.setattr(ops, 'FUSED_SPLITK', True)
    fused = [ops.gemm(A_st, B_st, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True, **lo) for _ in range(3)]
    assert (ops._gemm_counters(A_st.device) == 0).all()
    monkeypatch.setattr(ops, 'FUSED_SPLITK', False)
    two = ops.gemm(A_st, B_st, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True, **lo)
    assert torch.equal(fused[0], fused[1]) and torch.equal(fused[1], fused[2])
    ref = torch.relu(A.double() @ B.double().t() + bias.double())
    tol = 2e-5 if x3 else 2e-3
    assert (fused[0].double() - ref).norm() / ref.norm() < tol
    if sp.value > 1:
        # same split plan in both modes for these shapes? the planner's cost differs by mode, so compare
        # bit-for-bit only when it is; always compare numerically
        s2 = ctypes.c_int32()
        ex = _lib.GemmEx()
        ex.tile_counters, ex.n_counters = ops._gemm_counters(A_st.device).data_ptr(), 4096
        wsb_f = _lib.load().gist_gemm_ex_workspace_bytes(M, N, K, 0, 1 if x3 else 0, ctypes.byref(ex))
        assert wsb_f > 0
    assert (fused[0].double() - two.double()).abs().max() <= 1e-5 * ref.abs().max()


@pytest.mark.parametrize('shape', [(256, 1204, 2586), (41, 512, 2586), (256, 512, 2586), (32, 64, 2590), (47, 200, 2280),
                                   (300, 70, 1000)])
@pytest.mark.parametrize('x3', [False, True])
@pytest.mark.parametrize('inkernel', [False, True])
def test_rowsum_on_the_tensor_core_is_the_bias_gradient(shape, x3, inkernel, monkeypatch):
    """dW = dy^T z with db = colsum(dy) from the same launch: an extra 16-column MMA against a tile of
    ones per K step (in n-tile 0), reduced with the split-K partials."""
    from gist_b200 import _lib, ops
    if inkernel and not _lib.load().gist_gemm_has_inkernel_splitk():
        pytest.skip('library built without -DGIST_GEMM_INKERNEL_SPLITK (the default)')
    monkeypatch.setattr(ops, 'FUSED_SPLITK', inkernel)        # partial row sums folded in-kernel / by the second pass
    M, N, K = shape               # M = out features, N = in features, K = batch rows
    torch.manual_seed(M + N + K)
    dy = torch.randn(K, (M + 3) // 4 * 4, device='cuda')[:, :M]      # stored [K, M]: MN-major A
    z = torch.randn(K, (N + 3) // 4 * 4, device='cuda')[:, :N]
    lo = dict(A_lo=ops.split_tf32(dy), B_lo=ops.split_tf32(z)) if x3 else {}
    outs = [ops.gemm(dy, z, a_mn=True, b_mn=True, rowsum=True, **lo) for _ in range(2)]
    dW, db = outs[0]
    assert torch.equal(dW, outs[1][0]) and torch.equal(db, outs[1][1])
    plain = ops.gemm(dy, z, a_mn=True, b_mn=True, **lo)          # the product is unchanged (tile plans may differ)
    assert (dW - plain).abs().max() <= (2e-6 if x3 else 2e-3) * plain.abs().max()
    ref = dy.double().sum(0)
    scale = dy.abs().double().sum(0).max()
    assert ((db.double() - ref).abs() <= (2e-6 if x3 else 2e-3) * scale).all(), (db.double() - ref).abs().max().item()
    if x3:
        assert (db.double() - ref).norm() / ref.norm() < 1e-5
    ones = torch.ones(K, (M + 3) // 4 * 4, device='cuda')[:, :M]
    lo1 = dict(A_lo=ops.split_tf32(ones), B_lo=lo['B_lo']) if x3 else {}
    assert torch.equal(ops.gemm(ones, z, a_mn=True, b_mn=True, rowsum=True, **lo1)[1], torch.full((M,), float(K), device='cuda'))


@pytest.mark.parametrize('shape', [(2590, 32, 1204), (2590, 32, 64), (2586, 64, 128), (2586, 128, 256), (1000, 41, 512),
                                   (300, 100, 72), (129, 7, 36),
                                   # rows of 129..256: the norm rides in the split-K second pass (K split) or
                                   # runs as the row kernel behind the GEMM (K not split)
                                   (2141, 256, 1204), (2141, 256, 512), (2590, 200, 1204), (300, 132, 64), (20000, 256, 96)])
@pytest.mark.parametrize('relu', [True, False])
@pytest.mark.parametrize('inkernel', [False, True])
def test_layernorm_epilogue_matches_the_row_kernel(shape, relu, inkernel, monkeypatch):
    """y = act(LN(z W^T + b)) from the projection's epilogue (one tile holds the row) vs the projection
    followed by ln_act_fwd_kernel, and vs fp64; the saved (pre-norm, stats) drive the same backward."""
    from gist_b200 import _lib, ops
    if inkernel and not _lib.load().gist_gemm_has_inkernel_splitk():
        pytest.skip('library built without -DGIST_GEMM_INKERNEL_SPLITK (the default)')
    monkeypatch.setattr(ops, 'FUSED_SPLITK', inkernel)        # split shapes: LN in the last-arriver fold / in the second pass
    M, N, K = shape
    torch.manual_seed(M + N + K)
    z = torch.randn(M, (K + 3) // 4 * 4, device='cuda')[:, :K]
    W = torch.randn(N, (K + 3) // 4 * 4, device='cuda')[:, :K] / K ** 0.5
    b = torch.randn(N, device='cuda')
    z_lo, W_lo = ops.split_tf32(z), ops.split_tf32(W)
    assert ops.ln_fusable(M, N, True)
    x_pre, y, stats = ops.gemm(z, W, bias=b, A_lo=z_lo, B_lo=W_lo, ln=(1e-5, relu))
    x_ref = ops.gemm(z, W, bias=b, A_lo=z_lo, B_lo=W_lo)
    assert (x_pre - x_ref).abs().max() <= 2e-6 * x_ref.abs().max()         # split plans may differ (forced tile)
    y_ref, stats_ref = ops._ln_fwd_raw(x_pre, 1e-5, relu)
    assert (y - y_ref).abs().max() <= 2e-6 * max(y_ref.abs().max().item(), 1.0)
    assert (stats - stats_ref).abs().max() <= 2e-6 * stats_ref.abs().max()
    x64 = z.double() @ W.double().t() + b.double()
    y64 = torch.nn.functional.layer_norm(x64, (N,), None, None, 1e-5)
    if relu:
        y64 = torch.relu(y64)
    assert (y.double() - y64).abs().max() <= 1e-5 * max(y64.abs().max().item(), 1.0)
    assert y.stride(0) % 4 == 0 and x_pre.stride(0) % 4 == 0               # TMA-addressable for the next GEMM
    again = ops.gemm(z, W, bias=b, A_lo=z_lo, B_lo=W_lo, ln=(1e-5, relu))
    assert torch.equal(again[1], y) and torch.equal(again[0], x_pre)


def test_fused_gemm_epilogues_on_parallel_streams_do_not_share_counters(monkeypatch):
    """The weight-gradient branch runs split-K GEMMs beside the training branch: tile counters are per stream."""
    from gist_b200 import ops
    monkeypatch.setattr(ops, 'FUSED_SPLITK', True)        # (a no-op in the default build: counters are then unused)
    torch.manual_seed(0)
    dy = torch.randn(2586, 256, device='cuda')
    z = torch.randn(2586, 1204, device='cuda')
    dy_lo, z_lo = ops.split_tf32(dy), ops.split_tf32(z)
    ref = ops.gemm(dy, z, a_mn=True, b_mn=True, A_lo=dy_lo, B_lo=z_lo, rowsum=True)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    for _ in range(4):
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(ops.gemm(dy, z, a_mn=True, b_mn=True, A_lo=dy_lo, B_lo=z_lo, rowsum=True))
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1])
