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
    n, fin, fout = 1000, 1204, 256
    z = torch.randn(n, fin, device='cuda', requires_grad=True)
    W = (torch.randn(fout, fin, device='cuda') * 0.03).requires_grad_(True)
    b = torch.randn(fout, device='cuda', requires_grad=True)
    wy = torch.randn(n, fout, device='cuda')
    ops.set_matmul_precision('tf32')
    try:
        y = ops.linear(z, W, b)
        (y * wy).sum().backward()
    finally:
        ops.set_matmul_precision('fp32')
    z2, W2, b2 = (t.detach().double().requires_grad_(True) for t in (z, W, b))
    y2 = torch.nn.functional.linear(z2, W2, b2)
    (y2 * wy.double()).sum().backward()
    for got, ref, name in ((y, y2, 'y'), (z.grad, z2.grad, 'dz'), (W.grad, W2.grad, 'dW'), (b.grad, b2.grad, 'db')):
        rel = ((got.double() - ref).norm() / ref.norm()).item()
        assert rel < 2e-3, (name, rel)
