"""Fused row-wise kernels of the training step (csrc/fused.cu) through the C-ABI against torch
fp64 references of the same ops.  Tolerance: fp32 round-off of a few-hundred-term row reduction,
asserted at 2e-6 relative (norm-wise), written next to each check."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return ((got.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('n,d', [(1, 4), (2586, 256), (2586, 32), (1000, 41), (333, 100), (64, 4096), (7, 1)])
@pytest.mark.parametrize('relu', [False, True])
def test_layer_norm_act(n, d, relu):
    from gist_b200 import ops
    torch.manual_seed(n + d)
    x = (torch.randn(n, d, device='cuda') * 3 + 1).requires_grad_(True)
    w = torch.randn(n, d, device='cuda')
    y = ops.layer_norm_act(x, 1e-5, relu=relu)
    (y * w).sum().backward()
    x2 = x.detach().double().requires_grad_(True)
    y2 = F.layer_norm(x2, (d,), eps=1e-5)
    if relu:
        y2 = F.relu(y2)
    (y2 * w.double()).sum().backward()
    if d > 1:
        assert _rel(y, y2) < 2e-6
        assert _rel(x.grad, x2.grad) < 2e-5     # LN backward cancels: looser
    else:
        assert torch.equal(y, y2.float())


def test_layer_norm_act_strided_views():
    from gist_b200 import ops
    buf = torch.randn(500, 300, device='cuda')
    x = buf[:, 7:107]                   # unaligned column block of a wider buffer
    y = ops.layer_norm_act(x, 1e-5, relu=True)
    ref = F.relu(F.layer_norm(x.double(), (100,), eps=1e-5))
    assert _rel(y, ref) < 2e-6


@pytest.mark.parametrize('n,d', [(2586, 256), (2586, 41), (1, 5), (100000, 7), (4096, 4096), (63, 33)])
def test_colsum(n, d):
    from gist_b200 import ops
    torch.manual_seed(n)
    buf = torch.randn(n, d + 3, device='cuda')
    x = buf[:, :d]
    got = ops.colsum(x)
    ref = x.double().sum(0)
    scale = x.double().abs().sum(0).max().item()
    assert (got.double() - ref).abs().max().item() <= 1e-6 * scale
    assert torch.equal(got, ops.colsum(x))          # deterministic
    ones = torch.ones(n, d, device='cuda')
    if n < (1 << 24):
        assert torch.equal(ops.colsum(ones), torch.full((d,), float(n), device='cuda'))


@pytest.mark.parametrize('n,C', [(2586, 41), (1000, 47), (17, 3), (5000, 7), (300, 1000)])
@pytest.mark.parametrize('use_mask', [True, False])
def test_masked_cross_entropy(n, C, use_mask):
    from gist_b200 import ops
    torch.manual_seed(n + C)
    logits = (torch.randn(n, C, device='cuda') * 4).requires_grad_(True)
    labels = torch.randint(0, C, (n,), device='cuda')
    mask = (torch.rand(n, device='cuda') < 0.6) if use_mask else None
    loss = ops.masked_cross_entropy(logits, labels, mask)
    (loss * 1.7).backward()
    l2 = logits.detach().double().requires_grad_(True)
    sel = mask if use_mask else torch.ones(n, dtype=torch.bool, device='cuda')
    ref = F.cross_entropy(l2[sel], labels[sel])
    (ref * 1.7).backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert _rel(logits.grad, l2.grad) < 2e-6
    assert logits.grad.shape == (n, C)
    if use_mask:
        assert (logits.grad[~mask] == 0).all()


@pytest.mark.parametrize('mode', ['fp32', '3xtf32'])
@pytest.mark.parametrize('n,C', [(2586, 41), (1000, 47), (17, 3), (5000, 7), (300, 1000), (1, 5)])
@pytest.mark.parametrize('use_mask', [True, False, 1, 3])      # 1, 3: mask at a byte offset (unaligned words)
def test_masked_ce_loss_and_grad_one_launch(n, C, use_mask, mode):
    """The trainers' fused loss + gradient seed (ops.masked_ce_loss_and_grad) against fp64 torch and
    against the two-pass autograd op; repeated launches reuse the arrival counter (left at zero)."""
    from gist_b200 import ops
    torch.manual_seed(n + C)
    logits = torch.randn(n, C, device='cuda') * 4
    labels = torch.randint(0, C, (n,), device='cuda')
    mask = (torch.rand(n + 3, device='cuda') < 0.6)[int(use_mask) % 4 if use_mask is not True else 0:][:n] \
        if use_mask else None
    if use_mask:
        mask[0] = True
        assert mask.is_contiguous() and mask.data_ptr() % 4 == (0 if use_mask is True else use_mask)
    old = ops.get_matmul_precision()
    ops.set_matmul_precision(mode)
    try:
        for _ in range(3):
            loss, dl = ops.masked_ce_loss_and_grad(logits, labels, mask)
        lo = ops._lo_take(dl)
        assert (lo is not None) == (mode == '3xtf32')
    finally:
        ops.set_matmul_precision(old)
    l2 = logits.double().requires_grad_(True)
    sel = mask if use_mask else torch.ones(n, dtype=torch.bool, device='cuda')
    ref = F.cross_entropy(l2[sel], labels[sel])
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert dl.shape == (n, C) and _rel(dl, l2.grad) < 2e-6
    if use_mask:
        assert (dl[~mask] == 0).all()
    l3 = logits.clone().requires_grad_(True)
    ops.masked_cross_entropy(l3, labels, mask).backward()
    assert torch.equal(dl, l3.grad)                    # same arithmetic as the two-pass kernels
    if lo is not None:                                 # x = trunc_tf32(x) + lo up to tf32 rounding of lo
        hi = (dl.view(torch.int32) & ~0x1fff).view(torch.float32)
        assert ((dl - hi - lo).abs() <= 2.0 ** -10 * (dl - hi).abs() + 1e-45).all()
    assert dl.stride(0) % 4 == 0 and dl.data_ptr() % 16 == 0       # TMA-addressable, padding zeroed
    pad = torch.as_strided(dl, (n, dl.stride(0)), (dl.stride(0), 1))[:, C:]
    assert (pad == 0).all()


@pytest.mark.parametrize('use_mask', [True, False])
def test_cross_entropy_ignores_rows_with_out_of_range_labels(use_mask):
    """torch.nn.CrossEntropyLoss (…distrib.py:383) excludes ignore_index = -100 rows from the mean's
    denominator and gives them a zero gradient.  Both CE paths treat every label outside [0, C) that
    way (the one-launch kernel and the two-pass autograd op)."""
    from gist_b200 import ops
    torch.manual_seed(5)
    n, C = 1500, 41
    logits = torch.randn(n, C, device='cuda') * 3
    labels = torch.randint(0, C, (n,), device='cuda')
    ign = torch.rand(n, device='cuda') < 0.2
    labels[ign] = -100
    mask = (torch.rand(n, device='cuda') < 0.7) if use_mask else None
    sel = mask if use_mask else torch.ones(n, dtype=torch.bool, device='cuda')
    l2 = logits.double().requires_grad_(True)
    ref = F.cross_entropy(l2[sel], labels[sel], ignore_index=-100)
    ref.backward()
    loss, dl = ops.masked_ce_loss_and_grad(logits, labels, mask)
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert _rel(dl, l2.grad) < 2e-6 and (dl[ign] == 0).all()
    l3 = logits.clone().requires_grad_(True)
    loss3 = ops.masked_cross_entropy(l3, labels, mask)
    loss3.backward()
    assert abs(loss3.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert torch.equal(dl, l3.grad)
    labels[~ign] = C + 3                       # other invalid labels: same treatment, nothing counts
    loss, dl = ops.masked_ce_loss_and_grad(logits, labels, mask)
    assert torch.isnan(loss) and (dl == 0).all()


def test_adam_matches_torch():
    from gist_b200.optim import Adam
    torch.manual_seed(0)
    shapes = [(256, 1204), (256,), (256, 512), (256,), (41, 512), (41,), (3, 5), (1,)] + [(17,)] * 30
    ps = [torch.randn(*s, device='cuda').requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().double().requires_grad_(True) for p in ps]
    a = Adam(ps, lr=1e-2, weight_decay=5e-4)
    b = torch.optim.Adam(qs, lr=1e-2, weight_decay=5e-4)
    for it in range(12):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p) * (0.1 + it)
            p.grad = g
            q.grad = g.double()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        assert _rel(p, q) < 2e-6
    # reset_state == a fresh optimizer
    a.reset_state()
    c = torch.optim.Adam(qs, lr=1e-2, weight_decay=5e-4)
    for p, q in zip(ps, qs):
        q.data.copy_(p.data.double())
        g = torch.randn_like(p)
        p.grad = g
        q.grad = g.double()
    a.step()
    c.step()
    for p, q in zip(ps, qs):
        assert _rel(p, q) < 2e-6


def test_adam_cuda_graph_replay_advances_step():
    from gist_b200.optim import Adam
    torch.manual_seed(1)
    p = torch.randn(1000, device='cuda').requires_grad_(True)
    q = p.detach().clone().double().requires_grad_(True)
    g = torch.randn(1000, device='cuda')
    p.grad = g.clone()
    a = Adam([p], lr=1e-2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a.step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a.step()
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    b = torch.optim.Adam([q], lr=1e-2)
    for _ in range(6):       # 1 warm-up step + 5 replays (capture itself does not execute)
        q.grad = g.double()
        b.step()
    assert _rel(p, q) < 2e-6


@pytest.mark.parametrize('graphed', [False, True])
def test_adam_fused_tail_writes_low_halves_and_ticks_the_clock(graphed):
    """gist_adam_multi_ex_f32: the update itself is bit-identical to the plain launch, lo = tf32_lo(updated p)
    (what gist_split_tf32_multi_f32 computes at the head of the next step) and the clock advances by one per
    step — eagerly and from a replayed graph."""
    from gist_b200 import ops
    from gist_b200.optim import Adam
    torch.manual_seed(3)
    shapes = [(256, 1204), (256,), (32, 64), (41, 512), (41,), (3, 8)]
    ps = [torch.randn(*s, device='cuda').requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    a, b = Adam(ps, lr=1e-2, weight_decay=5e-4), Adam(qs, lr=1e-2, weight_decay=5e-4)
    los = {p: torch.full_like(p, float('nan')).detach() for p in ps if p.dim() == 2}
    a.lo_map = dict(los)
    a.tick = torch.full((1,), 7, dtype=torch.int64, device='cuda')
    gs = [torch.randn_like(p) for p in ps]
    for p, q, g in zip(ps, qs, gs):
        p.grad, q.grad = g.clone(), g.clone()
    n_steps = 4
    if graphed:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            a.step()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            a.step()
        for _ in range(n_steps - 1):
            graph.replay()
    else:
        for _ in range(n_steps):
            a.step()
    for _ in range(n_steps):
        b.step()
    torch.cuda.synchronize()
    assert int(a.tick.item()) == 7 + n_steps
    for p, q in zip(ps, qs):
        assert torch.equal(p, q)
    for p, lo in los.items():
        assert torch.equal(lo, ops.split_tf32(p.detach()))


def test_masked_ce_writes_the_loss_into_the_callers_slot():
    from gist_b200 import ops
    torch.manual_seed(4)
    logits = torch.randn(2141, 41, device='cuda') * 3
    labels = torch.randint(0, 41, (2141,), device='cuda')
    mask = torch.rand(2141, device='cuda') < 0.6
    ref, dref = ops.masked_ce_loss_and_grad(logits, labels, mask)
    slot = torch.zeros(2, device='cuda')
    loss, dl = ops.masked_ce_loss_and_grad(logits, labels, mask, out=slot)
    assert loss.data_ptr() == slot.data_ptr() and torch.equal(loss, ref) and torch.equal(dl, dref)
    assert abs(slot[1].item() - 1.0 / int(mask.sum())) < 1e-9
