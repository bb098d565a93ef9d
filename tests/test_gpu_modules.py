"""Module-level parity: drop-in layers / containers on the CUDA path vs the
functional CPU oracle (forward and gradients), eval mode and injected dropout."""
import pytest
import torch
import torch.nn.functional as F

from oracle import gist_oracle as O
from tests.util import assert_close, ograph, random_graph

pytestmark = pytest.mark.gpu


def _graphs(n, nnz, seed, loops=True):
    from gist_b200 import GistGraph
    src, dst = random_graph(n, nnz, seed=seed)
    if loops:      # GraphConv refuses zero-in-degree nodes, as DGL does
        ar = torch.arange(n)
        src, dst = torch.cat([src, ar]), torch.cat([dst, ar])
    return GistGraph.from_edges(src, dst, n, device='cuda'), ograph(src, dst, n)


def _params64(model):
    return [(l.linear.weight.detach().double().cpu().requires_grad_(True),
             l.linear.bias.detach().double().cpu().requires_grad_(True)) for l in model.layers]


@pytest.fixture(params=['fp32', '3xtf32'])
def matmul_precision(request):
    """Both fp32-accurate GEMM back ends must meet the same tolerance: cuBLAS sgemm and the
    tcgen05 kernel in 3xTF32 mode (what bench.py runs)."""
    from gist_b200 import ops
    ops.set_matmul_precision(request.param)
    yield request.param
    ops.set_matmul_precision('fp32')


@pytest.mark.parametrize('cfg', [(602, 256, 41, 2, True), (100, 64, 47, 3, True), (50, 32, 5, 1, False)])
def test_sage_gcn_forward_backward(cfg, matmul_precision):
    from gist_b200 import SageGCN
    fin, hid, ncls, L, ln = cfg
    n = 800
    g, og = _graphs(n, 12000, seed=L, loops=False)
    torch.manual_seed(0)
    model = SageGCN(fin, hid, ncls, L, F.relu, 0.0, ln, False, False, 1, True).cuda()
    x = torch.randn(n, fin)
    y = torch.randint(0, ncls, (n,))
    g.ndata['feat'] = x.cuda()
    params = _params64(model)
    ref = O.sage_gcn_forward(og, x.double(), params, ln)
    F.cross_entropy(ref, y).backward()
    out = model(g)
    F.cross_entropy(out, y.cuda()).backward()
    assert_close(out, ref, rtol=5e-5, what='logits')
    for l, (w, b) in zip(model.layers, params):
        assert_close(l.linear.weight.grad, w.grad, rtol=1e-4, what='dW')
        assert_close(l.linear.bias.grad, b.grad, rtol=1e-4, what='db')


def test_ist_sage_layer_train_mode_injected_dropout():
    """Train-mode parity with the SAME dropout mask: torch's Philox stream cannot be
    reproduced on the CPU, so the mask is captured from the CUDA run and injected
    into the oracle."""
    from gist_b200 import ISTSAGELayer
    n, fin, fout = 500, 64, 32
    g, og = _graphs(n, 6000, seed=4, loops=False)
    torch.manual_seed(0)
    layer = ISTSAGELayer(fin, fout, 0.5, True, activation=F.relu).cuda().train()
    masks = []
    layer.dropout.register_forward_hook(lambda m, i, o: masks.append((o / i[0]).nan_to_num(0.0)))
    x = torch.randn(n, fin)
    out = layer(g, x.cuda())
    mask = masks[0].cpu().double()
    # entries where the input was exactly 0 give 0/0 -> 0 in the mask; harmless (z*mask = 0)
    ref = O.ist_sage_layer(og, x.double(), layer.linear.weight.detach().double().cpu(),
                           layer.linear.bias.detach().double().cpu(), True, F.relu, mask)
    assert_close(out, ref, rtol=5e-5, what='train-mode layer')


@pytest.mark.parametrize('dims', [(1433, 16), (32, 32), (16, 7), (8, 24)])
@pytest.mark.parametrize('act', [None, 'relu', 'tanh'])
def test_graph_conv(dims, act):
    from gist_b200 import GraphConv
    fin, fout = dims
    n = 700
    g, og = _graphs(n, 8000, seed=fin)
    actf = {None: None, 'relu': F.relu, 'tanh': torch.tanh}[act]
    torch.manual_seed(0)
    conv = GraphConv(fin, fout, activation=actf).cuda()
    with torch.no_grad():
        conv.bias.uniform_(-0.5, 0.5)
    x = torch.randn(n, fin)
    w64 = conv.weight.detach().double().cpu().requires_grad_(True)
    b64 = conv.bias.detach().double().cpu().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    ref = O.graph_conv(og, x64, w64, b64, actf)
    wy = torch.randn(n, fout, dtype=torch.double)
    (ref * wy).sum().backward()
    xg = x.cuda().requires_grad_(True)
    out = conv(g, xg)
    (out * wy.float().cuda()).sum().backward()
    assert_close(out, ref, rtol=5e-5, what='GraphConv fwd')
    assert_close(xg.grad, x64.grad, rtol=1e-4, what='dX')
    assert_close(conv.weight.grad, w64.grad, rtol=1e-4, what='dW')
    assert_close(conv.bias.grad, b64.grad, rtol=1e-4, what='db')


def test_graph_conv_zero_in_degree_raises():
    from gist_b200 import GistGraph, GraphConv, GistError
    g = GistGraph.from_edges(torch.tensor([0, 1]), torch.tensor([1, 2]), 3, device='cuda')
    conv = GraphConv(4, 4).cuda()
    with pytest.raises(GistError):
        conv(g, torch.randn(3, 4, device='cuda'))


@pytest.mark.parametrize('split', [(False, False, 1), (False, True, 4), (True, True, 2)])
def test_graphconv_gcn_container(split):
    from gist_b200 import GCN
    si, so, k = split
    fin, hid, ncls, L = 48, 32, 3, 2
    n = 600
    g, og = _graphs(n, 7000, seed=9)
    torch.manual_seed(0)
    model = GCN(g, fin, hid, ncls, L, F.relu, 0.0, True, si, so, k).cuda()
    fin_eff = fin // k if si else fin
    x = torch.randn(n, fin_eff)
    params = [(l.weight.detach().double().cpu(), l.bias.detach().double().cpu()) for l in model.layers]
    ref = O.graphconv_gcn_forward(og, x.double(), params, True)
    out = model(x.cuda())
    assert_close(out, ref, rtol=5e-5, what='GCN(GraphConv) logits')
    assert list(model.state_dict().keys())[:2] == ['layers.0.weight', 'layers.0.bias']


def test_wrapper_dispatch_sync_single_process_matches_oracle():
    """m virtual sites simulated on one GPU: dispatch each site with the device K5
    gather, perturb, merge with the K5 scatter; compare with the oracle algebra."""
    import random
    from types import SimpleNamespace
    from gist_b200.ist import DistributedGNNWrapper, create_partition
    m, hid, L, fin, ncls = 4, 32, 2, 20, 5
    args = SimpleNamespace(rank=0, num_subnet=m, n_hidden=hid, n_layers=L, dropout=0.0,
                           use_layernorm=True)
    torch.manual_seed(0)
    w = DistributedGNNWrapper(args, None, fin, ncls, torch.device('cuda'))
    base0 = [(l.linear.weight.detach().cpu().clone(), l.linear.bias.detach().cpu().clone())
             for l in w.base_model.layers]
    random.seed(7)
    parts = [create_partition(m, hid) for _ in range(L)]
    random.seed(7)
    oparts = [O.create_partition(m, hid) for _ in range(L)]
    for a, b in zip(parts, oparts):
        for (i1, f1), (i2, f2) in zip(a, b):
            assert torch.equal(i1, i2) and torch.equal(f1, f2)
    dparts = w._to_dev(parts)
    flats, subs = [], []
    for site in range(m):
        w.args.rank = site
        w._dispatch_local(dparts)
        osub = O.sage_dispatch(base0, oparts, site)
        for lyr, (ow, ob) in zip(w.sub_model.layers, osub):
            assert torch.equal(lyr.linear.weight.detach().cpu(), ow)
            assert torch.equal(lyr.linear.bias.detach().cpu(), ob)
        # "train": deterministic perturbation
        trained = []
        for lyr in w.sub_model.layers:
            lyr.linear.weight.data = lyr.linear.weight.data * 1.5 + site
            lyr.linear.bias.data = lyr.linear.bias.data - 0.25 * (site + 1)
            trained.append((lyr.linear.weight.detach().cpu().clone(), lyr.linear.bias.detach().cpu().clone()))
        subs.append(trained)
        flats.append(w._pack().clone())
    w.args.rank = 0
    w.current_partition = dparts
    w._merge(torch.stack(flats), dparts)
    ref = O.sage_sync(base0, oparts, subs)
    for lyr, (rw, rb) in zip(w.base_model.layers, ref):
        assert torch.equal(lyr.linear.weight.detach().cpu(), rw)
        assert_close(lyr.linear.bias.detach(), rb, rtol=1e-6)


def test_weight_grad_branch_is_safe_when_the_side_stream_lags():
    """dW / db run on a low-priority side stream.  Their inputs (dy, z) are main-stream tensors that
    autograd drops as soon as the layer's backward returns; if the branch lags (here: a sleep
    queued on the side stream, in production the next batch's aggregation) the allocator must not
    recycle them under it.  Gradients must equal the single-stream run bit for bit."""
    from gist_b200 import SageGCN, ops
    n = 1500
    g, _ = _graphs(n, 30000, seed=9, loops=False)
    torch.manual_seed(1)
    model = SageGCN(128, 64, 10, 2, F.relu, 0.0, True, False, False, 1, True).cuda().train()
    g.ndata['feat'] = torch.randn(n, 128, device='cuda')
    y = torch.randint(0, 10, (n,), device='cuda')
    ops.set_matmul_precision('3xtf32')
    try:
        def grads(overlap, lag):
            ops.OVERLAP_WEIGHT_GRADS = overlap
            model.zero_grad(set_to_none=True)
            if lag:
                with torch.cuda.stream(ops._side_stream(torch.device('cuda', 0))):
                    torch.cuda._sleep(int(3e6))            # ~1.5 ms head start for the main stream
            F.cross_entropy(model(g), y).backward()
            junk = [torch.full((n, 128), float('nan'), device='cuda') for _ in range(8)]   # recycle freed blocks
            torch.cuda.synchronize()
            del junk
            return [p.grad.clone() for p in model.parameters()]
        ref = grads(False, False)
        for _ in range(3):
            got = grads(True, True)
            for a, b in zip(got, ref):
                assert torch.equal(a, b)
    finally:
        ops.OVERLAP_WEIGHT_GRADS = True
        ops.set_matmul_precision('fp32')


@pytest.mark.parametrize('cfg', [(602, 256, 41, 2, True), (100, 64, 47, 3, True), (50, 64, 5, 1, False)])
def test_sage_gcn_inference_project_first_vs_oracle(cfg, matmul_precision):
    """evaluate()'s no-grad eval forward takes the project-first form of every layer whose output is
    narrower than its input (ops.sage_project_first); it must meet the oracle like the training form,
    and agree with the aggregate-first forward of the same module."""
    from gist_b200 import SageGCN
    fin, hid, ncls, L, ln = cfg
    n = 900
    g, og = _graphs(n, 15000, seed=10 + L, loops=False)
    torch.manual_seed(1)
    model = SageGCN(fin, hid, ncls, L, F.relu, 0.3, ln, False, False, 1, True).cuda().eval()
    x = torch.randn(n, fin)
    g.ndata['feat'] = x.cuda()
    ref = O.sage_gcn_forward(og, x.double(), _params64(model), ln)
    with torch.no_grad():
        fast = model(g)
    slow = model(g)                         # grad enabled: aggregate-first path
    assert_close(fast, ref, rtol=5e-5, what='project-first logits')
    assert_close(fast, slow, rtol=5e-5, what='project-first vs aggregate-first')


def test_evaluate_masks_one_pass_equals_two_evaluate_calls():
    from gist_b200 import SageGCN
    from gist_b200.train import evaluate, evaluate_masks
    n = 1200
    g, _ = _graphs(n, 20000, seed=3, loops=False)
    torch.manual_seed(2)
    model = SageGCN(64, 32, 7, 2, F.relu, 0.2, True, False, False, 1, True).cuda()
    g.ndata['feat'] = torch.randn(n, 64, device='cuda')
    labels = torch.randint(0, 7, (n,), device='cuda')
    r = torch.rand(n, device='cuda')
    val, test, empty = r < 0.3, r > 0.6, torch.zeros(n, dtype=torch.bool, device='cuda')
    both = evaluate_masks(model, g, labels, [val, test, empty])
    assert both[0] == evaluate(model, g, labels, val)
    assert both[1] == evaluate(model, g, labels, test)
    assert both[2] == -1 == evaluate(model, g, labels, empty)
