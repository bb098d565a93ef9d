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
    ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)


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
    assert_close(out, ref, rtol=1e-5, what='logits')          # north_star: 1e-5 relative
    for l, (w, b) in zip(model.layers, params):
        assert_close(l.linear.weight.grad, w.grad, rtol=1e-5, what='dW')
        assert_close(l.linear.bias.grad, b.grad, rtol=1e-5, what='db')


def test_ist_sage_layer_train_mode_injected_dropout():
    """Train-mode parity with the SAME dropout mask: torch's Philox stream cannot be
    reproduced on the CPU, so the mask is captured from the CUDA run and injected
    into the oracle."""
    from gist_b200 import ISTSAGELayer, ops
    n, fin, fout = 500, 64, 32
    g, og = _graphs(n, 6000, seed=4, loops=False)
    torch.manual_seed(0)
    layer = ISTSAGELayer(fin, fout, 0.5, True, activation=F.relu).cuda().train()
    masks = []
    layer.dropout.register_forward_hook(lambda m, i, o: masks.append((o / i[0]).nan_to_num(0.0)))
    x = torch.randn(n, fin)
    # the nn.Dropout module itself only runs in the cuBLAS cross-check mode; on the default tensor-core
    # path dropout is fused into K1 / K4 and is covered, mask for mask, by tests/test_gpu_dropout_fusion.py
    ops.set_matmul_precision('fp32')
    try:
        out = layer(g, x.cuda())
    finally:
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
    mask = masks[0].cpu().double()
    # entries where the input was exactly 0 give 0/0 -> 0 in the mask; harmless (z*mask = 0)
    ref = O.ist_sage_layer(og, x.double(), layer.linear.weight.detach().double().cpu(),
                           layer.linear.bias.detach().double().cpu(), True, F.relu, mask)
    assert_close(out, ref, rtol=1e-5, what='train-mode layer')


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
    assert_close(out, ref, rtol=1e-5, what='GraphConv fwd')
    assert_close(xg.grad, x64.grad, rtol=1e-5, what='dX')
    assert_close(conv.weight.grad, w64.grad, rtol=1e-5, what='dW')
    assert_close(conv.bias.grad, b64.grad, rtol=1e-5, what='db')


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
    assert_close(out, ref, rtol=1e-5, what='GCN(GraphConv) logits')
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
        ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)


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
    assert_close(fast, ref, rtol=1e-5, what='project-first logits')
    assert_close(fast, slow, rtol=1e-5, what='project-first vs aggregate-first')


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


# ---------------------------------------------------------------------------------------------
# GraphSAGELayer / GraphSAGE / BaselineGCN (cluster_gcn/modules.py:100-189, 316-349): golden vectors
# recorded from the reference's own classes (oracle/gen_golden.py:gen_graphsage) and larger random
# cases against the oracle restatement.
# ---------------------------------------------------------------------------------------------
import os

import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _gold(name):
    return np.load(os.path.join(_GOLD, name + '.npz'))


def _T(x):
    return torch.from_numpy(np.asarray(x))


def _load_state(module, G, prefix):
    sd = {k: _T(G[prefix + k]) for k in module.state_dict().keys()}
    module.load_state_dict(sd)


@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_graphsage_layer_vs_reference_golden(ci, matmul_precision):
    from gist_b200 import GistGraph, GraphSAGELayer
    G = _gold('graphsage')
    p = 'gl%d_' % ci
    fin, fout, has_bias, pp, ln, train, act = (int(v) for v in G[p + 'cfg'])
    g = GistGraph.from_edges(_T(G[p + 'src']), _T(G[p + 'dst']), 48, device='cuda')
    layer = GraphSAGELayer(fin, fout, F.relu if act else None, 0.0, bias=bool(has_bias), use_pp=bool(pp),
                           use_lynorm=bool(ln))
    _load_state(layer, G, p + 'param.')
    layer = layer.cuda().train(bool(train))
    x = _T(G[p + 'x']).cuda().requires_grad_(True)
    y = layer(g, x)
    (y * _T(G[p + 'wy']).cuda()).sum().backward()
    assert_close(y, _T(G[p + 'out']), rtol=1e-5, what='GraphSAGELayer out')
    assert_close(x.grad, _T(G[p + 'dx']), rtol=1e-5, what='dx')
    for k, v in layer.named_parameters():
        assert_close(v.grad, _T(G[p + 'grad.' + k]), rtol=1e-5, what='grad ' + k)


@pytest.mark.parametrize('ci', [0, 1])
def test_graphsage_container_vs_reference_golden(ci, matmul_precision):
    from gist_b200 import GistGraph, GraphSAGE
    G = _gold('graphsage')
    p = 'gs%d_' % ci
    fin, hid, ncls, L, pp = (int(v) for v in G[p + 'cfg'])
    n = int(G[p + 'n'])
    g = GistGraph.from_edges(_T(G[p + 'src']), _T(G[p + 'dst']), n, device='cuda')
    model = GraphSAGE(fin, hid, ncls, L, F.relu, 0.4, bool(pp))
    _load_state(model, G, p + 'param.')
    model = model.cuda().eval()
    g.ndata['feat'] = _T(G[p + 'x']).cuda()
    logits = model(g)
    F.cross_entropy(logits, _T(G[p + 'y']).cuda()).backward()
    assert_close(logits, _T(G[p + 'logits']), rtol=1e-5, what='GraphSAGE logits')
    for k, v in model.named_parameters():
        assert_close(v.grad, _T(G[p + 'grad.' + k]), rtol=1e-5, what='grad ' + k)
    assert list(model.state_dict().keys())[:4] == ['layers.0.linear.weight', 'layers.0.linear.bias',
                                                   'layers.0.lynorm.weight', 'layers.0.lynorm.bias']


@pytest.mark.parametrize('cfg', [(602, 128, 41, 2), (100, 64, 47, 1)])
def test_graphsage_container_vs_oracle(cfg, matmul_precision):
    from gist_b200 import GraphSAGE
    fin, hid, ncls, L = cfg
    n = 900
    g, og = _graphs(n, 14000, seed=20 + L, loops=False)
    torch.manual_seed(3)
    model = GraphSAGE(fin, hid, ncls, L, F.relu, 0.3, False).cuda().eval()
    with torch.no_grad():
        for l in model.layers[:-1]:
            l.lynorm.weight.uniform_(0.5, 1.5)
            l.lynorm.bias.uniform_(-0.3, 0.3)
    x = torch.randn(n, fin)
    y = torch.randint(0, ncls, (n,))
    g.ndata['feat'] = x.cuda()
    params = []
    for i, l in enumerate(model.layers):
        t = [l.linear.weight, l.linear.bias] + ([l.lynorm.weight, l.lynorm.bias] if i < L else [None, None])
        params.append(tuple(v.detach().double().cpu().requires_grad_(True) if v is not None else None for v in t))
    ref = O.graphsage_forward(og, x.double(), params)
    F.cross_entropy(ref, y).backward()
    out = model(g)
    F.cross_entropy(out, y.cuda()).backward()
    assert_close(out, ref, rtol=1e-5, what='GraphSAGE logits')
    for l, tup in zip(model.layers, params):
        assert_close(l.linear.weight.grad, tup[0].grad, rtol=1e-5, what='dW')
        assert_close(l.linear.bias.grad, tup[1].grad, rtol=1e-5, what='db')
        if tup[2] is not None:
            assert_close(l.lynorm.weight.grad, tup[2].grad, rtol=1e-5, what='d ln.weight')
            assert_close(l.lynorm.bias.grad, tup[3].grad, rtol=1e-5, what='d ln.bias')


def test_graphsage_layer_use_pp_train_skips_aggregation():
    """use_pp=True in training mode: the layer consumes pre-aggregated [h ‖ ah] input and must not
    aggregate again (modules.py:133); in eval mode it aggregates."""
    from gist_b200 import GraphSAGELayer
    n, fin, fout = 300, 10, 6
    g, og = _graphs(n, 3000, seed=5, loops=False)
    torch.manual_seed(0)
    layer = GraphSAGELayer(fin, fout, None, 0.0, use_pp=True, use_lynorm=False).cuda()
    W, b = layer.linear.weight.detach().double().cpu(), layer.linear.bias.detach().double().cpu()
    x2 = torch.randn(n, 2 * fin)
    layer.train()
    assert_close(layer(g, x2.cuda()), F.linear(x2.double(), W, b), rtol=1e-5, what='use_pp train')
    layer.eval()
    x = torch.randn(n, fin)
    ref = O.graphsage_layer(og, x.double(), W, b, use_lynorm=False)
    assert_close(layer(g, x.cuda()), ref, rtol=1e-5, what='use_pp eval')


def test_baseline_gcn_vs_reference_golden_forward(matmul_precision):
    """graphconv.npz:base_* — the reference's BaselineGCN logits (recorded in round 1, until now only
    used to check the oracle)."""
    from gist_b200 import BaselineGCN, GistGraph
    G = _gold('graphconv')
    g = GistGraph.from_edges(_T(G['base_src']), _T(G['base_dst']), 50, device='cuda')
    model = BaselineGCN(9, 12, 4, 2, F.relu, 0.5, True)
    _load_state(model, G, 'base_param.')
    model = model.cuda().eval()
    g.ndata['feat'] = _T(G['base_x']).cuda()
    assert_close(model(g), _T(G['base_logits']), rtol=1e-5, what='BaselineGCN logits')


@pytest.mark.parametrize('ci', [0, 1])
def test_baseline_gcn_vs_reference_golden_with_grads(ci, matmul_precision):
    from gist_b200 import BaselineGCN, GistGraph
    G = _gold('graphsage')
    p = 'bg%d_' % ci
    fin, hid, ncls, L, ln = (int(v) for v in G[p + 'cfg'])
    n = int(G[p + 'n'])
    g = GistGraph.from_edges(_T(G[p + 'src']), _T(G[p + 'dst']), n, device='cuda')
    model = BaselineGCN(fin, hid, ncls, L, F.relu, 0.5, bool(ln))
    _load_state(model, G, p + 'param.')
    model = model.cuda().eval()
    g.ndata['feat'] = _T(G[p + 'x']).cuda()
    logits = model(g)
    F.cross_entropy(logits, _T(G[p + 'y']).cuda()).backward()
    assert_close(logits, _T(G[p + 'logits']), rtol=1e-5, what='BaselineGCN logits')
    for k, v in model.named_parameters():
        assert_close(v.grad, _T(G[p + 'grad.' + k]), rtol=1e-5, what='grad ' + k)


@pytest.mark.parametrize('cfg', [(602, 128, 41, 2, True), (64, 96, 7, 1, False)])
def test_baseline_gcn_vs_oracle(cfg, matmul_precision):
    from gist_b200 import BaselineGCN
    fin, hid, ncls, L, ln = cfg
    n = 1000
    g, og = _graphs(n, 16000, seed=30 + L)
    torch.manual_seed(4)
    model = BaselineGCN(fin, hid, ncls, L, F.relu, 0.5, ln).cuda().eval()
    with torch.no_grad():
        for l in model.layers:
            l.bias.uniform_(-0.3, 0.3)
    x = torch.randn(n, fin)
    y = torch.randint(0, ncls, (n,))
    g.ndata['feat'] = x.cuda()
    params = [(l.weight.detach().double().cpu().requires_grad_(True),
               l.bias.detach().double().cpu().requires_grad_(True)) for l in model.layers]
    ref = O.graphconv_gcn_forward(og, x.double(), params, ln)
    F.cross_entropy(ref, y).backward()
    out = model(g)
    F.cross_entropy(out, y.cuda()).backward()
    assert_close(out, ref, rtol=1e-5, what='BaselineGCN logits')
    for l, (w, b) in zip(model.layers, params):
        assert_close(l.weight.grad, w.grad, rtol=1e-5, what='dW')
        assert_close(l.bias.grad, b.grad, rtol=1e-5, what='db')
