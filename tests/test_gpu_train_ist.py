"""Config 2's path on the GPU: whole-tensor layer norm, the [in, out]-layout GraphConv matmul
and the gcn/train_ist.py trainer (gist_b200/train_ist.py) against

  * torch fp64 references of the same ops (tolerances written next to each check), and
  * tests/golden/train_ist.npz — what the reference's OWN gcn/train_ist.py main() split, trained
    and merged on the tiny dataset of tests/golden/train_ist_data.npz (oracle/gen_golden.py).
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _rel(got, ref):
    return ((got.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('n,d', [(19717, 32), (19717, 256), (2708, 16), (60, 8), (1, 1), (333, 7), (5, 4100)])
def test_tensor_layer_norm(n, d):
    from gist_b200 import ops
    torch.manual_seed(n * 7 + d)
    x = (torch.randn(n, d, device='cuda') * 2 + 0.5).requires_grad_(True)
    w = torch.randn(n, d, device='cuda')
    y = ops.tensor_layer_norm(x, 1e-5)
    (y * w).sum().backward()
    x2 = x.detach().double().requires_grad_(True)
    y2 = F.layer_norm(x2, x2.shape, eps=1e-5)
    (y2 * w.double()).sum().backward()
    if n * d > 1:
        assert _rel(y, y2) < 2e-6               # fp32 round-off of the normalisation itself
        assert _rel(x.grad, x2.grad) < 2e-5     # the backward cancels: looser
    else:
        assert torch.equal(y, y2.float())
    y_again = ops.tensor_layer_norm(x.detach(), 1e-5)
    assert torch.equal(y_again, y.detach())     # deterministic (fixed-order fold of the partials)


def test_tensor_layer_norm_strided_view():
    from gist_b200 import ops
    buf = torch.randn(700, 90, device='cuda')
    x = buf[:, 5:38]                           # unaligned, d = 33
    y = ops.tensor_layer_norm(x)
    assert _rel(y, F.layer_norm(x.double().contiguous(), (700, 33))) < 2e-6


@pytest.mark.parametrize('mode,tol', [('3xtf32', 1e-5), ('tf32', 3e-3), ('fp32', 1e-5)])
@pytest.mark.parametrize('n,fin,fout', [(19717, 496, 32), (2708, 1433, 16), (500, 32, 3), (61, 12, 8), (300, 7, 5)])
def test_graphconv_matmul_in_out_layout(mode, tol, n, fin, fout):
    """x @ W with W stored [in, out] (GraphConv's layout) and both gradients."""
    from gist_b200 import ops
    torch.manual_seed(n + fin + fout)
    old = ops.get_matmul_precision()
    ops.set_matmul_precision(mode)
    try:
        x = torch.randn(n, fin, device='cuda', requires_grad=True)
        W = (torch.randn(fin, fout, device='cuda') / fin ** 0.5).requires_grad_(True)
        gy = torch.randn(n, fout, device='cuda')
        y = ops.matmul(x, W)
        y.backward(gy)
        xd, Wd = x.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True)
        yd = xd @ Wd
        yd.backward(gy.double())
        assert _rel(y, yd) < tol
        assert _rel(x.grad, xd.grad) < tol
        assert _rel(W.grad, Wd.grad) < tol
    finally:
        ops.set_matmul_precision(old)


def _data():
    D = np.load(os.path.join(GOLD, 'train_ist_data.npz'))
    n = int(D['n'])
    tr, va, te = (np.zeros(n, dtype=bool) for _ in range(3))
    tr[:20], va[20:40], te[40:] = True, True, True       # the loader of oracle/gen_golden.py::gen_train_ist
    return SimpleNamespace(src=torch.from_numpy(D['src']), dst=torch.from_numpy(D['dst']),
                           features=torch.from_numpy(D['features']), labels=torch.from_numpy(D['labels']),
                           train_mask=torch.from_numpy(tr), val_mask=torch.from_numpy(va),
                           test_mask=torch.from_numpy(te), num_labels=int(D['num_labels']))


@pytest.mark.parametrize('use_graph', [False, True])
@pytest.mark.parametrize('mode', ['fp32', '3xtf32'])
@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_train_ist_trainer_vs_reference_golden(ci, mode, use_graph):
    """Same seed, same data: the trainer must draw the reference's partitions bit-exactly, start
    every round from the reference's weights, and after every round hold the reference's trained
    sub-models and merged model (fp32 training, 2 Adam steps per round: 2e-5 norm-wise)."""
    from gist_b200 import ops
    from gist_b200.graph import GistGraph
    from gist_b200.train_ist import ISTGCNTrainer
    G = np.load(os.path.join(GOLD, 'train_ist.npz'))
    p = 'ti%d_' % ci
    si, so, L, m, hid, fin, ncls = (int(v) for v in G[p + 'cfg'])
    keys = ['layers.%d.%s' % (l, t) for l in range(L + 1) for t in ('weight', 'bias')]
    d = _data()
    dev = torch.device('cuda')
    old = ops.get_matmul_precision()
    ops.set_matmul_precision(mode)
    try:
        args = SimpleNamespace(iter_per_site=2, num_subnet=m, dropout=0.0, split_output=str(bool(so)),
                               split_input=str(bool(si)), lr=0.01, n_epochs=4, n_hidden=hid, n_layers=L,
                               weight_decay=5e-4, use_layernorm='True')
        g = GistGraph.from_edges(d.src, d.dst, d.features.shape[0], device=dev)
        torch.manual_seed(ci)                      # gen_train_ist seeds right before ref.main(args)
        tr = ISTGCNTrainer(g, d.features.to(dev), d.labels.to(dev), d.train_mask.to(dev), ncls, args, dev,
                           use_graph=use_graph)
        nper = int(G[p + 'nperm']) // int(G[p + 'nrounds'])
        for r in range(int(G[p + 'nrounds'])):
            sd = tr.model.state_dict()
            for k in keys:                          # the round starts from the reference's full model
                ref = torch.from_numpy(G[p + 'r%d_main.%s' % (r, k)])
                if r == 0:
                    assert torch.equal(sd[k].cpu(), ref), k
                else:
                    assert _rel(sd[k].cpu(), ref) < 2e-5, (r, k)
            tr.train_epoch(2 * r)
            perms = [q for q in tr.feats_idx if q is not None]
            assert len(perms) == nper
            for i, chunks in enumerate(perms):      # partitions: bit-exact
                assert np.array_equal(torch.cat(list(chunks)).numpy(), G[p + 'perm%d' % (r * nper + i)])
            # merge happens at the end of the round's 2nd epoch; sub-models then hold 'trained'
            tr.train_epoch(2 * r + 1)
            for s in range(m):
                sub = tr.sub_models[s].state_dict()
                for k in keys:
                    ref = torch.from_numpy(G[p + 'r%d_trained%d.%s' % (r, s, k)])
                    assert sub[k].shape == ref.shape
                    assert _rel(sub[k].cpu(), ref) < 2e-5, (r, s, k, _rel(sub[k].cpu(), ref))
            sd = tr.model.state_dict()
            for k in keys:
                ref = torch.from_numpy(G[p + 'r%d_merged.%s' % (r, k)])
                assert _rel(sd[k].cpu(), ref) < 2e-5, (r, k)
    finally:
        ops.set_matmul_precision(old)


def test_train_ist_main_runs_pubmed_shape_slice():
    """main() end to end on a scaled PubMed-shaped graph with the config-2 flags
    (script/sweep.py: split_input False, split_output True, m = 8): loss decreases."""
    from gist_b200 import synth
    from gist_b200.train_ist import main
    ds = synth.make('pubmed', seed=0, scale=0.2)
    n = ds.num_nodes
    data = SimpleNamespace(src=ds.src, dst=ds.dst, features=ds.feat[:, :496], labels=ds.label,
                           train_mask=ds.train_mask, val_mask=ds.val_mask, test_mask=ds.test_mask,
                           num_labels=ds.num_classes)
    args = SimpleNamespace(iter_per_site=5, num_subnet=8, dropout=0.5, split_output='True', split_input='False',
                           lr=0.01, n_epochs=20, n_hidden=256, n_layers=2, weight_decay=5e-4,
                           use_layernorm='True', self_loop='True', use_random_proj='False')
    torch.manual_seed(0)
    losses = []
    out = main(args, data, device='cuda', log=lambda s: losses.append(float(s.split('Loss')[1].split('|')[0])))
    assert len(out.record) == 20 and n == data.features.shape[0]
    assert losses[-1] < losses[0]
    assert all(np.isfinite(losses))


def test_gcn_trainer_graph_replay_equals_eager():
    """gcn/train.py mirror: the CUDA-graph replayed step gives the eager step's weights (dropout 0)."""
    from gist_b200 import synth
    from gist_b200.graph import GistGraph
    from gist_b200.train_gcn import GCNTrainer
    from gist_b200.train_ist import add_self_loops
    ds = synth.make('cora', seed=0)
    dev = torch.device('cuda')
    src, dst = add_self_loops(ds.src, ds.dst, ds.num_nodes)
    g = GistGraph.from_edges(src, dst, ds.num_nodes, device=dev)
    args = SimpleNamespace(n_hidden=16, n_layers=1, dropout=0.0, lr=1e-2, weight_decay=5e-4, n_epochs=8,
                           use_layernorm='True', lr_scheduler=True)
    out = []
    for use_graph in (False, True):
        torch.manual_seed(1)
        tr = GCNTrainer(g, ds.feat.to(dev), ds.label.to(dev), ds.train_mask.to(dev), ds.num_classes, args, dev,
                        use_graph=use_graph)
        losses = [float(tr.train_epoch(e)) for e in range(8)]      # lr drops at epochs 4 and 6: two re-captures
        out.append((losses, [p.detach().clone() for p in tr.model.parameters()]))
    for a, b in zip(out[0][0], out[1][0]):
        assert abs(a - b) <= 1e-5 * max(abs(a), 1.0)
    for p, q in zip(out[0][1], out[1][1]):
        assert _rel(q, p) < 1e-5
