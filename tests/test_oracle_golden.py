"""CPU: the functional oracle against the golden vectors produced by running the
reference's own Python files (oracle/gen_golden.py)."""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gist_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def close(a, b, tol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    assert np.abs(a - b).max() <= tol * scale if b.size else True


def T(x):
    return torch.from_numpy(np.asarray(x))


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_sage_gcn_logits_and_grads(ci):
    G = load('sage')
    p = 'sage%d_' % ci
    fin, hid, ncls, L, ln = G[p + 'cfg']
    g = O.OGraph(G[p + 'src'], G[p + 'dst'], int(G[p + 'n']))
    params = [(T(G[p + 'param.layers.%d.linear.weight' % l]).requires_grad_(True),
               T(G[p + 'param.layers.%d.linear.bias' % l]).requires_grad_(True)) for l in range(L + 1)]
    logits = O.sage_gcn_forward(g, T(G[p + 'x']), params, bool(ln))
    close(logits.detach(), G[p + 'logits'])
    loss = F.cross_entropy(logits, T(G[p + 'y']))
    close(loss.item(), G[p + 'loss'])
    loss.backward()
    for l, (w, b) in enumerate(params):
        close(w.grad, G[p + 'grad.layers.%d.linear.weight' % l], 1e-4)
        close(b.grad, G[p + 'grad.layers.%d.linear.bias' % l], 1e-4)


def test_single_layer_and_norm():
    G = load('sage')
    g = O.OGraph(G['layer_src'], G['layer_dst'], 40)
    assert np.array_equal(O.sage_norm(g).numpy(), G['layer_norm'])       # bit-exact 1/deg, inf->0
    out = O.ist_sage_layer(g, T(G['layer_x']), T(G['layer_w']), T(G['layer_b']), True, F.relu)
    close(out, G['layer_out'])


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_graphconv_gcn(ci):
    G = load('graphconv')
    p = 'gc%d_' % ci
    fin, hid, ncls, L, si, so, k = G[p + 'cfg']
    g = O.OGraph(G[p + 'src'], G[p + 'dst'], int(G[p + 'n']))
    params = [(T(G[p + 'param.layers.%d.weight' % l]).requires_grad_(True),
               T(G[p + 'param.layers.%d.bias' % l]).requires_grad_(True)) for l in range(L + 1)]
    logits = O.graphconv_gcn_forward(g, T(G[p + 'x']), params, True)
    close(logits.detach(), G[p + 'logits'])
    F.cross_entropy(logits, T(G[p + 'y'])).backward()
    for l, (w, b) in enumerate(params):
        close(w.grad, G[p + 'grad.layers.%d.weight' % l], 1e-4)
        close(b.grad, G[p + 'grad.layers.%d.bias' % l], 1e-4)


def test_baseline_gcn():
    G = load('graphconv')
    g = O.OGraph(G['base_src'], G['base_dst'], 50)
    params = [(T(G['base_param.layers.%d.weight' % l]), T(G['base_param.layers.%d.bias' % l]))
              for l in range(3)]
    close(O.graphconv_gcn_forward(g, T(G['base_x']), params, True), G['base_logits'])


@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_graphsage_layer(ci):
    """GraphSAGELayer of the reference (affine LN, optional bias, use_pp) -> oracle restatement."""
    G = load('graphsage')
    p = 'gl%d_' % ci
    fin, fout, has_bias, pp, ln, train, act = (int(v) for v in G[p + 'cfg'])
    g = O.OGraph(G[p + 'src'], G[p + 'dst'], 48)
    w = T(G[p + 'param.linear.weight']).requires_grad_(True)
    b = T(G[p + 'param.linear.bias']).requires_grad_(True) if has_bias else None
    lw = T(G[p + 'param.lynorm.weight']).requires_grad_(True) if ln else None
    lb = T(G[p + 'param.lynorm.bias']).requires_grad_(True) if ln else None
    x = T(G[p + 'x']).requires_grad_(True)
    y = O.graphsage_layer(g, x, w, b, lw, lb, bool(ln), F.relu if act else None, aggregate=not (pp and train))
    close(y.detach(), G[p + 'out'])
    (y * T(G[p + 'wy'])).sum().backward()
    close(x.grad, G[p + 'dx'], 1e-4)
    close(w.grad, G[p + 'grad.linear.weight'], 1e-4)
    if has_bias:
        close(b.grad, G[p + 'grad.linear.bias'], 1e-4)
    if ln:
        close(lw.grad, G[p + 'grad.lynorm.weight'], 1e-4)
        close(lb.grad, G[p + 'grad.lynorm.bias'], 1e-4)


def graphsage_params(G, p, L, grad=True):
    out = []
    for l in range(L + 1):
        pre = p + 'param.layers.%d.' % l
        t = [T(G[pre + 'linear.weight']), T(G[pre + 'linear.bias']),
             T(G[pre + 'lynorm.weight']) if l < L else None, T(G[pre + 'lynorm.bias']) if l < L else None]
        out.append(tuple(v.requires_grad_(True) if (v is not None and grad) else v for v in t))
    return out


@pytest.mark.parametrize('ci', [0, 1])
def test_graphsage_container(ci):
    G = load('graphsage')
    p = 'gs%d_' % ci
    fin, hid, ncls, L, pp = (int(v) for v in G[p + 'cfg'])
    g = O.OGraph(G[p + 'src'], G[p + 'dst'], int(G[p + 'n']))
    params = graphsage_params(G, p, L)
    logits = O.graphsage_forward(g, T(G[p + 'x']), params)
    close(logits.detach(), G[p + 'logits'])
    F.cross_entropy(logits, T(G[p + 'y'])).backward()
    names = ('linear.weight', 'linear.bias', 'lynorm.weight', 'lynorm.bias')
    for l, tup in enumerate(params):
        for nm, t in zip(names, tup):
            if t is not None:
                close(t.grad, G[p + 'grad.layers.%d.%s' % (l, nm)], 1e-4)


@pytest.mark.parametrize('ci', [0, 1])
def test_baseline_gcn_with_grads(ci):
    G = load('graphsage')
    p = 'bg%d_' % ci
    fin, hid, ncls, L, ln = (int(v) for v in G[p + 'cfg'])
    g = O.OGraph(G[p + 'src'], G[p + 'dst'], int(G[p + 'n']))
    params = [(T(G[p + 'param.layers.%d.weight' % l]).requires_grad_(True),
               T(G[p + 'param.layers.%d.bias' % l]).requires_grad_(True)) for l in range(L + 1)]
    logits = O.graphconv_gcn_forward(g, T(G[p + 'x']), params, bool(ln))
    close(logits.detach(), G[p + 'logits'])
    F.cross_entropy(logits, T(G[p + 'y'])).backward()
    for l, (w, b) in enumerate(params):
        close(w.grad, G[p + 'grad.layers.%d.weight' % l], 1e-4)
        close(b.grad, G[p + 'grad.layers.%d.bias' % l], 1e-4)


def test_create_partition_bit_exact():
    G = load('partition')
    for t, (seed, m, size) in enumerate(G['cp_triples']):
        random.seed(int(seed))
        for tag in 'ab':
            part = O.create_partition(int(m), int(size))
            assert np.array_equal(np.stack([p[0].numpy() for p in part]), G['cp%d%s_idx' % (t, tag)])
            assert np.array_equal(np.stack([p[1].numpy() for p in part]), G['cp%d%s_full' % (t, tag)])


def test_cluster_iter_batches_bit_exact():
    G = load('cluster_iter')
    psize, bs, seed = (int(v) for v in G['ci_cfg'])
    g = O.OGraph(G['ci_src'], G['ci_dst'], int(G['ci_n'])).subgraph(G['ci_train_nid'])
    part = G['ci_part'][G['ci_train_nid']]
    par_li = [np.nonzero(part == p)[0].astype(np.int64) for p in range(psize)]
    random.seed(seed)
    random.shuffle(par_li)                          # sampler.py:55
    k = 0
    for epoch in range(2):
        for i in range(psize // bs):
            nid = O.batch_node_ids(par_li, i, psize, bs)
            assert np.array_equal(nid, G['ci_b%d_nid' % k])
            sg = g.subgraph(nid)
            key = np.sort(sg.dst.numpy() * len(nid) + sg.src.numpy())
            assert np.array_equal(key, G['ci_b%d_edges' % k])
            assert np.array_equal(G['ci_feat'][G['ci_train_nid']][nid][:, 0], G['ci_b%d_feat0' % k])
            k += 1
        random.shuffle(par_li)                      # sampler.py:92
    assert k == int(G['ci_nbatches'])
    assert random.random() == float(G['ci_next_random'])      # same number of RNG draws


@pytest.mark.parametrize('ci', [0, 1])
def test_dispatch_sync_algebra_vs_reference_processes(ci):
    """Golden = the reference's DistributedGNNWrapper run on m gloo processes."""
    G = load('wrapper')
    p = 'w%d_' % ci
    m, fin, hid, ncls, L, seed = (int(v) for v in G[p + 'cfg'])
    base0 = [(T(G[p + 'r0_base0.layers.%d.linear.weight' % l]), T(G[p + 'r0_base0.layers.%d.linear.bias' % l]))
             for l in range(L + 1)]
    # the partition every rank drew is identical and equals create_partition on the same stream
    random.seed(seed)
    parts0 = [O.create_partition(m, hid) for _ in range(L)]
    for r in range(m):
        for l in range(L):
            assert np.array_equal(np.stack([q[0].numpy() for q in parts0[l]]), G[p + 'r%d_part0.%d' % (r, l)])
    subs = []
    for r in range(m):
        sub = O.sage_dispatch(base0, parts0, r)
        for l, (w, b) in enumerate(sub):
            if r == 0 or l < L:
                assert np.array_equal(w.numpy(), G[p + 'r%d_sub0.layers.%d.linear.weight' % (r, l)])
            assert np.array_equal(b.numpy(), G[p + 'r%d_sub0.layers.%d.linear.bias' % (r, l)])
        subs.append([(T(G[p + 'r%d_trained.layers.%d.linear.weight' % (r, l)]),
                      T(G[p + 'r%d_trained.layers.%d.linear.bias' % (r, l)])) for l in range(L + 1)])
    base1 = O.sage_sync(base0, parts0, subs)
    for l, (w, b) in enumerate(base1):
        assert np.array_equal(w.numpy(), G[p + 'r0_base1.layers.%d.linear.weight' % l])
        close(b, G[p + 'r0_base1.layers.%d.linear.bias' % l], 1e-6)
    for r in range(m):
        close(base1[L][1], G[p + 'r%d_lastbias_after_sync' % r], 1e-6)
    parts1 = [O.create_partition(m, hid) for _ in range(L)]
    base1g = [(T(G[p + 'r0_base1.layers.%d.linear.weight' % l]), T(G[p + 'r0_base1.layers.%d.linear.bias' % l]))
              for l in range(L + 1)]
    for r in range(m):
        for l in range(L):
            assert np.array_equal(np.stack([q[0].numpy() for q in parts1[l]]), G[p + 'r%d_part1.%d' % (r, l)])
        sub = O.sage_dispatch(base1g, parts1, r)
        for l, (w, b) in enumerate(sub[:L]):
            assert np.array_equal(w.numpy(), G[p + 'r%d_sub1.layers.%d.linear.weight' % (r, l)])
            assert np.array_equal(b.numpy(), G[p + 'r%d_sub1.layers.%d.linear.bias' % (r, l)])
        assert np.array_equal(sub[L][0].numpy(), G[p + 'r%d_sub1.layers.%d.linear.weight' % (r, L)])


@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_train_ist_split_merge(ci):
    """Golden = what gcn/train_ist.py main() split and merged on a tiny dataset."""
    G = load('train_ist')
    p = 'ti%d_' % ci
    si, so, L, m, hid, fin, ncls = (int(v) for v in G[p + 'cfg'])
    keys = ['layers.%d.%s' % (l, t) for l in range(L + 1) for t in ('weight', 'bias')]
    perms = [T(G[p + 'perm%d' % i]) for i in range(int(G[p + 'nperm']))]
    per_round = len(perms) // int(G[p + 'nrounds'])
    for r in range(int(G[p + 'nrounds'])):
        it = iter(perms[r * per_round:(r + 1) * per_round])
        feats_idx = [torch.chunk(next(it), m) if si else None]
        for _ in range(1, L):
            feats_idx.append(torch.chunk(next(it), m))
        feats_idx.append(torch.chunk(next(it), m) if so else None)
        main = {k: T(G[p + 'r%d_main.%s' % (r, k)]) for k in keys}
        for s in range(m):
            sub = O.graphconv_split(main, feats_idx, s, L, bool(si), bool(so))
            for k in keys:
                assert np.array_equal(sub[k].numpy(), G[p + 'split%d.%s' % (r * m + s, k)]), (r, s, k)
        trained = [{k: T(G[p + 'r%d_trained%d.%s' % (r, s, k)]) for k in keys} for s in range(m)]
        merged = O.graphconv_merge(main, feats_idx, trained, L, bool(si), bool(so))
        for k in keys:
            close(merged[k], G[p + 'r%d_merged.%s' % (r, k)], 1e-6)


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_gat_layer_oracle_vs_reference_golden(ci):
    """oracle gat_layer == the reference's own GATLayer (modules.py:24-65) run on the restated
    DGL UDF path: output and all gradients."""
    G = load('gat')
    p = 'gat%d_' % ci
    g = O.OGraph(G[p + 'src'], G[p + 'dst'], int(G[p + 'n']))
    x = T(G[p + 'x']).requires_grad_(True)
    fc = T(G[p + 'fc']).requires_grad_(True)
    attn = T(G[p + 'attn']).requires_grad_(True)
    out = O.gat_layer(g, x, fc, attn)
    close(out.detach(), G[p + 'out'])
    (out * T(G[p + 'wy'])).sum().backward()
    close(x.grad, G[p + 'dx'], 1e-4)
    close(fc.grad, G[p + 'dfc'], 1e-4)
    close(attn.grad, G[p + 'dattn'], 1e-4)


def test_gat_multi_head_oracle_vs_golden():
    G = load('gat')
    g = O.OGraph(G['mh_src'], G['mh_dst'], int(G['mh_n']))
    layers = [[(T(G['mh_fc_0_%d' % h]), T(G['mh_attn_0_%d' % h])) for h in range(2)],
              [(T(G['mh_fc_1_0']), T(G['mh_attn_1_0']))]]
    close(O.gat_forward(g, T(G['mh_x']), layers), G['mh_out'])
