"""Data ingestion: the vectorised loaders of gist_b200/datasets.py against the
literal restatement of cluster_gcn/AmazonDataset.py::process (oracle/datasets_oracle.py) on small
files written in the reference's on-disk formats; CPU only."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sp
import torch


def _write_amazon(d, n, m, seed, list_labels, str_ids, name='amazon2M'):
    rng = np.random.RandomState(seed)
    ids = ['n%d' % i for i in range(n)] if str_ids else list(range(n))
    perm = rng.permutation(n)                       # id_map is NOT the identity
    nodes = []
    for i in range(n):
        r = rng.rand()
        nodes.append({'id': ids[i], 'val': bool(r < 0.15), 'test': bool(0.15 <= r < 0.35)})
    links = []
    seen = set()
    for _ in range(m):
        u, v = int(rng.randint(n)), int(rng.randint(n))
        if (u, v) in seen or (v, u) in seen:        # an undirected simple graph + a few self loops
            continue
        seen.add((u, v))
        links.append({'source': ids[u], 'target': ids[v]})
    G = {'directed': False, 'multigraph': False, 'graph': {}, 'nodes': nodes, 'links': links}
    # a node that is in the graph but missing from id_map: its edges are dropped (:50-52)
    id_map = {str(ids[i]): int(perm[i]) for i in range(n)}
    C = 5
    if list_labels:
        class_map = {str(ids[i]): [int(x) for x in (rng.rand(C) < 0.4)] for i in range(n)}
    else:
        class_map = {str(ids[i]): int(rng.randint(C)) for i in range(n)}
    feats = (rng.randn(n, 7) * rng.rand(7) * 10 + rng.randn(7)).astype(np.float32)
    feats[:, 3] = 2.5                                # a constant column: StandardScaler leaves scale 1
    np.save(os.path.join(d, '%s-feats.npy' % name), feats)
    for suffix, obj in (('G', G), ('id_map', id_map), ('class_map', class_map)):
        with open(os.path.join(d, '%s-%s.json' % (name, suffix)), 'w') as f:
            json.dump(obj, f)


def _canon(src, dst, n):
    key = np.asarray(dst, dtype=np.int64) * n + np.asarray(src, dtype=np.int64)
    return np.sort(key)


@pytest.mark.parametrize('list_labels,str_ids', [(False, False), (True, False), (False, True)])
def test_amazon2m_loader_matches_reference_restatement(tmp_path, list_labels, str_ids):
    from gist_b200 import datasets
    from oracle import datasets_oracle as O
    d = str(tmp_path)
    _write_amazon(d, n=300, m=2500, seed=3 + list_labels + 2 * str_ids, list_labels=list_labels, str_ids=str_ids)
    ref = O.amazon_process(d)
    got = datasets.load_amazon2m(d)
    g = got.g
    assert got.num_classes == ref['num_classes']
    assert g.number_of_nodes() == ref['n']
    assert g.number_of_edges() == len(ref['src'])
    # structure: bit-exact in canonical form (in-CSR, columns ascending within a row)
    rowptr, col = g.rowptr.numpy().astype(np.int64), g.col_buffer.numpy().astype(np.int64)
    dst = np.repeat(np.arange(ref['n']), np.diff(rowptr))
    assert np.array_equal(_canon(col, dst, ref['n']), _canon(ref['src'], ref['dst'], ref['n']))
    assert g.is_symmetric()
    for k in ('train_mask', 'val_mask', 'test_mask'):
        assert np.array_equal(g.ndata[k].numpy(), ref[k]), k
    assert np.array_equal(g.ndata['label'].numpy(), ref['labels'])
    assert g.ndata['label'].dtype == torch.int64 and g.ndata['feat'].dtype == torch.float32
    # features: sklearn's StandardScaler on float32 input vs float64 moments here
    np.testing.assert_allclose(g.ndata['feat'].numpy(), ref['feats'], rtol=2e-5, atol=2e-5)
    assert np.allclose(g.ndata['feat'].numpy()[:, 3], 0.0)          # constant column: centred, not scaled


def test_standardize_features_is_sklearn_standard_scaler():
    import sklearn.preprocessing
    from gist_b200 import datasets
    rng = np.random.RandomState(0)
    x = (rng.randn(500, 9) * 3 + 1).astype(np.float32)
    x[:, 4] = -1.0
    mask = rng.rand(500) < 0.6
    sc = sklearn.preprocessing.StandardScaler().fit(x[mask])         # …distrib.py:493-499
    got, ref = datasets.standardize_features(x, mask), sc.transform(x)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-6)
    assert np.array_equal(got, ref)          # same float32 arithmetic as the installed sklearn, bit for bit


@pytest.mark.parametrize('self_loop', [False, True])
def test_reddit_raw_files(tmp_path, self_loop):
    from gist_b200 import datasets
    rng = np.random.RandomState(1)
    n = 200
    a = sp.random(n, n, density=0.05, random_state=rng, format='coo')
    a = ((a + a.T) > 0).astype(np.float32)
    if self_loop:
        a = a + sp.eye(n, dtype=np.float32)
    a = sp.coo_matrix(a)
    sp.save_npz(os.path.join(str(tmp_path), 'reddit_self_loop_graph.npz' if self_loop else 'reddit_graph.npz'), a)
    node_types = rng.randint(1, 4, size=n)
    feat = rng.randn(n, 6).astype(np.float32)
    label = rng.randint(0, 41, size=n)
    label[0] = 40
    np.savez(os.path.join(str(tmp_path), 'reddit_data.npz'), feature=feat, label=label, node_types=node_types)
    args = SimpleNamespace(dataset='reddit-self-loop' if self_loop else 'reddit')
    data = datasets.load_data(args, raw_dir=str(tmp_path))
    g = data.g
    assert data.num_classes == 41 and g.number_of_nodes() == n and g.number_of_edges() == a.nnz
    assert np.array_equal(g.in_degrees().numpy(), np.bincount(a.col, minlength=n))     # one edge row -> col per non-zero
    assert np.array_equal(g.ndata['train_mask'].numpy(), node_types == 1)
    assert np.array_equal(g.ndata['val_mask'].numpy(), node_types == 2)
    assert np.array_equal(g.ndata['test_mask'].numpy(), node_types == 3)
    assert np.array_equal(g.ndata['feat'].numpy(), feat) and np.array_equal(g.ndata['label'].numpy(), label)


def test_load_data_rejects_unknown_dataset():
    from gist_b200 import datasets
    with pytest.raises(ValueError):
        datasets.load_data(SimpleNamespace(dataset='ppi'))


def _write_planetoid(d, name, n, n_lab, n_test, F, C, seed, gaps):
    import pickle
    from collections import defaultdict
    rng = np.random.RandomState(seed)
    n_all = n - n_test - gaps                       # allx rows; test ids live in [n_all, n)
    feat = sp.random(n, F, density=0.2, random_state=rng, format='csr', dtype=np.float64)
    feat.data[:] = rng.randint(1, 4, size=feat.nnz)
    feat = feat.tolil()
    feat[5, :] = 0                                   # an empty row: stays zero after normalisation
    feat = feat.tocsr()
    lab = rng.randint(0, C, size=n)
    onehot = np.eye(C, dtype=np.int32)[lab]
    test_ids = np.sort(rng.choice(np.arange(n_all, n), size=n_test, replace=False))    # gaps: citeseer's isolated nodes
    graph = defaultdict(list)
    for _ in range(4 * n):
        u, v = int(rng.randint(n)), int(rng.randint(n))
        graph[u].append(v)                           # asymmetric listings and duplicates on purpose
        if rng.rand() < 0.5:
            graph[v].append(u)
    graph[7].append(7)                               # a self-loop
    order = rng.permutation(n_test)
    objs = {'x': feat[:n_lab], 'y': onehot[:n_lab], 'allx': feat[:n_all], 'ally': onehot[:n_all],
            'tx': feat[test_ids], 'ty': onehot[test_ids], 'graph': graph}
    for k, v in objs.items():
        with open(os.path.join(d, 'ind.%s.%s' % (name, k)), 'wb') as f:
            pickle.dump(v, f)
    with open(os.path.join(d, 'ind.%s.test.index' % name), 'w') as f:
        f.write('\n'.join(str(int(t)) for t in test_ids[order]) + '\n')


@pytest.mark.parametrize('name,gaps', [('cora', 0), ('pubmed', 0), ('citeseer', 6)])
def test_citation_loader_matches_restatement(tmp_path, name, gaps):
    from gist_b200 import datasets
    from oracle import datasets_oracle as O
    d = str(tmp_path)
    _write_planetoid(d, name, n=900, n_lab=60, n_test=150, F=40, C=6, seed=len(name), gaps=gaps)
    ref = O.citation_load(d, name)
    got = datasets.load_data(SimpleNamespace(dataset=name), raw_dir=d)
    assert got.num_nodes == ref['n'] and got.num_labels == ref['num_labels']
    assert np.array_equal(_canon(got.src.numpy(), got.dst.numpy(), ref['n']), _canon(ref['src'], ref['dst'], ref['n']))
    for k in ('train_mask', 'val_mask', 'test_mask'):
        assert np.array_equal(getattr(got, k), ref[k]), k
    assert got.train_mask.sum() == 60 and got.val_mask.sum() == 500 and got.test_mask.sum() == 150
    assert np.array_equal(got.labels, ref['labels'])
    np.testing.assert_allclose(got.features, ref['features'], rtol=1e-6, atol=0)
    rs = got.features.sum(1)
    assert np.allclose(rs[rs > 0], 1.0, atol=1e-5) and got.features.dtype == np.float32
