"""Data ingestion either side of the hot path: the on-disk formats the
reference trains from, read into the graph object the kernels use.

* Amazon2M — the GraphSAGE / Cluster-GCN file set ``{name}-feats.npy``, ``{name}-G.json``,
  ``{name}-id_map.json``, ``{name}-class_map.json`` exactly as ``AmazonDataset.process`` reads it
  (cluster_gcn/AmazonDataset.py:25-118), without TensorFlow's gfile, networkx or DGL: the JSON is
  parsed directly and the edge / mask / label loops are vectorised.
* Reddit — the raw files of ``dgl.data.RedditDataset`` the reference loads through
  ``dgl.data.load_data`` (cluster_gcn/utils.py:7, :110-ff): ``reddit_data.npz`` (feature, label,
  node_types 1/2/3 = train/val/test) and ``reddit_graph.npz`` / ``reddit_self_loop_graph.npz``
  (scipy ``save_npz``) [DGL-recall: file names and fields of DGL 0.5.x].
* Cora / Citeseer / PubMed — the Planetoid ``ind.{name}.{x,y,tx,ty,allx,ally,graph,test.index}``
  files behind ``dgl.data.load_data`` in gcn/train.py:33-41 and gcn/train_ist.py:64-92
  (``load_citation``) [DGL-recall: CitationGraphDataset of DGL 0.5.x = Kipf & Welling's loader].
* ``standardize_features`` — the StandardScaler step of get_data (…distrib.py:493-499).
* ``load_data`` / ``get_data`` — the two reference entry points (utils.py:82-ff,
  cluster_gcn_ist_distrib.py:484-518) on top of them.

Everything here is host-side preprocessing, as in the reference; the result is a ``GistGraph``
(int32 in-CSR) with the reference's ``ndata`` keys, ready for ``.to('cuda')`` / ``ClusterIter``.
"""
import json
import os
import pickle
from collections import namedtuple
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import torch

from .graph import GistGraph

Dataset = namedtuple('Dataset', ['num_classes', 'g'])


def _key(k, is_digit):
    return int(k) if is_digit else k


def standardize_features(feats, train_rows):
    """sklearn.preprocessing.StandardScaler fitted on the training rows and applied to all rows
    (AmazonDataset.py:87-90; cluster_gcn_ist_distrib.py:493-499): population variance, a
    zero-variance column is left unscaled.  Moments are accumulated in float64; the transform
    mirrors sklearn's arithmetic on float32 input: ``X -= mean_.astype(float32)`` then
    ``X /= scale_.astype(float32)`` in place on the float32 copy (sklearn >= 1.x casts the moments
    to the data type first; the reference pins only ``scikit-learn>=0.20.0``, and releases that
    subtract the float64 moments directly differ from this by at most 1 ulp).  Bit-identical to
    the installed StandardScaler — pinned by tests/test_datasets.py against the real sklearn."""
    feats = np.asarray(feats)
    tr = feats[train_rows].astype(np.float64)
    mean = tr.mean(axis=0)
    var = tr.var(axis=0)
    scale = np.sqrt(var)
    # sklearn's _handle_zeros_in_scale: (near-)constant columns get scale 1
    eps = 10 * np.finfo(scale.dtype).eps
    scale[scale < eps] = 1.0
    out = feats.astype(np.float32, copy=True)
    out -= mean.astype(np.float32)
    out /= scale.astype(np.float32)
    return out


def load_amazon2m(raw_dir, name='amazon2M', standardize=True):
    """AmazonDataset.process (AmazonDataset.py:25-118) -> GistGraph with ndata
    feat / label / train_mask / val_mask / test_mask, and the number of classes."""
    feats = np.load(os.path.join(raw_dir, '%s-feats.npy' % name)).astype(np.float32)      # :29
    with open(os.path.join(raw_dir, '%s-G.json' % name)) as f:
        G = json.load(f)                                                                    # :33
    with open(os.path.join(raw_dir, '%s-id_map.json' % name)) as f:
        id_map = json.load(f)                                                               # :37
    is_digit = list(id_map.keys())[0].isdigit()                                             # :38
    id_map = {_key(k, is_digit): int(v) for k, v in id_map.items()}
    with open(os.path.join(raw_dir, '%s-class_map.json' % name)) as f:
        class_map = json.load(f)                                                            # :41
    is_list = isinstance(list(class_map.values())[0], list)                                 # :42
    class_map = {_key(k, is_digit): (v if is_list else int(v)) for k, v in class_map.items()}

    n = len(id_map)                                                                         # :55
    nodes = G['nodes']
    node_ids = [nd['id'] for nd in nodes]
    links = G.get('links', G.get('edges'))
    # node_link_graph: source / target are node ids (networkx >= 2)
    src_ids = [e['source'] for e in links]
    dst_ids = [e['target'] for e in links]
    get = id_map.get
    s = np.fromiter((get(u, -1) for u in src_ids), dtype=np.int64, count=len(links))
    d = np.fromiter((get(v, -1) for v in dst_ids), dtype=np.int64, count=len(links))
    keep = (s >= 0) & (d >= 0)                                                              # :50-52
    s, d = s[keep], d[keep]

    pos = np.fromiter((get(i, -1) for i in node_ids), dtype=np.int64, count=len(nodes))
    if (pos < 0).any():                 # the reference indexes id_map[n] for every node of G (:58-59)
        raise KeyError(node_ids[int(np.argmax(pos < 0))])
    is_val = np.fromiter((bool(nd['val']) for nd in nodes), dtype=bool, count=len(nodes))
    is_test = np.fromiter((bool(nd['test']) for nd in nodes), dtype=bool, count=len(nodes))
    val_mask = np.zeros(n, dtype=bool)
    test_mask = np.zeros(n, dtype=bool)
    val_mask[pos[is_val]] = True                                                            # :58, :104
    test_mask[pos[is_test]] = True                                                          # :59, :105
    train_mask = ~(val_mask | test_mask)                                                    # :60-63, :103

    if is_list:                                                                             # :75-79
        num_classes = len(list(class_map.values())[0])
        lab = np.zeros((n, num_classes), dtype=np.float32)
        for k, v in class_map.items():
            lab[id_map[k], :] = np.array(v)
    else:                                                                                   # :80-84
        num_classes = len(set(class_map.values()))
        lab = np.zeros((n, num_classes), dtype=np.float32)
        ks = np.fromiter((id_map[k] for k in class_map), dtype=np.int64, count=len(class_map))
        vs = np.fromiter(class_map.values(), dtype=np.int64, count=len(class_map))
        lab[ks, vs] = 1
    labels = np.argmax(lab, 1)                                                              # :86

    if standardize:
        # the scaler is fitted on the rows of the nodes that are neither val nor test, in the
        # G.nodes() order (:88-92) — the moments do not depend on the order
        train_ids = pos[~is_val & ~is_test]
        feats = standardize_features(feats, train_ids)

    adj = sp.csr_matrix((np.ones(s.shape[0], dtype=np.float32), (s, d)), shape=(n, n))      # :94-97
    adj = adj + adj.transpose()
    g = GistGraph.from_scipy(adj)                                                           # :111 (pattern only)
    g.ndata['train_mask'] = torch.from_numpy(train_mask)
    g.ndata['val_mask'] = torch.from_numpy(val_mask)
    g.ndata['test_mask'] = torch.from_numpy(test_mask)
    g.ndata['feat'] = torch.tensor(feats, dtype=torch.float32)
    g.ndata['label'] = torch.tensor(labels, dtype=torch.int64)
    return Dataset(num_classes=num_classes, g=g)


def load_reddit(raw_dir, self_loop=False):
    """dgl.data.RedditDataset's raw files -> GistGraph with the same ndata keys [DGL-recall]."""
    data = np.load(os.path.join(raw_dir, 'reddit_data.npz'))
    adj = sp.load_npz(os.path.join(raw_dir, 'reddit_self_loop_graph.npz' if self_loop else 'reddit_graph.npz'))
    g = GistGraph.from_scipy(adj)
    node_types = data['node_types']
    g.ndata['train_mask'] = torch.from_numpy(node_types == 1)
    g.ndata['val_mask'] = torch.from_numpy(node_types == 2)
    g.ndata['test_mask'] = torch.from_numpy(node_types == 3)
    g.ndata['feat'] = torch.tensor(data['feature'], dtype=torch.float32)
    g.ndata['label'] = torch.tensor(data['label'], dtype=torch.int64)
    return Dataset(num_classes=int(data['label'].max()) + 1, g=g)


def load_citation(raw_dir, name):
    """The citation datasets as DGL 0.5's ``CitationGraphDataset`` hands them to gcn/train.py:33-41
    and gcn/train_ist.py:64-92: ``features`` (row-normalised, dense float32), ``labels`` (int64),
    ``train_mask`` / ``val_mask`` / ``test_mask``, ``num_labels``, and the graph as the directed edge
    list of ``nx.DiGraph(nx.from_dict_of_lists(graph))`` (``src`` / ``dst``; the trainers add the
    self-loops themselves, train.py:66-68).  [DGL-recall]"""
    def rd(suffix):
        with open(os.path.join(raw_dir, 'ind.%s.%s' % (name, suffix)), 'rb') as f:
            return pickle.load(f, encoding='latin1')
    x, y, tx, ty, allx, ally, graph = (rd(sfx) for sfx in ('x', 'y', 'tx', 'ty', 'allx', 'ally', 'graph'))
    with open(os.path.join(raw_dir, 'ind.%s.test.index' % name)) as f:
        test_idx_reorder = np.array([int(line.strip()) for line in f if line.strip()], dtype=np.int64)
    test_idx_range = np.sort(test_idx_reorder)
    if name == 'citeseer':
        # isolated test nodes are missing from tx / ty: zero rows at their positions
        full = np.arange(test_idx_range.min(), test_idx_range.max() + 1)
        tx_ext = sp.lil_matrix((len(full), x.shape[1]))
        tx_ext[test_idx_range - test_idx_range.min(), :] = tx
        tx = tx_ext
        ty_ext = np.zeros((len(full), y.shape[1]))
        ty_ext[test_idx_range - test_idx_range.min(), :] = ty
        ty = ty_ext
    features = sp.vstack((allx, tx)).tolil()
    features[test_idx_reorder, :] = features[test_idx_range, :]
    onehot = np.vstack((ally, ty))
    onehot[test_idx_reorder, :] = onehot[test_idx_range, :]
    labels = np.argmax(onehot, 1).astype(np.int64)
    n = features.shape[0]
    train_mask = np.zeros(n, dtype=bool)
    val_mask = np.zeros(n, dtype=bool)
    test_mask = np.zeros(n, dtype=bool)
    train_mask[:len(y)] = True
    val_mask[len(y):len(y) + 500] = True
    test_mask[test_idx_range] = True
    # _preprocess_features: rows scaled to sum 1, empty rows stay 0
    features = sp.csr_matrix(features, dtype=np.float64)
    rowsum = np.asarray(features.sum(1)).reshape(-1)
    with np.errstate(divide='ignore'):
        r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.0
    features = np.asarray(sp.diags(r_inv).dot(features).todense()).astype(np.float32)
    # nx.from_dict_of_lists -> undirected simple graph (a self-loop stays one edge); DiGraph: both directions
    keys = np.fromiter(graph.keys(), dtype=np.int64, count=len(graph))
    lens = np.fromiter((len(v) for v in graph.values()), dtype=np.int64, count=len(graph))
    u = np.repeat(keys, lens)
    v = np.fromiter((w for nb in graph.values() for w in nb), dtype=np.int64, count=int(lens.sum()))
    a = sp.coo_matrix((np.ones(len(u), dtype=np.float32), (u, v)), shape=(n, n)).tocsr()
    a = (a + a.transpose()).tocoo()
    return SimpleNamespace(src=torch.from_numpy(a.row.astype(np.int64)), dst=torch.from_numpy(a.col.astype(np.int64)),
                           features=features, labels=labels, train_mask=train_mask, val_mask=val_mask,
                           test_mask=test_mask, num_labels=int(onehot.shape[1]), num_nodes=n)


def load_data(args, raw_dir=None):
    """cluster_gcn/utils.py::load_data for the datasets of the hot path's configs."""
    name = args.dataset
    if name == 'amazon2m':
        return load_amazon2m(raw_dir or './amazon2m_data/amazon2M')       # AmazonDataset(save_dir=…) default
    if name.startswith('reddit'):
        return load_reddit(raw_dir or os.path.expanduser('~/.dgl/reddit'), self_loop='self-loop' in name)
    if name in ('cora', 'citeseer', 'pubmed'):                          # gcn/train.py:33-34
        return load_citation(raw_dir or os.path.expanduser('~/.dgl/%s' % name), name)
    raise ValueError('gist_b200.datasets.load_data: unknown dataset %r (amazon2m, reddit, reddit-self-loop, cora, citeseer, pubmed)' % name)


def get_data(args, device, raw_dir=None):
    """cluster_gcn_ist_distrib.py::get_data (:484-518): load, standardise on the training rows,
    build the cluster iterator over the training subgraph, move the graph to the device."""
    from .sampler import ClusterIter
    data = load_data(args, raw_dir)
    g = data.g
    train_mask, val_mask, test_mask = g.ndata['train_mask'], g.ndata['val_mask'], g.ndata['test_mask']
    labels = g.ndata['label']
    train_nid = np.nonzero(train_mask.numpy())[0].astype(np.int64)
    if getattr(args, 'normalize', False):
        g.ndata['feat'] = torch.from_numpy(standardize_features(g.ndata['feat'].numpy(), train_mask.numpy()))
    in_feats = g.ndata['feat'].shape[1]
    n_edges = g.number_of_edges()
    g = g.to(device)                                                        # resident in HBM for ClusterIter
    cluster_iterator = ClusterIter(args.dataset, g, args.psize, args.batch_size, train_nid,
                                   use_pp=getattr(args, 'use_pp', False))
    return (g, cluster_iterator, train_mask, val_mask.to(device), test_mask.to(device), labels, train_nid,
            in_feats, data.num_classes, n_edges)
