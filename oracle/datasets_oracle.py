"""TEST INFRASTRUCTURE ONLY — literal CPU restatement of the reference's Amazon2M ingestion.

Follows cluster_gcn/AmazonDataset.py::process line by line (the loops, networkx's
``node_link_graph``, sklearn's ``StandardScaler``, scipy's symmetrisation), with the two pieces
that are absent here replaced by what they do: ``tf.io.gfile.GFile`` -> ``open`` and
``dgl.convert.from_scipy`` -> the (row, col) pairs of the matrix's non-zeros [DGL-recall].
The product loader (gist_b200/datasets.py) is the vectorised form and is checked against this.
"""
import json

import numpy as np
import scipy.sparse as sp
import sklearn.preprocessing
from networkx.readwrite import json_graph


def sample_mask(idx, n):
    """dgl.data.utils-style mask (AmazonDataset.py:103-105)."""
    mask = np.zeros(n)
    mask[idx] = 1
    return mask.astype(bool)


def amazon_process(raw_path, name='amazon2M'):
    feats = np.load('{}/{}-feats.npy'.format(raw_path, name)).astype(np.float32)                       # :29
    G = json_graph.node_link_graph(json.load(open('{}/{}-G.json'.format(raw_path, name))), edges='links')   # :33
    id_map = json.load(open('{}/{}-id_map.json'.format(raw_path, name)))                               # :37
    is_digit = list(id_map.keys())[0].isdigit()
    id_map = {(int(k) if is_digit else k): int(v) for k, v in id_map.items()}
    class_map = json.load(open('{}/{}-class_map.json'.format(raw_path, name)))                         # :41
    is_instance = isinstance(list(class_map.values())[0], list)
    class_map = {(int(k) if is_digit else k): (v if is_instance else int(v)) for k, v in class_map.items()}

    edges = []                                                                                          # :48-52
    for edge in G.edges():
        if edge[0] in id_map and edge[1] in id_map:
            edges.append((id_map[edge[0]], id_map[edge[1]]))
    _nodes = len(id_map)                                                                                # :55
    val_nodes = np.array([id_map[n] for n in G.nodes() if G.nodes[n]['val']], dtype=np.int32)           # :58
    test_nodes = np.array([id_map[n] for n in G.nodes() if G.nodes[n]['test']], dtype=np.int32)         # :59
    is_train = np.ones((_nodes), dtype=bool)                                                            # :60 (np.bool)
    is_train[test_nodes] = False
    is_train[val_nodes] = False
    train_nodes = np.array([n for n in range(_nodes) if is_train[n]], dtype=np.int32)                   # :63
    _edges = np.array(edges, dtype=np.int32).reshape(-1, 2)                                             # :70

    if isinstance(list(class_map.values())[0], list):                                                   # :75-84
        num_classes = len(list(class_map.values())[0])
        _labels = np.zeros((_nodes, num_classes), dtype=np.float32)
        for k in class_map.keys():
            _labels[id_map[k], :] = np.array(class_map[k])
    else:
        num_classes = len(set(class_map.values()))
        _labels = np.zeros((_nodes, num_classes), dtype=np.float32)
        for k in class_map.keys():
            _labels[id_map[k], class_map[k]] = 1
    _labels = np.argmax(_labels, 1)                                                                     # :86

    train_ids = np.array([id_map[n] for n in G.nodes() if not G.nodes[n]['val'] and not G.nodes[n]['test']])  # :88
    train_feats = feats[train_ids]
    scaler = sklearn.preprocessing.StandardScaler()
    scaler.fit(train_feats)
    _feats = scaler.transform(feats)                                                                    # :92

    adj = sp.csr_matrix((np.ones((_edges.shape[0]), dtype=np.float32), (_edges[:, 0], _edges[:, 1])),
                        shape=(_nodes, _nodes))                                                         # :94-97
    adj += adj.transpose()
    coo = adj.tocoo()                                                                                   # from_scipy: one edge per non-zero
    return dict(src=coo.row.astype(np.int64), dst=coo.col.astype(np.int64), n=_nodes, feats=_feats.astype(np.float32),
                labels=_labels.astype(np.int64), train_mask=sample_mask(train_nodes, _nodes),
                val_mask=sample_mask(val_nodes, _nodes), test_mask=sample_mask(test_nodes, _nodes),
                num_classes=num_classes)


def citation_load(raw_path, name):
    """Kipf & Welling's ``load_data`` as DGL 0.5's CitationGraphDataset runs it [DGL-recall], with
    networkx building the graph exactly as gcn/train.py receives it (``data.graph``)."""
    import pickle
    import networkx as nx
    objects = []
    for suffix in ['x', 'y', 'tx', 'ty', 'allx', 'ally', 'graph']:
        with open('{}/ind.{}.{}'.format(raw_path, name, suffix), 'rb') as f:
            objects.append(pickle.load(f, encoding='latin1'))
    x, y, tx, ty, allx, ally, graph = tuple(objects)
    test_idx_reorder = [int(line.strip()) for line in open('{}/ind.{}.test.index'.format(raw_path, name))]
    test_idx_range = np.sort(test_idx_reorder)
    if name == 'citeseer':
        test_idx_range_full = range(min(test_idx_reorder), max(test_idx_reorder) + 1)
        tx_extended = sp.lil_matrix((len(test_idx_range_full), x.shape[1]))
        tx_extended[test_idx_range - min(test_idx_range), :] = tx
        tx = tx_extended
        ty_extended = np.zeros((len(test_idx_range_full), y.shape[1]))
        ty_extended[test_idx_range - min(test_idx_range), :] = ty
        ty = ty_extended
    features = sp.vstack((allx, tx)).tolil()
    features[test_idx_reorder, :] = features[test_idx_range, :]
    g = nx.DiGraph(nx.from_dict_of_lists(graph))
    onehot_labels = np.vstack((ally, ty))
    onehot_labels[test_idx_reorder, :] = onehot_labels[test_idx_range, :]
    labels = np.argmax(onehot_labels, 1)
    idx_test = test_idx_range.tolist()
    idx_train = range(len(y))
    idx_val = range(len(y), len(y) + 500)
    n = labels.shape[0]
    rowsum = np.asarray(features.sum(1))
    with np.errstate(divide='ignore'):
        r_inv = np.power(rowsum, -1.).flatten()
    r_inv[np.isinf(r_inv)] = 0.
    feats = np.asarray(sp.diags(r_inv).dot(features).todense())
    e = np.array(list(g.edges()), dtype=np.int64).reshape(-1, 2)
    return dict(src=e[:, 0], dst=e[:, 1], n=n, features=feats.astype(np.float32), labels=labels.astype(np.int64),
                train_mask=sample_mask(idx_train, n), val_mask=sample_mask(idx_val, n),
                test_mask=sample_mask(idx_test, n), num_labels=onehot_labels.shape[1])
