"""K3 parity: device cluster-batch builder vs oracle induced subgraph (bit-exact in
canonical form), ndata row gathers, scan, K5 slice gather/scatter."""
import os

import numpy as np
import pytest
import torch

from oracle import gist_oracle as O
from tests.util import canonical_csr_from_gist, ograph, random_graph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('symmetric', [False, True])
@pytest.mark.parametrize('n_b', [1, 37, 500])
def test_subgraph_bit_exact(symmetric, n_b):
    from gist_b200 import GistGraph
    n, nnz = 2000, 60000
    src, dst = random_graph(n, nnz, seed=n_b)
    if symmetric:
        src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    g = GistGraph.from_edges(src, dst, n, device='cuda')
    g.ndata['feat'] = torch.randn(n, 602, device='cuda')
    g.ndata['label'] = torch.randint(0, 41, (n,), device='cuda')
    g.ndata['train_mask'] = torch.rand(n, device='cuda') < 0.5
    rng = np.random.RandomState(0)
    nids = rng.permutation(n)[:n_b].astype(np.int64)          # arbitrary order, unique
    sg = g.subgraph(nids)
    osg = ograph(src, dst, n).subgraph(nids)
    rp, col = canonical_csr_from_gist(sg)
    orp, ocol = osg.canonical_csr()
    assert torch.equal(rp, orp) and torch.equal(col, ocol)
    assert sg.number_of_edges() == osg.src.shape[0]
    assert sg.is_symmetric() == symmetric or osg.src.shape[0] == 0
    # csc of the batch == canonical csr of the reversed oracle graph
    colptr, row = sg.csc()
    rev = O.OGraph(osg.dst, osg.src, osg.n).canonical_csr()
    nn = sg.number_of_nodes()
    r_of_e = torch.repeat_interleave(torch.arange(nn), colptr.cpu().long()[1:] - colptr.cpu().long()[:-1])
    key, _ = torch.sort(r_of_e * nn + row.cpu().long()[:r_of_e.shape[0]])
    assert torch.equal(colptr.cpu().long(), rev[0]) and torch.equal(key % nn, rev[1])
    # ndata gathered exactly, node order preserved
    t = torch.from_numpy(nids)
    assert torch.equal(sg.ndata['feat'].cpu(), g.ndata['feat'].cpu()[t])
    assert torch.equal(sg.ndata['label'].cpu(), g.ndata['label'].cpu()[t])
    assert torch.equal(sg.ndata['train_mask'].cpu(), g.ndata['train_mask'].cpu()[t])
    assert torch.equal(sg.ndata['_ID'].cpu(), t)
    assert torch.equal(sg.inv_in_degree().cpu(), O.sage_norm(osg).reshape(-1))
    # scratch restored: a second build gives the same answer
    sg2 = g.subgraph(nids)
    assert torch.equal(sg2.rowptr, sg.rowptr) and torch.equal(sg2.col, sg.col)


def test_subgraph_empty():
    from gist_b200 import GistGraph
    src, dst = random_graph(100, 500, seed=0)
    g = GistGraph.from_edges(src, dst, 100, device='cuda')
    sg = g.subgraph(np.zeros(0, dtype=np.int64))
    assert sg.number_of_nodes() == 0 and sg.number_of_edges() == 0


@pytest.mark.parametrize('n', [0, 1, 31, 2048, 2049, 8192, 8193, 100000, 1234567])
def test_exclusive_scan(n):
    from gist_b200 import ops
    x = torch.randint(0, 50, (n,), dtype=torch.int32)
    got = ops.exclusive_scan(x.cuda()).cpu()
    ref = torch.zeros(n + 1, dtype=torch.int64)
    ref[1:] = torch.cumsum(x.long(), 0)
    assert torch.equal(got.long(), ref)


def test_cluster_iter_matches_oracle_batches():
    import random
    from gist_b200 import ClusterIter, synth
    ds = synth.make('reddit', seed=0, device='cpu', scale=0.01)
    g = synth.to_gist_graph(ds, device='cuda')
    train_nid = np.nonzero(ds.train_mask.numpy())[0].astype(np.int64)
    psize, bs = ds.part.max().item() + 1, 3
    for mode in ('step', 'epoch'):
        random.seed(3)
        it = ClusterIter('', g, psize, bs, train_nid, use_pp=False, h2d=mode)
        # oracle: same training graph, same partition, same python-random stream
        og = ograph(ds.src, ds.dst, ds.num_nodes).subgraph(train_nid)
        part = ds.part[torch.from_numpy(train_nid)].numpy()
        order = np.argsort(part, kind='stable')
        bounds = np.searchsorted(part[order], np.arange(psize + 1))
        par_li = [order[bounds[p]:bounds[p + 1]].astype(np.int64) for p in range(psize)]
        rng = random.Random(3)          # private copy of the stream the sampler draws from
        rng.shuffle(par_li)
        for epoch in range(2):
            for i, batch in enumerate(it):
                nids = O.batch_node_ids(par_li, i, psize, bs)
                assert torch.equal(batch.ndata['_ID'].cpu(), torch.from_numpy(nids))
                osg = og.subgraph(nids)
                rp, col = canonical_csr_from_gist(batch)
                orp, ocol = osg.canonical_csr()
                assert torch.equal(rp, orp) and torch.equal(col, ocol)
            assert i == len(it) - 1
            rng.shuffle(par_li)        # sampler.py:92


@pytest.mark.parametrize('shape', [(64, 48), (300, 1204), (41, 512)])
def test_slice_gather_scatter(shape):
    from gist_b200 import ops
    R, C = shape
    torch.manual_seed(0)
    W = torch.randn(R, C)
    ridx = torch.randperm(R)[:R // 2]
    cidx = torch.randperm(C)[:C // 3]
    Wd = W.cuda()
    for r, c in [(ridx, cidx), (ridx, None), (None, cidx), (None, None)]:
        ref = W
        if r is not None:
            ref = ref[r, :]
        if c is not None:
            ref = ref[:, c]
        got = ops.slice_gather(Wd, None if r is None else r.cuda(), None if c is None else c.cuda())
        assert torch.equal(got.cpu(), ref)
        new = torch.randn_like(ref)
        D = W.clone()
        if r is not None and c is not None:
            tmp = D[:, c]
            tmp[r, :] = new
            D[:, c] = tmp
        elif r is not None:
            D[r, :] = new
        elif c is not None:
            D[:, c] = new
        else:
            D = new.clone()
        Dd = W.clone().cuda()
        ops.slice_scatter_(Dd, new.cuda(), None if r is None else r.cuda(), None if c is None else c.cuda())
        assert torch.equal(Dd.cpu(), D)
    b = torch.randn(R)
    assert torch.equal(ops.slice_gather(b.cuda(), None, ridx.cuda()).cpu(), b[ridx])


def test_subgraph_hub_rows_and_padding_sentinel():
    """Parent rows far longer than the CTA-cooperative threshold (1024), and nids padded
    with -1 (isolated, zero-feature rows) as used for fixed-shape replay."""
    from gist_b200 import GistGraph
    from tests.util import powerlaw_graph
    n = 6000
    src, dst = powerlaw_graph(n, 60, seed=2)
    g = GistGraph.from_edges(src, dst, n, device='cuda')
    assert int(g.in_degrees().max()) > 5000
    g.ndata['feat'] = torch.randn(n, 10, device='cuda')
    g.ndata['label'] = torch.randint(0, 7, (n,), device='cuda')
    hubs = torch.argsort(g.in_degrees(), descending=True)[:40].cpu().numpy()
    rng = np.random.RandomState(1)
    rest = np.setdiff1d(rng.permutation(n)[:800], hubs)
    nids = np.concatenate([rest[:300], hubs, rest[300:]]).astype(np.int64)
    sg = g.subgraph(nids)
    osg = ograph(src, dst, n).subgraph(nids)
    rp, col = canonical_csr_from_gist(sg)
    orp, ocol = osg.canonical_csr()
    assert torch.equal(rp, orp) and torch.equal(col, ocol)
    # padded: same structure on the real rows, pad rows isolated with zero ndata
    pad = 37
    nids_p = np.concatenate([nids, -np.ones(pad, dtype=np.int64)])
    cap = int(g.in_degrees()[torch.from_numpy(nids).cuda()].sum().item())
    sgp = g.subgraph(nids_p, col_capacity=cap)
    rpp, colp = canonical_csr_from_gist(sgp)
    assert torch.equal(rpp[:len(nids) + 1], orp) and (rpp[len(nids):] == orp[-1]).all()
    assert torch.equal(colp, ocol)
    assert (sgp.ndata['feat'][len(nids):] == 0).all() and (sgp.ndata['label'][len(nids):] == 0).all()
    assert torch.equal(sgp.ndata['feat'][:len(nids)], sg.ndata['feat'])
    assert (sgp.inv_in_degree()[len(nids):] == 0).all()
    # scratch map fully restored
    assert (g._node_map() == -1).all()


def test_cluster_iter_metis_and_partition_cache(tmp_path):
    """No '_part' on the graph -> get_partition_list runs METIS (partition_utils.py:11-18); with a
    dataset name the partition is cached in the reference's file format and format-compatible
    files are replayed (sampler.py:44-51)."""
    import random
    from gist_b200 import ClusterIter, partition, synth
    ds = synth.make('reddit', seed=0, device='cpu', scale=0.01)
    g = synth.to_gist_graph(ds, device='cuda')
    del g.ndata['_part']
    train_nid = np.nonzero(ds.train_mask.numpy())[0].astype(np.int64)
    psize, bs = 12, 3
    random.seed(5)
    it = ClusterIter('synth-reddit', g, psize, bs, train_nid, use_pp=False, cache_dir=str(tmp_path))
    fn = partition.cache_path('synth-reddit', psize, str(tmp_path))
    assert os.path.exists(fn)
    cached = partition.load_partition(fn)
    n_train = it.g.number_of_nodes()
    allids = np.sort(np.concatenate(cached))
    assert np.array_equal(allids, np.arange(n_train))                   # a partition of the training graph
    assert all(np.all(np.diff(p) > 0) for p in cached)                  # ascending inside a part
    sizes = np.array([len(p) for p in cached])
    assert sizes.max() <= 1.05 * n_train / psize + 1
    # second construction replays the cached file instead of partitioning again: same batches
    random.seed(5)
    it2 = ClusterIter('synth-reddit', g, psize, bs, train_nid, use_pp=False, cache_dir=str(tmp_path))
    for b1, b2 in zip(it, it2):
        assert torch.equal(b1.ndata['_ID'], b2.ndata['_ID'])
        assert torch.equal(b1.rowptr, b2.rowptr) and torch.equal(b1.col, b2.col)
    # METIS cuts no more edges than the id-order split of the same sizes (this scaled-down graph is
    # nearly dense, so there is little to gain; tests/test_partition.py checks quality on a sparse one)
    rp = it.g.rowptr.cpu().numpy().astype(np.int64)
    col = it.g.col.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(n_train), np.diff(rp))
    part = np.empty(n_train, dtype=np.int64)
    for k, p in enumerate(cached):
        part[p] = k
    naive = np.arange(n_train) * psize // n_train
    assert (part[rows] != part[col]).sum() <= (naive[rows] != naive[col]).sum()
