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


@pytest.mark.parametrize('case', ['random', 'hubs_padded', 'isolated', 'one_row', 'reddit_batch'])
def test_chunked_builder_gives_the_same_csr_as_the_row_builder(case, monkeypatch):
    """K3 v2 (one warp per 128-edge chunk of the parent rows) vs K3 v1 (one warp per row), array for array:
    rowptr, columns in parent order, 1 / degree; building into existing buffers (the pipelined trainer) and
    the scratch relabel map restored."""
    from gist_b200 import GistGraph, graph as gg, synth
    from tests.util import powerlaw_graph
    rng = np.random.RandomState(3)
    if case == 'reddit_batch':
        ds = synth.make('reddit', seed=0, device='cpu', scale=0.05)
        g = synth.to_gist_graph(ds, device='cuda')
        n = ds.num_nodes
        nids = rng.permutation(n)[:2100].astype(np.int64)
    else:
        n = 6000
        if case == 'hubs_padded':
            src, dst = powerlaw_graph(n, 60, seed=2)
        else:
            src, dst = random_graph(n, 90000, seed=7)
            src, dst = torch.cat([src, dst]), torch.cat([dst, src])
        g = GistGraph.from_edges(src, dst, n, device='cuda')
        if case == 'hubs_padded':
            hubs = torch.argsort(g.in_degrees(), descending=True)[:40].cpu().numpy()
            rest = np.setdiff1d(rng.permutation(n)[:900], hubs)
            nids = np.concatenate([rest[:300], hubs, -np.ones(5, dtype=np.int64), rest[300:], -np.ones(37, dtype=np.int64)])
        elif case == 'isolated':
            deg = g.in_degrees().cpu().numpy()
            nids = np.concatenate([np.nonzero(deg == 0)[0][:50], rng.permutation(n)[:3]]).astype(np.int64)
            nids = np.unique(nids)
        elif case == 'one_row':
            nids = np.array([int(torch.argmax(g.in_degrees()))], dtype=np.int64)
        else:
            nids = rng.permutation(n)[:1500].astype(np.int64)
    nids = nids.astype(np.int64)
    real = torch.from_numpy(nids[nids >= 0]).cuda()
    cap = int(g.in_degrees()[real].sum().item()) + 11
    monkeypatch.setattr(gg, 'BUILDER', 'rows')
    a = g.subgraph(nids, col_capacity=cap, walk_capacity=cap, gather_ndata=False)
    launches = __import__('gist_b200')._lib.launch_count()
    monkeypatch.setattr(gg, 'BUILDER', 'chunks')
    b = g.subgraph(nids, col_capacity=cap, walk_capacity=cap, gather_ndata=False)
    assert __import__('gist_b200')._lib.launch_count() - launches >= 6          # the v2 pipeline really ran
    nnz = a.number_of_edges()
    assert b.number_of_edges() == nnz
    assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.col_buffer[:nnz], b.col_buffer[:nnz])
    assert torch.equal(a.inv_in_degree(), b.inv_in_degree())
    assert (g._node_map() == -1).all()
    # rebuild a different batch into b's buffers, then this one again
    other = np.where(nids >= 0, (nids * 7 + 1) % n, -1)
    _, first = np.unique(other, return_index=True)
    keep = np.zeros(len(other), dtype=bool)
    keep[first] = True
    other = np.where(keep | (other < 0), other, -1).astype(np.int64)              # unique ids, same length
    capo = max(cap, int(g.in_degrees()[torch.from_numpy(other[other >= 0]).cuda()].sum().item()))
    b2 = g.subgraph(nids, col_capacity=capo, walk_capacity=capo, gather_ndata=False)
    g.subgraph(other, col_capacity=capo, walk_capacity=capo, gather_ndata=False, out=b2)
    monkeypatch.setattr(gg, 'BUILDER', 'rows')
    ao = g.subgraph(other, col_capacity=capo, walk_capacity=capo, gather_ndata=False)
    nz = ao.number_of_edges()
    assert torch.equal(ao.rowptr, b2.rowptr) and torch.equal(ao.col_buffer[:nz], b2.col_buffer[:nz])
    assert torch.equal(ao.inv_in_degree(), b2.inv_in_degree())
    assert (g._node_map() == -1).all()


@pytest.mark.parametrize('staged', [True, False])
@pytest.mark.parametrize('shape', [(64, 48, 96), (300, 1204, 1208), (41, 100, 104), (128, 8192, 8192), (37, 40000, 40000)])
def test_row_streaming_scatter_equals_indexing(shape, staged, monkeypatch):
    """gist_slice_scatter_rows_f32 (whole sectors of the owned rows, left to right) against torch indexing:
    several jobs with disjoint rows on one destination, ragged column counts, identity rows; source rows staged
    in shared memory (when two fit 72 KB) or read through L2."""
    from gist_b200 import ops
    if not staged:
        monkeypatch.setenv('GIST_MERGE_NO_STAGE', '1')
    R, C, ld = shape
    torch.manual_seed(R + C)
    buf = torch.randn(R, ld, device='cuda')
    dst = buf[:, :C]
    ref = dst.clone()
    perm = torch.randperm(R, device='cuda')
    jobs = []
    for s in range(3):
        ridx = perm[s * (R // 3):(s + 1) * (R // 3)].contiguous()
        cidx = torch.randperm(C, device='cuda')[:max(C // (s + 2), 1)].contiguous()
        src = torch.randn(ridx.numel(), cidx.numel(), device='cuda')
        assert ops.slice_scatter_rows_ok(dst, ridx, cidx)
        jobs.append((src, ridx, ops.index_invert(cidx, C), dst))
        ref[ridx.unsqueeze(1), cidx.unsqueeze(0)] = src
    pad_before = buf[:, C:].clone()
    ops.slice_scatter_rows_(jobs)
    assert torch.equal(dst, ref)
    assert torch.equal(buf[:, C:], pad_before)                  # row padding untouched
    # identity rows
    cidx = torch.randperm(C, device='cuda')[:C // 2].contiguous()
    src = torch.randn(R, cidx.numel(), device='cuda')
    ref[:, cidx] = src
    ops.slice_scatter_rows_([(src, None, ops.index_invert(cidx, C), dst)])
    assert torch.equal(dst, ref)
    inv = ops.index_invert(cidx, C).cpu()
    assert (inv >= 0).sum() == cidx.numel() and torch.equal(cidx.cpu()[inv[inv >= 0].long()], torch.nonzero(inv >= 0).reshape(-1))


@pytest.mark.parametrize('m', [2, 8])
def test_wrapper_merge_row_streaming_equals_scatter(m, monkeypatch):
    """sync_model's merge of all m sites' slices into the full-model replica: the row-streaming kernel on the
    wide layers (forced on at this small width) leaves the same replica, bit for bit, as the 4-byte scatters."""
    import random
    from types import SimpleNamespace
    from gist_b200.ist import DistributedGNNWrapper
    random.seed(1); torch.manual_seed(1); np.random.seed(1)
    args = SimpleNamespace(rank=0, num_subnet=m, n_hidden=512, n_layers=3, dropout=0.0, use_layernorm=True)
    w = DistributedGNNWrapper(args, None, 100, 41, torch.device('cuda', 0))
    parts = w._to_dev(w.sample_partitions())
    numel = sum(t.numel() for lyr in w.sub_model.layers for t in (lyr.linear.weight, lyr.linear.bias))
    gathered = torch.randn(m, numel, device='cuda')
    base0 = [p.detach().clone() for p in w.base_model.parameters()]
    out = {}
    for mode in ('scatter', 'rows'):
        monkeypatch.setenv('GIST_MERGE', mode)
        with torch.no_grad():
            for p, q in zip(w.base_model.parameters(), base0):
                p.copy_(q)
            from gist_b200 import _lib
            l0 = _lib.launch_count()
            w._merge(gathered, parts)
            out[mode] = ([p.detach().clone() for p in w.base_model.parameters()], _lib.launch_count() - l0)
    for a, b in zip(out['scatter'][0], out['rows'][0]):
        assert torch.equal(a, b)
    assert out['rows'][1] > out['scatter'][1]                   # the streaming launch (+ index inversions) really ran
    assert any(not torch.equal(a, b) for a, b in zip(out['rows'][0], base0))
