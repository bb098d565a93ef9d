"""Deterministic synthetic graphs of the shapes BASELINE.json names (no dataset
files exist here, no network).  Degree-corrected stochastic block model:
power-law node weights, `n_blocks` planted communities scattered over the id
space (ids are NOT community-ordered, like real Reddit ids), symmetric, no self
loops, no multi-edges.  The planted block of a node doubles as its METIS-style
part (`ndata['_part']`) — METIS itself is third-party and out of scope.

Works on CPU (small test graphs) and CUDA (the 114.6 M-edge Reddit shape is
generated on the GPU in a few seconds); the torch RNG stream differs per device
type, so a given (shape, seed) is reproducible per device type.
"""
from collections import namedtuple

import torch

SHAPES = {
    # name: (nodes, undirected edges, feats, classes, parts, p_in, train/val/test fractions)
    'cora': (2708, 5278, 1433, 7, 10, 0.6, (140 / 2708, 500 / 2708, 1000 / 2708)),
    'pubmed': (19717, 44324, 500, 3, 50, 0.6, (60 / 19717, 500 / 19717, 1000 / 19717)),
    'reddit': (232965, 57307946, 602, 41, 1500, 0.25, (153431 / 232965, 23831 / 232965, 55703 / 232965)),
    'amazon2m': (2449029, 30929570, 100, 47, 15000, 0.5, (0.70, 0.05, 0.25)),
}

Synthetic = namedtuple('Synthetic', ['src', 'dst', 'num_nodes', 'feat', 'label', 'train_mask',
                                     'val_mask', 'test_mask', 'part', 'num_classes'])


def dc_sbm_edges(n, m_und, n_blocks, seed, device='cpu', p_in=0.5, alpha=2.3, oversample=1.25):
    """Returns (src, dst, block): 2*m_und directed edges (both directions of m_und
    distinct unordered pairs) and the node->block vector."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
    w = (1.0 - u).clamp_min(1e-9).pow(-1.0 / (alpha - 1.0)).clamp(max=float(n) ** 0.5)
    block = torch.randint(0, n_blocks, (n,), generator=gen, device=dev)
    order = torch.argsort(block, stable=True)              # nodes grouped by block
    cw = torch.cumsum(w[order], 0)
    total = cw[-1]
    bcount = torch.bincount(block, minlength=n_blocks)
    bend = torch.cumsum(bcount, 0)                          # end position of each block in `order`
    bstart = bend - bcount
    cw0 = torch.cat([cw.new_zeros(1), cw])                  # cw0[i] = weight before position i
    need = m_und
    pairs = None
    for _ in range(8):
        k = int(need * oversample) + 1024
        r = torch.rand(k, generator=gen, device=dev, dtype=torch.float64) * total
        pu = torch.searchsorted(cw, r).clamp(max=n - 1)
        b = block[order[pu]]
        lo, hi = cw0[bstart[b]], cw0[bend[b]]
        inblk = torch.rand(k, generator=gen, device=dev) < p_in
        r2 = torch.rand(k, generator=gen, device=dev, dtype=torch.float64)
        r2 = torch.where(inblk, lo + r2 * (hi - lo), r2 * total)
        pv = torch.searchsorted(cw, r2).clamp(max=n - 1)
        a, c = order[pu], order[pv]
        keep = a != c
        a, c = a[keep], c[keep]
        key = torch.minimum(a, c) * n + torch.maximum(a, c)
        pairs = key if pairs is None else torch.cat([pairs, key])
        pairs = torch.unique(pairs)
        if pairs.numel() >= m_und:
            break
        need = m_und - pairs.numel()
    if pairs.numel() > m_und:
        sel = torch.randperm(pairs.numel(), generator=gen, device=dev)[:m_und]
        pairs = pairs[sel]
    a = torch.div(pairs, n, rounding_mode='floor')
    c = pairs % n
    return torch.cat([a, c]), torch.cat([c, a]), block


def make(shape, seed=0, device='cpu', scale=1.0, self_loops=False, feat_dim=None):
    """Synthetic dataset of a named shape (optionally scaled down for tests)."""
    n, m_und, f, ncls, parts, p_in, fr = SHAPES[shape]
    if scale != 1.0:
        n = max(int(n * scale), 64)
        m_und = max(int(m_und * scale), 64)
        parts = max(int(parts * scale), 4)
        m_und = min(m_und, n * (n - 1) // 4)
    if feat_dim is not None:
        f = feat_dim
    dev = torch.device(device)
    src, dst, block = dc_sbm_edges(n, m_und, parts, seed, dev, p_in=p_in)
    if self_loops:
        loops = torch.arange(n, device=dev)
        src, dst = torch.cat([src, loops]), torch.cat([dst, loops])
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed + 1)
    feat = torch.randn(n, f, generator=gen, device=dev)
    label = torch.randint(0, ncls, (n,), generator=gen, device=dev)
    r = torch.rand(n, generator=gen, device=dev)
    train = r < fr[0]
    val = (r >= fr[0]) & (r < fr[0] + fr[1])
    test = (r >= fr[0] + fr[1]) & (r < fr[0] + fr[1] + fr[2])
    return Synthetic(src, dst, n, feat, label, train, val, test, block, ncls)


def to_gist_graph(ds, device=None):
    """GistGraph carrying the dataset's ndata (feat, label, masks, _part)."""
    from .graph import GistGraph
    g = GistGraph.from_edges(ds.src, ds.dst, ds.num_nodes, device=device)
    dev = g.device
    g.ndata['feat'] = ds.feat.to(dev)
    g.ndata['label'] = ds.label.to(dev)
    g.ndata['train_mask'] = ds.train_mask.to(dev)
    g.ndata['val_mask'] = ds.val_mask.to(dev)
    g.ndata['test_mask'] = ds.test_mask.to(dev)
    g.ndata['_part'] = ds.part.to(dev)
    return g
