"""Shared helpers for the parity tests."""
import numpy as np
import torch

from oracle import gist_oracle as O

# Tolerance of north_star: fp32 aggregation outputs and gradients within 1e-5
# RELATIVE.  Summation order differs between any two SpMM implementations, so
# "relative" is taken norm-wise (max |a-b| <= RTOL * max |ref|), plus a loose
# element-wise check.
RTOL = 1e-5


def assert_close(got, ref, rtol=RTOL, what=''):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.numel() == 0:
        return
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= rtol * max(scale, 1e-30) + 1e-30, '%s: max|err| %.3e vs scale %.3e (rel %.3e)' % (
        what, err, scale, err / max(scale, 1e-30))


def random_graph(n, nnz, seed, self_loops=True, isolated=0):
    """Random multigraph edge list (u->v) with optional self loops, duplicate edges
    and `isolated` nodes that have no in-edges."""
    rng = np.random.RandomState(seed)
    src = rng.randint(0, n, size=nnz)
    dst = rng.randint(0, max(n - isolated, 1), size=nnz)   # last `isolated` nodes: zero in-degree
    if not self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    # force some duplicates
    if nnz > 8:
        src[:4] = src[4:8]
        dst[:4] = dst[4:8]
    return torch.from_numpy(src.astype(np.int64)), torch.from_numpy(dst.astype(np.int64))


def powerlaw_graph(n, avg_deg, seed):
    rng = np.random.RandomState(seed)
    w = (1 - rng.rand(n)) ** (-1 / 1.2)
    w = w / w.sum()
    nnz = n * avg_deg
    src = rng.choice(n, size=nnz, p=w)
    dst = rng.choice(n, size=nnz, p=w)
    return torch.from_numpy(src.astype(np.int64)), torch.from_numpy(dst.astype(np.int64))


def canonical_csr_from_gist(g):
    """(rowptr, col) int64 on CPU with columns sorted inside each row."""
    rowptr = g.rowptr.cpu().long()
    col = g.col.cpu().long()
    n = g.number_of_nodes()
    row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    key, _ = torch.sort(row * max(n, 1) + col)
    return rowptr, key % max(n, 1)


def ograph(src, dst, n):
    return O.OGraph(src, dst, n)
