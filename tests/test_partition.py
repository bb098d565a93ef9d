"""CPU: the host-side partition producer (METIS through libgist_partition.so) and the
reference's on-disk partition cache format (sampler.py:44-51)."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_library_exports_every_declared_symbol():
    from gist_b200 import partition
    header = open(os.path.join(ROOT, 'include', 'gist_partition.h')).read()
    declared = set(re.findall(r'\b(gist_[a-z0-9_]+)\s*\(', header))
    assert declared == set(partition.SIGNATURES), declared ^ set(partition.SIGNATURES)
    lib = partition.load()
    for name in declared:
        assert hasattr(lib, name), name
    # METIS itself stays private to the library
    assert not hasattr(lib, 'METIS_PartGraphKway')


def _in_csr(src, dst, n):
    A = sp.csr_matrix((np.ones(len(src)), (dst, src)), shape=(n, n))      # row v lists sources u
    A.sort_indices()
    # keep multi-edges: rebuild from sorted COO
    order = np.lexsort((src, dst))
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(rowptr, dst + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), src[order].astype(np.int32)


def test_symmetrize_matches_scipy():
    from gist_b200 import partition
    rng = np.random.RandomState(0)
    n = 300
    src, dst = rng.randint(0, n, 4000), rng.randint(0, n, 4000)
    src[:50] = dst[:50]                        # self loops
    src[50:100], dst[50:100] = src[100:150], dst[100:150]      # multi-edges
    rowptr, col = _in_csr(src, dst, n)
    xadj, adjncy = partition.symmetrize(rowptr, col)
    A = sp.csr_matrix((np.ones(len(src)), (dst, src)), shape=(n, n))
    S = ((A + A.T) > 0).astype(np.int8).tolil()
    S.setdiag(0)
    S = S.tocsr()
    S.eliminate_zeros()
    S.sort_indices()
    assert np.array_equal(xadj, S.indptr.astype(np.int64))
    assert np.array_equal(adjncy, S.indices.astype(np.int64))
    # empty graph / isolated nodes
    xadj, adjncy = partition.symmetrize(np.zeros(6, dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert np.array_equal(xadj, np.zeros(6, dtype=np.int64)) and len(adjncy) == 0


def _planted(n, k, deg_in, deg_out, seed):
    rng = np.random.RandomState(seed)
    block = rng.randint(0, k, n)
    members = [np.nonzero(block == b)[0] for b in range(k)]
    src, dst = [], []
    for v in range(n):
        mine = members[block[v]]
        src.append(np.full(deg_in, v)); dst.append(mine[rng.randint(0, len(mine), deg_in)])
        src.append(np.full(deg_out, v)); dst.append(rng.randint(0, n, deg_out))
    src, dst = np.concatenate(src), np.concatenate(dst)
    return np.concatenate([src, dst]), np.concatenate([dst, src]), block


class _G:
    """The two attributes metis_assignment reads from a GistGraph."""

    def __init__(self, rowptr, col):
        self.rowptr, self.col = torch.from_numpy(rowptr), torch.from_numpy(col)


def test_metis_recovers_planted_communities():
    from gist_b200 import partition
    n, k = 4000, 16
    src, dst, block = _planted(n, k, 12, 1, seed=1)
    rowptr, col = _in_csr(src, dst, n)
    g = _G(rowptr, col)
    part, cut = partition.metis_assignment(g, k, seed=0, return_edgecut=True)
    assert part.shape == (n,) and part.min() >= 0 and part.max() == k - 1
    sizes = np.bincount(part, minlength=k)
    assert sizes.max() <= 1.05 * n / k + 1            # METIS's default 3 % imbalance (+ rounding)
    # edge cut reported == edge cut recomputed on the symmetrized simple graph
    xadj, adjncy = partition.symmetrize(rowptr, col)
    rows = np.repeat(np.arange(n), np.diff(xadj))
    assert cut == int((part[rows] != part[adjncy]).sum()) // 2
    planted_cut = int((block[rows] != block[adjncy]).sum()) // 2
    # the planted blocks are unbalanced (random sizes) while METIS enforces balance, so it has to
    # cut somewhat more than the planted cut — but far less than an arbitrary balanced assignment
    arbitrary_cut = int(((rows % k) != (adjncy % k)).sum()) // 2
    assert cut <= 1.5 * planted_cut and cut <= 0.2 * arbitrary_cut, (cut, planted_cut, arbitrary_cut)
    # deterministic for a fixed seed
    assert np.array_equal(part, partition.metis_assignment(g, k, seed=0))
    # k = 1 and an edgeless graph are served without METIS
    assert partition.metis_assignment(g, 1).max() == 0
    e = _G(np.zeros(11, dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert sorted(np.bincount(partition.metis_assignment(e, 5), minlength=5)) == [2] * 5


def test_partition_list_and_cache_roundtrip(tmp_path):
    from gist_b200 import partition
    part = np.array([2, 0, 1, 2, 2, 0, 1, 0, 0], dtype=np.int64)
    li = partition.partition_list(part, 4)
    assert [p.tolist() for p in li] == [[1, 5, 7, 8], [2, 6], [0, 3, 4], []]
    assert all(p.dtype == np.int64 for p in li)
    fn = partition.cache_path('reddit-self-loop', 4, str(tmp_path))
    assert fn.endswith('reddit-self-loop_4.npy')
    partition.save_partition(fn, li)
    back = partition.load_partition(fn)
    assert [p.tolist() for p in back] == [p.tolist() for p in li]
    # what the reference reads with np.load(fn, allow_pickle=True) (sampler.py:47): a 1-D object
    # array whose items are the int64 id arrays
    raw = np.load(fn, allow_pickle=True)
    assert raw.dtype == object and raw.shape == (4,) and raw[0].dtype == np.int64
    # a file the reference wrote when all parts had equal size is a plain 2-D int array
    fn2 = os.path.join(str(tmp_path), 'eq_2.npy')
    np.save(fn2, np.array([[0, 3], [1, 2]]))
    assert [p.tolist() for p in partition.load_partition(fn2)] == [[0, 3], [1, 2]]
