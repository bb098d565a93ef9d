"""Partition producer and its on-disk format.

``metis_assignment`` replaces ``dgl.transform.metis_partition`` as the reference calls it
(cluster_gcn/partition_utils.py:11-18): a k-way METIS partition of the training graph.  METIS
5.x comes from the CUDA toolkit's ``libmetis_static.a`` through the host-only C-ABI library
``csrc/libgist_partition.so`` (include/gist_partition.h).  Partition *assignment* parity with
DGL's bundled METIS is unpinned (seeded third-party heuristic); what is pinned
is everything downstream: given an assignment, ``partition_list`` yields the reference's
``par_li`` (per part: node ids ascending, parts in id order), and ``save_partition`` /
``load_partition`` read and write the reference's cache file ``../data/{dn}_{psize}.npy``
(sampler.py:44-51: ``np.save`` of the ragged list of int64 arrays, i.e. a pickled object
array) so cached partitions from a real GIST run replay here bit-for-bit.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PARTLIB_PATH = os.path.join(_HERE, 'csrc', 'libgist_partition.so')

_P, _I64 = ctypes.c_void_p, ctypes.c_int64
# name -> (restype, argtypes); must list every symbol include/gist_partition.h declares
SIGNATURES = {
    'gist_partition_symmetrize': (ctypes.c_int, [_I64, _P, _P, _P, _P]),
    'gist_metis_part_kway': (ctypes.c_int, [_I64, _P, _P, _I64, _I64, _P, _P]),
}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(PARTLIB_PATH):
            raise RuntimeError('gist_b200: %s not found; build it with `make -C gist_b200/csrc`' % PARTLIB_PATH)
        lib = ctypes.CDLL(PARTLIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def symmetrize(rowptr, col):
    """Undirected simple CSR (xadj, adjncy: int64) of a directed in-CSR (int32 arrays)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    n = len(rowptr) - 1
    lib = load()
    xadj = np.zeros(n + 1, dtype=np.int64)
    st = lib.gist_partition_symmetrize(n, _np_ptr(rowptr), _np_ptr(col), _np_ptr(xadj), None)
    if st != 0:
        raise RuntimeError('gist_partition_symmetrize failed: status %d' % st)
    adjncy = np.zeros(max(int(xadj[n]), 1), dtype=np.int64)
    st = lib.gist_partition_symmetrize(n, _np_ptr(rowptr), _np_ptr(col), _np_ptr(xadj), _np_ptr(adjncy))
    if st != 0:
        raise RuntimeError('gist_partition_symmetrize failed: status %d' % st)
    return xadj, adjncy[:int(xadj[n])]


def metis_assignment(g, psize, seed=0, return_edgecut=False):
    """node -> part vector (int64 [n]) of a k-way METIS partition of graph ``g`` (a GistGraph, on
    any device; the structure is copied to the host once — this is setup, not the hot path)."""
    rowptr = g.rowptr.detach().cpu().numpy()
    col = g.col.detach().cpu().numpy()
    n = len(rowptr) - 1
    xadj, adjncy = symmetrize(rowptr, col)
    part = np.zeros(n, dtype=np.int64)
    cut = ctypes.c_int64(0)
    st = load().gist_metis_part_kway(n, _np_ptr(xadj), _np_ptr(adjncy) if len(adjncy) else None, int(psize),
                                     int(seed), _np_ptr(part), ctypes.byref(cut))
    if st != 0:
        raise RuntimeError('gist_metis_part_kway failed: status %d' % st)
    return (part, int(cut.value)) if return_edgecut else part


def partition_list(part, psize):
    """par_li as partition_utils.py:11-18 builds it from DGL's {part id: subgraph} dict: one
    int64 array per part id 0..psize-1 holding that part's node ids in ascending order."""
    part = np.asarray(part).astype(np.int64)
    order = np.argsort(part, kind='stable')
    bounds = np.searchsorted(part[order], np.arange(psize + 1))
    return [order[bounds[p]:bounds[p + 1]].astype(np.int64) for p in range(psize)]


def cache_path(dn, psize, cache_dir='../data/'):
    return os.path.join(cache_dir, dn + '_{}.npy'.format(psize))        # sampler.py:45


def save_partition(fn, par_li):
    """np.save(fn, par_li) as sampler.py:51 does under numpy 1.19: a 1-D object array of int64
    arrays (newer numpy refuses to build a ragged array implicitly, so it is built explicitly)."""
    d = os.path.dirname(fn)
    if d:
        os.makedirs(d, exist_ok=True)
    arr = np.empty(len(par_li), dtype=object)
    for i, p in enumerate(par_li):
        arr[i] = np.asarray(p).astype(np.int64)
    np.save(fn, arr, allow_pickle=True)


def load_partition(fn):
    """sampler.py:47: np.load(fn, allow_pickle=True) -> list of int64 arrays.  Also accepts the
    2-D int array numpy 1.19 produces when every part happens to have the same size."""
    arr = np.load(fn, allow_pickle=True)
    return [np.asarray(p).astype(np.int64).reshape(-1) for p in arr]
