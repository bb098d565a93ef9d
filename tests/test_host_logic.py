"""CPU: host-side logic of the product package + C-ABI surface (no compute calls)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def T(x):
    return torch.from_numpy(np.asarray(x))


def cpu_slice_ops():
    """Checker executors (torch indexing) injected in place of the CUDA K5 kernels so the
    host-side plan can be tested without a GPU.  Test infrastructure only."""
    def gather(src, ridx=None, cidx=None):
        out = src
        if src.dim() == 1:
            return out[cidx] if cidx is not None else out.clone()
        if ridx is not None:
            out = out[ridx, :]
        if cidx is not None:
            out = out[:, cidx]
        return out.clone()

    def scatter_(dst, src, ridx=None, cidx=None):
        if dst.dim() == 1:
            if cidx is None:
                dst.copy_(src)
            else:
                dst[cidx] = src
            return dst
        if ridx is not None and cidx is not None:
            tmp = dst[:, cidx]
            tmp[ridx, :] = src
            dst[:, cidx] = tmp
        elif ridx is not None:
            dst[ridx, :] = src
        elif cidx is not None:
            dst[:, cidx] = src
        else:
            dst.copy_(src)
        return dst
    return gather, scatter_


def test_library_exports_every_declared_symbol():
    from gist_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'gist_b200.h')).read()
    declared = set(re.findall(r'\b(gist_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gist_abi_version() == 1
    assert lib.gist_status_string(-1) == b'bad argument'
    assert lib.gist_scan_workspace_bytes(100) == 0
    assert lib.gist_scan_workspace_bytes(10 ** 6) >= (10 ** 6 // 2048) * 4
    assert isinstance(_lib.launch_count(), int)


def test_python_constants_match_the_header():
    """The flag / enum values the ctypes host passes are the header's #defines, and the ctypes
    argument lists have one entry per declared parameter."""
    from gist_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'gist_b200.h')).read()
    defs = {m.group(1): m.group(2) for m in re.finditer(r'#define\s+GIST_([A-Z0-9_]+)\s+\(?(-?\d+)u?\)?', header)}
    for py, c in (('SPMM_RELU', 'SPMM_RELU'), ('SPMM_NARROW', 'SPMM_NARROW'), ('SPMM_WIDE', 'SPMM_WIDE'),
                  ('SPMM_BG_SHIFT', 'SPMM_BG_SHIFT'), ('NORM_INV', 'NORM_INV'),
                  ('NORM_RSQRT_CLAMP', 'NORM_RSQRT_CLAMP'), ('GEMM_RELU', 'GEMM_RELU'),
                  ('GEMM_NO_SPLITK', 'GEMM_NO_SPLITK'), ('GEMM_TILE_N64', 'GEMM_TILE_N64'),
                  ('GEMM_TILE_N128', 'GEMM_TILE_N128'), ('GEMM_TILE_N256', 'GEMM_TILE_N256'),
                  ('GEMM_K_MAJOR', 'GEMM_K_MAJOR'), ('GEMM_MN_MAJOR', 'GEMM_MN_MAJOR'), ('ACT_RELU', 'ACT_RELU')):
        assert getattr(_lib, py) == int(defs[c]), (py, defs[c])
    # parameter counts: text between the parentheses of every prototype
    flat = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r'\b%s\s*\(([^)]*)\)\s*;' % name, flat)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        assert n == len(argtypes), (name, n, len(argtypes))


def test_compute_on_cpu_tensor_fails_loudly():
    from gist_b200 import GistGraph, ops
    from gist_b200._lib import GistLibraryError
    g = GistGraph.from_edges(torch.tensor([0, 1]), torch.tensor([1, 2]), 3)
    with pytest.raises(GistLibraryError):
        ops.copy_src_sum(g, torch.ones(3, 4))
    with pytest.raises(GistLibraryError):
        g.subgraph(np.array([0, 1]))
    with pytest.raises(GistLibraryError):
        g.inv_in_degree()


def test_gist_graph_structure_on_cpu():
    from gist_b200 import GistGraph
    src = torch.tensor([0, 2, 2, 3, 1])
    dst = torch.tensor([1, 1, 1, 3, 0])
    g = GistGraph.from_edges(src, dst, 4)
    assert g.rowptr.tolist() == [0, 1, 4, 4, 5]
    assert g.col.tolist() == [1, 0, 2, 2, 3]
    assert g.number_of_edges() == 5 and g.number_of_nodes() == 4
    assert g.in_degrees().tolist() == [1, 3, 0, 1]
    assert g.out_degrees().tolist() == [1, 1, 2, 1]
    colptr, row = g.csc()
    assert colptr.tolist() == [0, 1, 2, 4, 5] and row.tolist() == [1, 0, 1, 1, 3]
    assert not g.is_symmetric()
    s = GistGraph.from_edges(torch.cat([src, dst]), torch.cat([dst, src]), 4)
    assert s.is_symmetric() and s.csc()[0] is s.rowptr
    lv = g.local_var()
    lv.ndata['h'] = torch.zeros(4, 2)
    assert 'h' not in g.ndata
    assert g.long().idtype == torch.int64 and g.int().idtype == torch.int32
    assert g.in_degrees().dtype == torch.int64 and g.int().in_degrees().dtype == torch.int32
    with pytest.raises(Exception):
        g.ndata['bad'] = torch.zeros(3)


def test_create_partition_matches_reference_golden():
    from gist_b200 import create_partition
    G = np.load(os.path.join(GOLD, 'partition.npz'))
    for t, (seed, m, size) in enumerate(G['cp_triples']):
        random.seed(int(seed))
        for tag in 'ab':
            part = create_partition(int(m), int(size))
            assert np.array_equal(np.stack([p[0].numpy() for p in part]), G['cp%d%s_idx' % (t, tag)])
            assert np.array_equal(np.stack([p[1].numpy() for p in part]), G['cp%d%s_full' % (t, tag)])
            assert part[0][0].dtype == torch.int64


def test_container_shapes_and_keys_match_reference():
    import torch.nn.functional as F
    from gist_b200 import SageGCN, GCN
    G = np.load(os.path.join(GOLD, 'sage.npz'))
    for ci in range(3):
        fin, hid, ncls, L, ln = (int(v) for v in G['sage%d_cfg' % ci])
        model = SageGCN(fin, hid, ncls, L, F.relu, 0.3, bool(ln), False, False, 1, True)
        ref = {k[len('sage%d_param.' % ci):]: G[k].shape for k in G.files if k.startswith('sage%d_param.' % ci)}
        assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == ref
    G = np.load(os.path.join(GOLD, 'graphconv.npz'))
    for ci in range(3):
        fin, hid, ncls, L, si, so, k = (int(v) for v in G['gc%d_cfg' % ci])
        model = GCN(None, fin, hid, ncls, L, F.relu, 0.5, True, bool(si), bool(so), k)
        ref = {kk[len('gc%d_param.' % ci):]: G[kk].shape for kk in G.files if kk.startswith('gc%d_param.' % ci)}
        assert {kk: tuple(v.shape) for kk, v in model.state_dict().items()} == ref


def test_module_init_consumes_rng_like_reference():
    """Same seed -> bit-identical initial parameters as the reference's modules (golden)."""
    import torch.nn.functional as F
    from gist_b200 import SageGCN
    G = np.load(os.path.join(GOLD, 'sage.npz'))
    for ci, seed in enumerate([1, 2, 3]):
        fin, hid, ncls, L, ln = (int(v) for v in G['sage%d_cfg' % ci])
        n = int(G['sage%d_n' % ci])
        torch.manual_seed(seed)
        torch.randn(n, fin); torch.randint(0, ncls, (n,))      # noqa: E702 - the generator drew x, y first
        model = SageGCN(fin, hid, ncls, L, F.relu, 0.3, bool(ln), False, False, 1, True)
        for k, v in model.state_dict().items():
            assert np.array_equal(v.numpy(), G['sage%d_param.%s' % (ci, k)]), k


@pytest.mark.parametrize('ci', [0, 1, 2, 3])
def test_graphconv_split_merge_plan_vs_reference_golden(ci):
    from gist_b200 import ist_graphconv as IG
    G = np.load(os.path.join(GOLD, 'train_ist.npz'))
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
            sub = IG.split_state_dict(main, feats_idx, s, L, bool(si), bool(so), slice_ops=cpu_slice_ops())
            for k in keys:
                assert np.array_equal(sub[k].numpy(), G[p + 'split%d.%s' % (r * m + s, k)]), (r, s, k)
        trained = [{k: T(G[p + 'r%d_trained%d.%s' % (r, s, k)]) for k in keys} for s in range(m)]
        merged = IG.merge_state_dicts(main, feats_idx, trained, L, bool(si), bool(so), slice_ops=cpu_slice_ops())
        for k in keys:
            assert np.allclose(merged[k].numpy(), G[p + 'r%d_merged.%s' % (r, k)], rtol=1e-6, atol=1e-7), k


def test_sample_feature_partitions_rng_order():
    from gist_b200 import ist_graphconv as IG
    torch.manual_seed(5)
    a = IG.sample_feature_partitions(12, 8, 2, 4, True, True)
    torch.manual_seed(5)
    b = [torch.chunk(torch.randperm(12), 4), torch.chunk(torch.randperm(8), 4), torch.chunk(torch.randperm(8), 4)]
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            assert torch.equal(u, v)
    c = IG.sample_feature_partitions(12, 8, 2, 4, False, False)
    assert c[0] is None and c[-1] is None and len(c) == 3


def test_synth_shapes():
    from gist_b200 import synth
    ds = synth.make('cora', seed=0)
    assert ds.num_nodes == 2708 and ds.src.shape[0] == 10556 and ds.feat.shape == (2708, 1433)
    key = ds.src * ds.num_nodes + ds.dst
    assert torch.unique(key).numel() == key.numel()            # no multi-edges
    assert (ds.src != ds.dst).all()                              # no self loops
    rev = ds.dst * ds.num_nodes + ds.src
    assert torch.equal(torch.sort(key)[0], torch.sort(rev)[0])   # symmetric
    ds2 = synth.make('cora', seed=0)
    assert torch.equal(ds.src, ds2.src) and torch.equal(ds.feat, ds2.feat)
    dl = synth.make('cora', seed=0, self_loops=True)
    assert dl.src.shape[0] == 10556 + 2708


def test_persistent_weight_low_halves_are_served_only_while_valid(monkeypatch):
    """ops' registry of persistent 3xTF32 low halves (the fused Adam launch keeps them current): an entry is
    served only while no raw-pointer writer of this library (note_raw_write) and no torch in-place op
    (Tensor._version) has touched the weight since the last refresh; dead parameters drop out."""
    import gc
    import torch
    from gist_b200 import ops
    calls = []
    monkeypatch.setattr(ops, '_tma_ok', lambda t: t.dim() == 2)
    monkeypatch.setattr(ops, 'split_tf32_multi', lambda ws, los: calls.append(len(ws)))
    monkeypatch.setattr(ops, '_WEIGHT_LO_PERSIST', {})
    monkeypatch.setattr(ops, '_PERSIST_EPOCH', [0, -1])
    p = torch.nn.Parameter(torch.randn(8, 12))
    q = torch.nn.Parameter(torch.randn(4, 12))
    lo_p, lo_q = torch.zeros(8, 12), torch.zeros(4, 12)
    assert ops.register_persistent_lo(p, lo_p) is lo_p
    assert ops.register_persistent_lo(q, lo_q) is lo_q
    assert ops.register_persistent_lo(p, torch.zeros(8, 12)) is lo_p       # a second owner of the same parameter shares it
    assert not ops.persistent_lo_valid() and ops._persistent_lo(p.data) is None     # never refreshed
    ops.refresh_persistent_lo()
    assert calls == [2] and ops.persistent_lo_valid()
    assert ops._persistent_lo(p) is lo_p and ops._persistent_lo(p.data) is lo_p and ops._persistent_lo(q) is lo_q
    assert ops._persistent_lo(torch.randn(8, 12)) is None                          # some other tensor
    with torch.no_grad():
        p.mul_(2.0)                                                                 # torch in-place op: version bump
    assert not ops.persistent_lo_valid() and ops._persistent_lo(p) is None
    assert ops._persistent_lo(q) is lo_q                                            # q itself is untouched
    ops.refresh_persistent_lo()
    assert ops._persistent_lo(p) is lo_p
    ops.note_raw_write(torch.randn(3, 3))                                           # not a registered weight: nothing happens
    assert ops.persistent_lo_valid()
    ops.note_raw_write(q.data)                                                      # a K5 kernel wrote q through its pointer
    assert not ops.persistent_lo_valid() and ops._persistent_lo(q) is None and ops._persistent_lo(p) is None
    ops.refresh_persistent_lo()
    assert ops.persistent_lo_valid() and calls == [2, 2, 2]
    del p
    gc.collect()
    ops.refresh_persistent_lo()
    assert calls[-1] == 1 and len(ops._WEIGHT_LO_PERSIST) == 1                      # the dead parameter dropped out


def test_padded_epoch_ids_row_copies_equal_the_per_part_form():
    """ClusterIter.padded_epoch_ids (one slice copy per batch row) against the literal per-part loop, including a
    part count that is not a multiple of the batch size and empty parts."""
    from types import SimpleNamespace
    from gist_b200.sampler import ClusterIter
    rng = np.random.RandomState(4)
    for psize, bs in [(1500, 20), (37, 5), (7, 3), (4, 8), (1, 1)]:
        par_li = [np.sort(rng.choice(100000, int(s), replace=False)).astype(np.int64) for s in rng.randint(0, 40, psize)]
        mx = max(psize // bs, 1) if psize >= bs else 1
        it = SimpleNamespace(par_li=par_li, batch_size=bs, psize=psize, max=mx)
        n_pad = int(sum(sorted((len(p) for p in par_li), reverse=True)[:bs])) + 3
        ref = np.full((mx, n_pad), -1, dtype=np.int64)
        for i in range(mx):
            pos = 0
            for s in range(i * bs, min((i + 1) * bs, psize)):
                ref[i, pos:pos + len(par_li[s])] = par_li[s]
                pos += len(par_li[s])
        got = ClusterIter.padded_epoch_ids(it, n_pad).numpy()
        assert np.array_equal(got, ref), (psize, bs)
        buf = np.zeros((mx, n_pad), dtype=np.int64)
        assert ClusterIter.padded_epoch_ids(it, n_pad, out=buf).numpy() is not None and np.array_equal(buf, ref)
