"""GistGraph — the device-resident graph object behind the reference's module API.

It exposes the duck-typed slice of ``DGLGraph`` that GIST's models, sampler and
trainers touch: ``ndata``, ``local_var``, ``in_degrees``,
``update_all(copy_src, sum)``, ``subgraph``, ``to``, ``long``/``int``, ``cpu``,
``number_of_nodes``/``number_of_edges``.  Storage is an int32 in-edge CSR
(``rowptr[n+1]``, ``col[nnz]``: row v lists the sources of edges u->v) plus, on
demand, the CSC (out-edge lists) used by the backward SpMM.  No edge values.

Structure may live on the CPU (construction, host-side tests) but every compute
method (``update_all``, ``subgraph``, norms) requires CUDA and raises otherwise —
there is no CPU path.
"""
import os

import numpy as np
import torch

from . import _lib, function as fn, ops

NID = '_ID'


# A/B switch for the segment-balanced cluster-batch SpMM (GIST_SPMM_SEG=0: one row per group)
SEG_ENABLED = os.environ.get('GIST_SPMM_SEG', '1') != '0'
# K3 variant: 'rows' = one warp per batch row (gist_cluster_batch_build), 'chunks' = one warp per 128-edge chunk of
# the batch's parent rows (gist_cluster_batch_build_v2; same CSR).  Chunks need a bound on the batch's parent-degree
# sum (``walk_capacity``; it covers in-edge rows, so the CSC pass of a directed graph keeps the row kernel); very
# large node sets keep the row kernel too (the chunk counts are scanned by one CTA).
BUILDER = os.environ.get('GIST_BUILDER', 'chunks')
MAX_CHUNKS_V2 = 1 << 17

class GistError(RuntimeError):
    """Counterpart of dgl.DGLError."""


class NodeFrame(dict):
    def __init__(self, n, *a, **k):
        super().__init__(*a, **k)
        self._n = n

    def __setitem__(self, key, val):
        if val.shape[0] != self._n:
            raise GistError('Expect number of features to match number of nodes. Got %d and %d instead.'
                            % (val.shape[0], self._n))
        # device feature matrices live in HBM with 16-byte-aligned rows (ops.pad_rows): same values,
        # same shape, padded leading dimension — 128-bit gathers and direct TMA addressing
        super().__setitem__(key, ops.pad_rows(val))


class GistGraph:
    def __init__(self, rowptr, col, num_nodes=None, *, csc=None, symmetric=None,
                 idtype=torch.int64, nnz=None):
        assert rowptr.dtype == torch.int32 and col.dtype == torch.int32
        self.rowptr = rowptr
        self.col_buffer = col            # may be longer than nnz (capacity of a built batch)
        self._n = int(num_nodes) if num_nodes is not None else rowptr.shape[0] - 1
        self._nnz = nnz                  # None -> read rowptr[n] lazily (one sync)
        self._csc = csc                  # (colptr, row) or None
        self._symmetric = symmetric      # True -> CSC arrays == CSR arrays
        self._idtype = idtype
        self.ndata = NodeFrame(self._n)
        self._cache = {}                 # degree norms, node_map scratch, zero-in-degree flag

    # ------------------------------------------------------------ builders --
    @staticmethod
    def from_edges(src, dst, num_nodes, device=None, idtype=torch.int64):
        """COO (u -> v) to canonical in-CSR: rows by dst, columns ascending.
        One-time graph setup; uses torch sort on whatever device the edges are on."""
        src = torch.as_tensor(src).long().reshape(-1)
        dst = torch.as_tensor(dst).long().reshape(-1)
        if device is not None:
            src, dst = src.to(device), dst.to(device)
        n = int(num_nodes)
        key = dst * n + src
        key, _ = torch.sort(key)
        col = (key % n).to(torch.int32)
        deg = torch.bincount(torch.div(key, n, rounding_mode='floor'), minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=src.device)
        rowptr[1:] = torch.cumsum(deg, 0)
        return GistGraph(rowptr.to(torch.int32), col, n, idtype=idtype, nnz=int(src.shape[0]))

    @staticmethod
    def from_scipy(sp_mat, device=None, idtype=torch.int64):
        """dgl.from_scipy (AmazonDataset.py:111): pattern only, edge row -> col."""
        coo = sp_mat.tocoo()
        return GistGraph.from_edges(torch.from_numpy(coo.row.astype(np.int64)),
                                    torch.from_numpy(coo.col.astype(np.int64)), coo.shape[0],
                                    device=device, idtype=idtype)

    @staticmethod
    def from_networkx(g, device=None):
        """DGLGraph(nx_graph) (gcn/train.py:69): undirected graphs give both directions."""
        n = g.number_of_nodes()
        if not g.is_directed():
            g = g.to_directed()
        e = np.asarray(list(g.edges()), dtype=np.int64).reshape(-1, 2)
        return GistGraph.from_edges(torch.from_numpy(e[:, 0].copy()), torch.from_numpy(e[:, 1].copy()),
                                    n, device=device)

    # ------------------------------------------------------- DGL duck type --
    @property
    def device(self):
        return self.rowptr.device

    @property
    def idtype(self):
        return self._idtype

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self):
        if self._nnz is None:
            self._nnz = int(self.rowptr[self._n].item())
        return self._nnz

    num_edges = number_of_edges

    @property
    def col(self):
        """Column indices trimmed to nnz (syncs once for a freshly built batch)."""
        return self.col_buffer[:self.number_of_edges()]

    def _shallow(self, idtype=None):
        g = GistGraph(self.rowptr, self.col_buffer, self._n, csc=self._csc, symmetric=self._symmetric,
                      idtype=idtype or self._idtype, nnz=self._nnz)
        g.ndata = NodeFrame(self._n, self.ndata)
        g._cache = self._cache           # structure-derived, shareable
        return g

    def local_var(self):
        return self._shallow()

    def long(self):
        return self._shallow(torch.int64)

    def int(self):
        return self._shallow(torch.int32)

    def to(self, device, non_blocking=False):
        device = torch.device(device)
        if device == self.device and all(v.device == device for v in self.ndata.values()):
            return self._shallow()
        mv = lambda t: t.to(device, non_blocking=non_blocking)  # noqa: E731
        csc = (mv(self._csc[0]), mv(self._csc[1])) if self._csc is not None else None
        g = GistGraph(mv(self.rowptr), mv(self.col_buffer), self._n, csc=csc,
                      symmetric=self._symmetric, idtype=self._idtype, nnz=self._nnz)
        for k, v in self.ndata.items():
            g.ndata[k] = mv(v)
        return g

    def cpu(self):
        return self.to('cpu')

    def in_degrees(self):
        d = self.rowptr[1:] - self.rowptr[:-1]
        return d.to(self._idtype)

    def out_degrees(self):
        colptr, _ = self.csc()
        return (colptr[1:] - colptr[:-1]).to(self._idtype)

    # ----------------------------------------------------------- structure --
    def csc(self):
        """(colptr, row): out-edge lists, rows ascending per column. Built once."""
        if self._csc is None:
            if self._symmetric:
                self._csc = (self.rowptr, self.col_buffer)
            else:
                n = self._n
                col = self.col.long()
                deg_in = (self.rowptr[1:] - self.rowptr[:-1]).long()
                row_of_edge = torch.repeat_interleave(torch.arange(n, device=self.device), deg_in)
                order = torch.argsort(col, stable=True)
                row = row_of_edge[order].to(torch.int32)
                colptr = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
                colptr[1:] = torch.cumsum(torch.bincount(col, minlength=n), 0)
                colptr = colptr.to(torch.int32)
                if self._symmetric is None:
                    self._symmetric = bool(torch.equal(colptr, self.rowptr)
                                           and torch.equal(row, self.col.to(torch.int32)))
                self._csc = (self.rowptr, self.col_buffer) if self._symmetric else (colptr, row)
        return self._csc

    def is_symmetric(self):
        if self._symmetric is None:
            self.csc()
        return bool(self._symmetric)

    def inv_in_degree(self):
        """[n] fp32: 1/in_degree, 0 for isolated nodes (get_norm, modules.py:239-243)."""
        if 'inv_in' not in self._cache:
            self._cache['inv_in'] = ops.degree_norm(self.rowptr, self._n, _lib.NORM_INV)
        return self._cache['inv_in']

    def rsqrt_in_degree(self):
        if 'rsqrt_in' not in self._cache:
            self._cache['rsqrt_in'] = ops.degree_norm(self.rowptr, self._n, _lib.NORM_RSQRT_CLAMP)
        return self._cache['rsqrt_in']

    def rsqrt_out_degree(self):
        if 'rsqrt_out' not in self._cache:
            colptr, _ = self.csc()
            self._cache['rsqrt_out'] = ops.degree_norm(colptr, self._n, _lib.NORM_RSQRT_CLAMP)
        return self._cache['rsqrt_out']

    # rows are cut into segments for the balanced SpMM only where rows are scarce (cluster batches)
    SEG_MAX_NODES = 32768

    def seg_schedule(self, transpose=False):
        """ops.SegSchedule of the in-CSR (or of the CSC for the transpose SpMM), built once per
        structure; None for large graphs, where one row per warp already balances."""
        if self._n == 0 or self._n > self.SEG_MAX_NODES or not self.rowptr.is_cuda or not SEG_ENABLED:
            return None
        if transpose and not self.is_symmetric():
            key, (ptr_, idx_) = 'seg_sched_t', self.csc()
        else:
            key, ptr_, idx_ = 'seg_sched', self.rowptr, self.col_buffer
        sch = self._cache.get(key)
        if sch is None:
            sch = self._cache[key] = ops.SegSchedule(ptr_, self._n, int(idx_.shape[0]))
        return sch

    def has_zero_in_degree(self):
        if 'zero_in' not in self._cache:
            self._cache['zero_in'] = bool(((self.rowptr[1:] - self.rowptr[:-1]) == 0).any().item())
        return self._cache['zero_in']

    # ------------------------------------------------------ message passing --
    def update_all(self, message_func, reduce_func, apply_node_func=None):
        """Only the builtin pair the reference uses: copy_src + sum."""
        if not (isinstance(message_func, fn.CopySrc) and isinstance(reduce_func, fn.Sum)
                and message_func.out == reduce_func.msg and apply_node_func is None):
            raise GistError('gist_b200 implements update_all(fn.copy_src, fn.sum) only')
        h = self.ndata[message_func.src]
        h2 = h.reshape(h.shape[0], -1)
        out = ops.copy_src_sum(self, h2)
        self.ndata[reduce_func.out] = out.reshape((self._n,) + tuple(h.shape[1:]))

    # -------------------------------------------------------- K3: subgraph --
    def _node_map(self):
        if 'node_map' not in self._cache:
            self._cache['node_map'] = torch.full((self._n,), -1, dtype=torch.int32, device=self.device)
        return self._cache['node_map']

    def subgraph(self, nids, col_capacity=None, gather_ndata=True, ndata_keys=None, out=None, walk_capacity=None):
        """Node-induced subgraph built on the device (K3).  new node i <-> nids[i].

        ``col_capacity``: capacity of the edge array (default: the ids' parent-degree sum, one sync).
        ``walk_capacity``: a bound on the ids' parent-degree sum, if the caller has one (ClusterIter: the sum
        of the parts' degree sums) — what the chunked builder sizes its chunk arrays from.

        ``out``: a subgraph previously returned for the same number of ids, capacity and
        ndata keys; its buffers are overwritten in place (fixed addresses: what the pipelined
        CUDA-graph trainer needs to build batch k+1 while batch k trains)."""
        _lib.require_cuda(self.rowptr)
        dev = self.device
        if isinstance(nids, np.ndarray):
            nids = torch.from_numpy(np.ascontiguousarray(nids.reshape(-1).astype(np.int64)))
        nids = torch.as_tensor(nids).reshape(-1)
        if nids.dtype != torch.int64:
            nids = nids.long()
        if nids.device != dev:
            nids = nids.to(dev, non_blocking=True)
        nids = nids.contiguous()
        n_b = int(nids.shape[0])
        if col_capacity is None:
            deg = (self.rowptr[1:] - self.rowptr[:-1])
            col_capacity = int(deg[nids.clamp_min(0)].sum().item()) if n_b else 0
            walk_capacity = col_capacity

        if out is not None:
            assert out.number_of_nodes() == n_b and out.col_buffer.shape[0] == max(col_capacity, 1)

        def build(prow, pcol, want_inv, reuse=None):
            lib = _lib.load()
            if reuse is not None:
                rowptr, col, inv = reuse
            else:
                rowptr = torch.empty(n_b + 1, dtype=torch.int32, device=dev)
                col = torch.empty(max(col_capacity, 1), dtype=torch.int32, device=dev)
                inv = torch.empty(max(n_b, 1), dtype=torch.float32, device=dev) if want_inv else None
            max_chunks = (walk_capacity // 128 + n_b + 1) if walk_capacity is not None else None
            if (BUILDER == 'chunks' and max_chunks is not None and max_chunks <= MAX_CHUNKS_V2 and n_b > 0
                    and prow is self.rowptr):      # the in-edge pass: the rows walk_capacity bounds
                wsb = lib.gist_cluster_batch_build_v2_workspace_bytes(n_b, max_chunks)
                ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
                _lib.check(lib.gist_cluster_batch_build_v2(
                    _lib.ptr(prow), _lib.ptr(pcol), self._n, _lib.ptr(nids), n_b, _lib.ptr(self._node_map()),
                    _lib.ptr(rowptr), _lib.ptr(col), col_capacity, _lib.ptr(inv), None, max_chunks,
                    _lib.ptr(ws), wsb, _lib.stream_ptr(dev)), 'cluster_batch_build_v2')
                return rowptr, col, inv
            wsb = lib.gist_scan_workspace_bytes(n_b)
            ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
            _lib.check(lib.gist_cluster_batch_build(
                _lib.ptr(prow), _lib.ptr(pcol), self._n, _lib.ptr(nids), n_b, _lib.ptr(self._node_map()),
                _lib.ptr(rowptr), _lib.ptr(col), col_capacity, _lib.ptr(inv), None,
                _lib.ptr(ws), wsb, _lib.stream_ptr(dev)), 'cluster_batch_build')
            return rowptr, col, inv

        if out is not None:
            sg = out
            inv_buf = sg._cache['inv_in_buf']
            build(self.rowptr, self.col_buffer, True, reuse=(sg.rowptr, sg.col_buffer, inv_buf))
            if not self.is_symmetric():
                pcolptr, prow_idx = self.csc()
                build(pcolptr, prow_idx, False, reuse=(sg._csc[0], sg._csc[1], None))
            if gather_ndata:
                for k, v in self.ndata.items():
                    if ndata_keys is None or k in ndata_keys:
                        ops.gather_rows(v, nids, out=sg.ndata[k])
            if sg.ndata[NID].data_ptr() != nids.data_ptr():
                sg.ndata[NID].copy_(nids)
            sg._nnz = None
            keep = ('inv_in', 'inv_in_buf', 'seg_sched', 'seg_sched_t')
            for k in [k for k in sg._cache if k not in keep]:
                del sg._cache[k]                 # structure-derived values of the previous batch
            if 'seg_sched' in sg._cache:         # same buffers, new contents
                sg._cache['seg_sched'].rebuild(sg.rowptr)
            if 'seg_sched_t' in sg._cache:
                sg._cache['seg_sched_t'].rebuild(sg._csc[0])
            return sg
        rowptr, col, inv = build(self.rowptr, self.col_buffer, True)
        sg = GistGraph(rowptr, col, n_b, idtype=self._idtype)
        sg._cache['inv_in'] = inv[:n_b]
        sg._cache['inv_in_buf'] = inv
        if self.is_symmetric():
            # an induced subgraph of a symmetric pattern is symmetric, and it is built
            # from equal arrays in equal order, so CSC == CSR array-for-array
            sg._symmetric, sg._csc = True, (rowptr, col)
        else:
            pcolptr, prow_idx = self.csc()
            cptr, crow, _ = build(pcolptr, prow_idx, False)
            sg._symmetric, sg._csc = False, (cptr, crow)
        if gather_ndata:
            for k, v in self.ndata.items():
                if ndata_keys is None or k in ndata_keys:
                    sg.ndata[k] = ops.gather_rows(v, nids)
        sg.ndata[NID] = nids.to(self._idtype)
        return sg

    def __repr__(self):
        return 'GistGraph(num_nodes=%d, device=%s)' % (self._n, self.device)
