"""CPU restatement of the slice of DGL 0.5.3 that GIST's hot path calls.

TEST INFRASTRUCTURE ONLY. Nothing under ``oracle/`` may be imported by the
product package ``gist_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs use it, as the checker.

Why it exists: the arithmetic of the reference's hot path lives in the
third-party wheel ``dgl-cu101==0.5.3`` (/root/reference/requirements.txt:6),
which is absent from /root/reference and cannot be installed here.  This
package restates the *published semantics* of the calls the reference makes
 in plain torch-CPU so that the reference's own Python
files (cluster_gcn/modules.py, gcn/gcn.py, sampler.py, partition_utils.py,
cluster_gcn_ist_distrib.py) import and run unmodified on top of it
(``oracle/gen_golden.py`` puts this directory on sys.path as ``dgl``).

PARITY STATUS: **DGL semantics unpinned** (no DGL wheel, and the reference
has no tests / golden vectors); everything *above* DGL is pinned, because the
golden fixtures in tests/golden/ are produced by executing the reference's own
modules on top of this restatement.

Reference call sites restated here:
  update_all(copy_src, sum)  cluster_gcn/modules.py:136-137, :224-225; sampler.py:64-66
  in_degrees                 cluster_gcn/modules.py:156, :240; sampler.py:73
  subgraph                   cluster_gcn/sampler.py:34; partition_utils.py:23
  local_var / ndata / to / long / int   modules.py:219; cluster_gcn_ist_distrib.py:409, :504, :514
  GraphConv                  gcn/gcn.py:30-56; cluster_gcn/modules.py:331-338
  metis_partition            cluster_gcn/partition_utils.py:12 (assignment is an INPUT here)
"""
import numpy as np
import torch

from . import function  # noqa: F401
from . import backend  # noqa: F401

NID = '_ID'
EID = '_ID'


class DGLError(Exception):
    pass


class _Frame(dict):
    """ndata / edata container: a dict whose writes are row-count checked."""

    def __init__(self, n, *a, **k):
        super().__init__(*a, **k)
        self._n = n

    def __setitem__(self, key, val):
        if val.shape[0] != self._n:
            raise DGLError('Expect number of features to match number of nodes/edges '
                           '(len(u)). Got %d and %d instead.' % (val.shape[0], self._n))
        super().__setitem__(key, val)


class _EdgeBatch:
    def __init__(self, g):
        self.src = {k: v[g._src] for k, v in g.ndata.items()}
        self.dst = {k: v[g._dst] for k, v in g.ndata.items()}
        self.data = g.edata


class _NodeBatch:
    def __init__(self, mailbox, data):
        self.mailbox = mailbox
        self.data = data


class DGLGraph:
    """Homogeneous directed multigraph stored as a COO edge list (edge id =
    position).  A_in[v, u] = number of edges u -> v."""

    def __init__(self, graph_data=None, num_nodes=None, idtype=torch.int64):
        self._idtype = idtype
        if graph_data is None:
            src = dst = torch.zeros(0, dtype=torch.int64)
            n = num_nodes or 0
        elif hasattr(graph_data, 'number_of_nodes') and hasattr(graph_data, 'edges') \
                and not isinstance(graph_data, DGLGraph):
            # networkx graph (gcn/train.py:69, gcn/train_ist.py:115): undirected nx
            # graphs contribute both directions; edge order = nx iteration order.
            import networkx as nx
            g = graph_data
            n = g.number_of_nodes()
            if not g.is_directed():
                g = g.to_directed()
            e = np.asarray(list(g.edges()), dtype=np.int64).reshape(-1, 2)
            src, dst = torch.from_numpy(e[:, 0].copy()), torch.from_numpy(e[:, 1].copy())
        else:
            src, dst = graph_data
            src = torch.as_tensor(np.asarray(src), dtype=torch.int64) if not torch.is_tensor(src) else src.long()
            dst = torch.as_tensor(np.asarray(dst), dtype=torch.int64) if not torch.is_tensor(dst) else dst.long()
            n = num_nodes if num_nodes is not None else (int(max(src.max(), dst.max())) + 1 if len(src) else 0)
        self._src, self._dst, self._n = src, dst, int(n)
        self.ndata = _Frame(self._n)
        self.edata = _Frame(len(src))
        self._device = torch.device('cpu')

    # ---- structure queries -------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.shape[0])

    num_nodes = number_of_nodes
    num_edges = number_of_edges

    def edges(self):
        return self._src.to(self._idtype), self._dst.to(self._idtype)

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n).to(self._idtype).to(self._device)

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n).to(self._idtype).to(self._device)

    @property
    def idtype(self):
        return self._idtype

    @property
    def device(self):
        return self._device

    # ---- shallow copies ----------------------------------------------------
    def _clone(self, idtype=None):
        g = DGLGraph.__new__(DGLGraph)
        g._src, g._dst, g._n = self._src, self._dst, self._n
        g._idtype = idtype or self._idtype
        g._device = self._device
        g.ndata = _Frame(self._n, self.ndata)
        g.edata = _Frame(len(self._src), self.edata)
        return g

    def local_var(self):
        return self._clone()

    def long(self):
        return self._clone(torch.int64)

    def int(self):
        return self._clone(torch.int32)

    def to(self, device):
        g = self._clone()
        g._device = torch.device(device)
        for k, v in list(g.ndata.items()):
            dict.__setitem__(g.ndata, k, v.to(device))
        return g

    def cpu(self):
        return self.to('cpu')

    # ---- induced subgraph --------------------------------------------------
    def subgraph(self, nids):
        """Node-induced subgraph; new node i <-> nids[i] (order kept, no sort, no
        de-dup); keeps every edge with both endpoints selected, with
        multiplicity; ndata row-gathered; adds ndata[NID], edata[EID]."""
        nids = torch.as_tensor(np.asarray(nids) if not torch.is_tensor(nids) else nids).long().reshape(-1)
        relabel = torch.full((self._n,), -1, dtype=torch.int64)
        relabel[nids] = torch.arange(len(nids))
        keep = (relabel[self._src] >= 0) & (relabel[self._dst] >= 0)
        eid = torch.nonzero(keep).reshape(-1)
        sg = DGLGraph((relabel[self._src[eid]], relabel[self._dst[eid]]),
                      num_nodes=len(nids), idtype=self._idtype)
        for k, v in self.ndata.items():
            sg.ndata[k] = v[nids]
        for k, v in self.edata.items():
            sg.edata[k] = v[eid]
        sg.ndata[NID] = nids.to(self._idtype)
        sg.edata[EID] = eid.to(self._idtype)
        return sg

    # ---- message passing ---------------------------------------------------
    def apply_edges(self, func):
        out = func(_EdgeBatch(self))
        for k, v in out.items():
            self.edata[k] = v

    def update_all(self, message_func, reduce_func, apply_node_func=None):
        if isinstance(message_func, function.CopyMessage) and isinstance(reduce_func, function.SumReduce):
            assert message_func.out == reduce_func.msg
            h = self.ndata[message_func.src]
            out = torch.zeros((self._n,) + tuple(h.shape[1:]), dtype=h.dtype, device=h.device)
            out = out.index_add(0, self._dst.to(h.device), h[self._src.to(h.device)])
            self.ndata[reduce_func.out] = out
            return
        # generic UDF path (GAT, cluster_gcn/modules.py:57-65): degree bucketing
        msgs = message_func(_EdgeBatch(self))
        indeg = torch.bincount(self._dst, minlength=self._n)
        order = torch.argsort(self._dst, stable=True)
        results = {}
        for deg in torch.unique(indeg).tolist():
            if deg == 0:
                continue
            nodes = torch.nonzero(indeg == deg).reshape(-1)
            sel = torch.isin(self._dst[order], nodes)
            eids = order[sel].reshape(len(nodes), deg)  # dst sorted => grouped per node
            mailbox = {k: v[eids] for k, v in msgs.items()}
            out = reduce_func(_NodeBatch(mailbox, {k: v[nodes] for k, v in self.ndata.items()}))
            for k, v in out.items():
                if k not in results:
                    results[k] = torch.zeros((self._n,) + tuple(v.shape[1:]), dtype=v.dtype)
                results[k] = results[k].index_copy(0, nodes, v)
        for k, v in results.items():
            self.ndata[k] = v


def graph(data, num_nodes=None, idtype=torch.int64):
    return DGLGraph(data, num_nodes=num_nodes, idtype=idtype)


def from_scipy(sp_mat, idtype=torch.int64):
    """AmazonDataset.py:111 — pattern only, values dropped; edge (row -> col)."""
    coo = sp_mat.tocoo()
    return DGLGraph((coo.row.astype(np.int64), coo.col.astype(np.int64)),
                    num_nodes=coo.shape[0], idtype=idtype)


from . import transform  # noqa: E402,F401
from . import nn  # noqa: E402,F401
from . import data  # noqa: E402,F401
