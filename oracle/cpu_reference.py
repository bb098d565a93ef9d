"""CPU stand-in for "the reference's DGL CPU path" used by bench.py's cpu_baseline and
--impl reference legs.  TEST / MEASUREMENT INFRASTRUCTURE ONLY — never imported by
gist_b200.

DGL 0.5.3 cannot be installed here (oracle/README.md), so the reference's CPU path is
timed as this restatement: the same loop as cluster_gcn_ist_distrib.py:398-417 —
CPU induced-subgraph extraction per step (partition_utils.py:20-25; DGL slices the
CSR per selected node, restated with scipy's CSR row/column slicing), feature/label
row gather, ISTSAGELayer stack (modules.py:218-237) with the SpMM done by torch's
multi-threaded CSR `torch.sparse.mm` (DGL's CPU SpMM is an OpenMP CSR loop), autograd
backward, Adam.  kind = "port".
"""
import time

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, At, x):
        ctx.At = At
        return torch.sparse.mm(A, x)

    @staticmethod
    def backward(ctx, dy):
        return None, None, torch.sparse.mm(ctx.At, dy)


def _to_torch_csr(m):
    m = m.tocsr()
    return torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)),
                                   torch.from_numpy(m.indices.astype(np.int64)),
                                   torch.ones(m.nnz, dtype=torch.float32), size=m.shape)


class CpuClusterTrainer:
    """Holds the training graph (in-CSR as scipy) + features on the host."""

    def __init__(self, rowptr, col, feat, label, in_feats, n_hidden, n_classes, n_layers,
                 num_subnet=1, dropout=0.2, use_layernorm=True, lr=1e-2, weight_decay=5e-4, seed=0):
        n = rowptr.shape[0] - 1
        self.A = sp.csr_matrix((np.ones(col.shape[0], dtype=np.float32), col, rowptr), shape=(n, n))
        self.feat, self.label = feat, label
        torch.manual_seed(seed)
        hk = n_hidden // num_subnet
        dims = [(in_feats, hk)] + [(hk, hk)] * (n_layers - 1) + [(hk, n_classes)]
        self.params = []
        for fin, fout in dims:
            lin = torch.nn.Linear(2 * fin, fout)
            stdv = 1. / (2 * fin) ** 0.5
            lin.weight.data.uniform_(-stdv, stdv)
            lin.bias.data.uniform_(-stdv, stdv)
            self.params += [lin.weight, lin.bias]
        self.layers = [(self.params[2 * i], self.params[2 * i + 1]) for i in range(len(dims))]
        self.dropout, self.use_layernorm = dropout, use_layernorm
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay)

    def step(self, nids):
        """One reference training step on the batch whose node ids (into the training
        graph) are `nids` (int64)."""
        sub = self.A[nids][:, nids].tocsr()              # CPU induced subgraph, per step
        A = _to_torch_csr(sub)
        At = _to_torch_csr(sub.T)
        deg = torch.from_numpy(np.diff(sub.indptr).astype(np.float32)).unsqueeze(1)
        norm = 1. / deg
        norm[torch.isinf(norm)] = 0
        t = torch.from_numpy(nids)
        h = self.feat[t]
        y = self.label[t]
        self.opt.zero_grad()
        L = len(self.layers)
        for l, (w, b) in enumerate(self.layers):
            ah = _SpMM.apply(A, At, h) * norm
            z = torch.cat((h, ah), dim=1)
            z = F.dropout(z, self.dropout, True)
            h = F.linear(z, w, b)
            if l < L - 1:
                if self.use_layernorm:
                    h = F.layer_norm(h, (h.shape[-1],))
                h = F.relu(h)
        loss = F.cross_entropy(h, y)
        loss.backward()
        self.opt.step()
        return float(loss.detach())


def time_steps(trainer, batches, warmup=1):
    """Seconds per step over `batches` (list of int64 node-id arrays) after `warmup`."""
    for b in batches[:warmup]:
        trainer.step(b)
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        trainer.step(b)
    return (time.perf_counter() - t0) / max(len(batches) - warmup, 1)


def time_full_graph_spmm(rowptr, col, n, d, row_sample=None, seed=0):
    """Seconds for Y = A·X on the host (torch CSR, all threads).  If row_sample is given,
    only that many leading rows of A are multiplied (bounded sample) and the time is
    scaled by nnz_total / nnz_sample."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    rows = n if row_sample is None else min(row_sample, n)
    nnz = int(rowptr[rows])
    A = torch.sparse_csr_tensor(torch.from_numpy(rowptr[:rows + 1].copy()), torch.from_numpy(col[:nnz].copy()),
                                torch.ones(nnz, dtype=torch.float32), size=(rows, n))
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    torch.sparse.mm(A, x[:, :8].contiguous())
    t0 = time.perf_counter()
    torch.sparse.mm(A, x)
    dt = time.perf_counter() - t0
    return dt * (float(rowptr[n]) / max(nnz, 1)), nnz


# ---------------------------------------------------------------------------------------------
# Full-graph GraphConv trainers (configs 1 and 2): gcn/train.py:86-121 and gcn/train_ist.py:140-286
# with DGL's GraphConv(norm='both') restated as one torch CSR product with the symmetric
# normalisation folded into the edge values, whole-tensor layer norm
# (gcn/gcn.py:65-66), dropout before every layer but the first (:62-63).
# ---------------------------------------------------------------------------------------------
def _normalised_adj(rowptr, col, n):
    indeg = np.diff(rowptr).astype(np.float64)
    outdeg = np.bincount(col, minlength=n).astype(np.float64)
    t = np.maximum(indeg, 1.0) ** -0.5
    s = np.maximum(outdeg, 1.0) ** -0.5
    row = np.repeat(np.arange(n), np.diff(rowptr))
    vals = (t[row] * s[col]).astype(np.float32)
    A = sp.csr_matrix((vals, col, rowptr), shape=(n, n))

    def tcsr(m):
        m = m.tocsr()
        return torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)),
                                       torch.from_numpy(m.indices.astype(np.int64)),
                                       torch.from_numpy(m.data.astype(np.float32)), size=m.shape)
    return tcsr(A), tcsr(A.T)


def _graphconv_forward(A, At, h, params, dropout, use_layernorm, training):
    L = len(params)
    for i, (W, b) in enumerate(params):
        if i != 0:
            h = F.dropout(h, dropout, training)
        if W.shape[0] > W.shape[1]:
            h = _SpMM.apply(A, At, h @ W) + b
        else:
            h = _SpMM.apply(A, At, h) @ W + b
        if i < L - 1:
            h = F.relu(h)
            if use_layernorm:
                h = F.layer_norm(h, h.shape)
    return h


def _xavier(fin, fout):
    W = torch.empty(fin, fout)
    torch.nn.init.xavier_uniform_(W)
    return W.requires_grad_(True), torch.zeros(fout, requires_grad=True)


class CpuGCNTrainer:
    """gcn/train.py: one Adam step on the full graph per epoch."""

    def __init__(self, rowptr, col, feat, label, train_mask, n_hidden, n_classes, n_layers, dropout=0.5,
                 use_layernorm=True, lr=1e-3, weight_decay=5e-4, seed=0):
        n = rowptr.shape[0] - 1
        self.A, self.At = _normalised_adj(np.asarray(rowptr, np.int64), np.asarray(col, np.int64), n)
        self.feat, self.label, self.mask = feat, label, train_mask.bool()
        torch.manual_seed(seed)
        dims = [(feat.shape[1], n_hidden)] + [(n_hidden, n_hidden)] * (n_layers - 1) + [(n_hidden, n_classes)]
        self.params = [_xavier(a, b) for a, b in dims]
        self.dropout, self.use_layernorm = dropout, use_layernorm
        self.opt = torch.optim.Adam([t for p in self.params for t in p], lr=lr, weight_decay=weight_decay)

    def epoch(self, e=0):
        self.opt.zero_grad()
        out = _graphconv_forward(self.A, self.At, self.feat, self.params, self.dropout, self.use_layernorm, True)
        loss = F.cross_entropy(out[self.mask], self.label[self.mask])
        loss.backward()
        self.opt.step()
        return float(loss.detach())


class CpuISTGCNTrainer:
    """gcn/train_ist.py: m sub-GCNs trained one after the other every epoch, split / merged every
    iter_per_site epochs (split_input False, split_output True: the config-2 flags of
    script/sweep.py:12-13; index algebra from oracle/gist_oracle.py)."""

    def __init__(self, rowptr, col, feat, label, train_mask, n_hidden, n_classes, n_layers, num_subnet,
                 iter_per_site=5, dropout=0.5, use_layernorm=True, lr=1e-2, weight_decay=5e-4, seed=0):
        from oracle import gist_oracle as O
        self.O = O
        n = rowptr.shape[0] - 1
        self.A, self.At = _normalised_adj(np.asarray(rowptr, np.int64), np.asarray(col, np.int64), n)
        self.feat, self.label, self.mask = feat, label, train_mask.bool()
        torch.manual_seed(seed)
        self.L, self.m, self.h, self.ips = n_layers, num_subnet, n_hidden, iter_per_site
        dims = [(feat.shape[1], n_hidden)] + [(n_hidden, n_hidden)] * (n_layers - 1) + [(n_hidden, n_classes)]
        self.main = {}
        for l, (a, b) in enumerate(dims):
            W, bb = _xavier(a, b)
            self.main['layers.%d.weight' % l], self.main['layers.%d.bias' % l] = W.detach(), bb.detach()
        self.dropout, self.use_layernorm, self.lr, self.wd = dropout, use_layernorm, lr, weight_decay

    def epoch(self, e):
        O, L, m = self.O, self.L, self.m
        if e % self.ips == 0:
            self.feats_idx = [None] + [torch.chunk(torch.randperm(self.h), m) for _ in range(1, L)] + \
                             [torch.chunk(torch.randperm(self.h), m)]
            self.subs, self.opts = [], []
            for s in range(m):
                sd = O.graphconv_split(self.main, self.feats_idx, s, L, False, True)
                ps = [(sd['layers.%d.weight' % l].clone().requires_grad_(True),
                       sd['layers.%d.bias' % l].clone().requires_grad_(True)) for l in range(L + 1)]
                self.subs.append(ps)
                self.opts.append(torch.optim.Adam([t for p in ps for t in p], lr=self.lr, weight_decay=self.wd))
        loss = None
        for s in range(m):
            self.opts[s].zero_grad()
            out = _graphconv_forward(self.A, self.At, self.feat, self.subs[s], self.dropout, self.use_layernorm, True)
            loss = F.cross_entropy(out[self.mask], self.label[self.mask])
            loss.backward()
            self.opts[s].step()
        if (e + 1) % self.ips == 0:
            sds = [{('layers.%d.%s' % (l, nm)): t.detach() for l, p in enumerate(ps) for nm, t in zip(('weight', 'bias'), p)}
                   for ps in self.subs]
            self.main = O.graphconv_merge(self.main, self.feats_idx, sds, L, False, True)
        return float(loss.detach())


def time_epochs(trainer, n_epochs, warmup=1):
    """Seconds per epoch of a Cpu*GCNTrainer after `warmup` epochs."""
    for e in range(warmup):
        trainer.epoch(e)
    t0 = time.perf_counter()
    for e in range(warmup, warmup + n_epochs):
        trainer.epoch(e)
    return (time.perf_counter() - t0) / max(n_epochs, 1)
