"""Drop-in modules for cluster_gcn/modules.py and DGL's GraphConv.

Same constructor signatures, parameter names / shapes / initialisers and
``forward`` signatures as the reference, so state dicts and the GIST
split/merge code (which re-assigns ``layer.linear.weight.data`` with tensors of
a different shape, cluster_gcn_ist_distrib.py:208-226) work unchanged.  The
aggregation, degree normalisation, concat, bias and ReLU run in the sm_100a
kernels of ``csrc/``; nothing here falls back to a CPU or library SpMM.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import GistError


# graph-cache key under which a batch may carry the first SAGE layer's prepared input (ops.SagePre)
SAGE_PRE0 = 'sage_pre0'


def _is_relu(act):
    return act is F.relu or act is torch.relu or isinstance(act, nn.ReLU)


class GraphConv(nn.Module):
    """dgl.nn.pytorch.GraphConv (0.5.x) with norm='both' semantics:
    weight is [in, out] (Xavier uniform), bias zeros; W is applied first iff in > out.
    Used by gcn/gcn.py:30-56 and cluster_gcn/modules.py:331-338."""

    def __init__(self, in_feats, out_feats, norm='both', weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm not in ('none', 'both', 'right'):
            raise GistError('Invalid norm value. Must be either "none", "both" or "right".')
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats)) if weight else None
        self.bias = nn.Parameter(torch.Tensor(out_feats)) if bias else None
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, feat):
        if not self._allow_zero_in_degree and graph.has_zero_in_degree():
            raise GistError('There are 0-in-degree nodes in the graph, output for those nodes will '
                            'be invalid. Add self-loops or set allow_zero_in_degree=True.')
        s = graph.rsqrt_out_degree() if self._norm == 'both' else None
        if self._norm == 'both':
            t = graph.rsqrt_in_degree()
        elif self._norm == 'right':
            t = graph.inv_in_degree()
        else:
            t = None
        fuse_relu = _is_relu(self._activation)
        # the width test uses the CURRENT weight shape: GIST re-assigns .data slices
        w = self.weight
        in_f, out_f = (w.shape if w is not None else (self._in_feats, self._out_feats))
        if in_f > out_f:
            # (diag(s) X) W == diag(s) (X W): the src-norm moves into the SpMM gather
            h = ops.matmul(feat, w) if w is not None else feat
            rst = ops.gspmm(graph, h, s, t, self.bias, fuse_relu)
            if self._activation is not None and not fuse_relu:
                rst = self._activation(rst)
            return rst
        # aggregate first; diag(t) commutes with the right-multiplication by W
        rst = ops.gspmm(graph, feat, s, t, None, False)
        if w is not None:
            if ops.get_matmul_precision() == 'fp32':
                rst = torch.addmm(self.bias, rst, w) if self.bias is not None else rst @ w
            else:
                rst = ops.matmul(rst, w)
                if self.bias is not None:
                    rst = rst + self.bias
        elif self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class GATLayer(nn.Module):
    """cluster_gcn/modules.py:24-65: one attention head.  fc: Linear(in, out, bias=False),
    attn_fc: Linear(2*out, 1, bias=False), both Xavier-normal with the ReLU gain.  The edge UDF
    (leaky_relu(attn_fc([z_u ‖ z_v]))), the mailbox softmax and the weighted sum run as one fused
    kernel (K6) instead of DGL's degree-bucketed Python UDFs."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.fc = nn.Linear(in_dim, out_dim, bias=False)
        self.attn_fc = nn.Linear(2 * out_dim, 1, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_fc.weight, gain=gain)

    def forward(self, g, h):
        z = ops.linear(h, self.fc.weight, None)
        return ops.gat_aggregate(g, z, self.attn_fc.weight, 0.01)     # F.leaky_relu default slope


PARALLEL_HEADS = True
# All heads of a layer as ONE projection GEMM + one K6 launch per kernel (heads = grid dimension) instead of
# a GEMM + three K6 launches per head on per-head streams.  OPT-IN (GIST_GAT_BATCH_HEADS=1): it halves the
# launches of the config-5 step (39 vs 76 per step) but measured SLOWER, 0.868 vs 0.799 ms/step
# (profiles/r2_bench_gat*.json): the per-head streams overlap the heads' tail-bound hub rows and their
# small GEMMs, one batched launch serialises the projection in front of the aggregation.
BATCH_HEADS = os.environ.get('GIST_GAT_BATCH_HEADS', '0') != '0'
_HEAD_STREAMS = {}


def _head_streams(device, k):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    pool = _HEAD_STREAMS.setdefault(idx, [])
    while len(pool) < k:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:k]


class MultiHeadGATLayer(nn.Module):
    """cluster_gcn/modules.py:67-76.  The committed forward returns
    ``torch.mean(torch.stack(head_outs))`` with no dim — a 0-d scalar, after which the next
    layer cannot run.  This implements the intended mean over the HEAD axis
    (``dim=0``); pass ``reduce='scalar'`` for the literal behaviour of the committed code."""

    def __init__(self, in_dim, out_dim, num_heads, reduce='heads'):
        super().__init__()
        assert reduce in ('heads', 'scalar')
        self.reduce = reduce
        self.heads = nn.ModuleList()
        for _ in range(num_heads):
            self.heads.append(GATLayer(in_dim, out_dim))

    def _batched(self, g, h):
        """The heads' fc weights stacked into one [H*D, in] projection, their attention vectors into
        [H, 2D]; autograd splits the gradients back to the per-head parameters."""
        H = len(self.heads)
        W_all = torch.cat([hd.fc.weight for hd in self.heads], dim=0)
        attn_all = torch.stack([hd.attn_fc.weight.reshape(-1) for hd in self.heads])
        z_all = ops.linear(h, W_all, None)
        out_all = ops.gat_aggregate_heads(g, z_all, attn_all, H, 0.01)
        if self.reduce == 'scalar':
            return torch.mean(out_all)
        D = out_all.shape[1] // H
        return out_all.reshape(out_all.shape[0], H, D).mean(dim=1)

    def forward(self, g, h):
        if (h.is_cuda and len(self.heads) > 1 and BATCH_HEADS
                and len({tuple(hd.fc.weight.shape) for hd in self.heads}) == 1):
            return self._batched(g, h)
        if h.is_cuda and len(self.heads) > 1 and PARALLEL_HEADS:
            # the heads share only read-only inputs (graph, h): one stream per head, so their
            # projection GEMMs and row-per-warp attention kernels (tail-bound on a cluster batch's hub
            # rows) overlap — parallel branches under CUDA-graph capture; autograd runs each head's
            # backward on its forward stream.  fork: side waits for main; join: main waits for side.
            main = torch.cuda.current_stream(h.device)
            streams = _head_streams(h.device, len(self.heads))
            head_outs = []
            for attn_head, st in zip(self.heads, streams):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    head_outs.append(attn_head(g, h))
            for st in streams:
                main.wait_stream(st)
        else:
            head_outs = [attn_head(g, h) for attn_head in self.heads]
        if self.reduce == 'scalar':
            return torch.mean(torch.stack(head_outs))
        if len(head_outs) == 1:
            return head_outs[0]
        return torch.mean(torch.stack(head_outs), dim=0)


class GAT(nn.Module):
    """cluster_gcn/modules.py:78-98: num_layers MultiHeadGATLayers (the last with one head),
    ELU after every layer including the last."""

    def __init__(self, num_layers, in_dim, hidden_dim, out_dim, num_heads):
        super().__init__()
        layers = [MultiHeadGATLayer(in_dim, hidden_dim, num_heads)]
        for _ in range(num_layers - 2):
            layers.append(MultiHeadGATLayer(hidden_dim, hidden_dim, num_heads))
        layers.append(MultiHeadGATLayer(hidden_dim, out_dim, 1))
        self.layers = nn.ModuleList(layers)

    def forward(self, g):
        h = g.ndata['feat']
        for layer in self.layers:
            h = layer(g, h)
            h = F.elu(h)
        return h


class GraphSAGELayer(nn.Module):
    """cluster_gcn/modules.py:100-159."""

    def __init__(self, in_feats, out_feats, activation, dropout, bias=True, use_pp=False,
                 use_lynorm=True):
        super().__init__()
        self.linear = nn.Linear(2 * in_feats, out_feats, bias=bias)
        self.activation = activation
        self.use_pp = use_pp
        self.dropout = nn.Dropout(p=dropout) if dropout else 0.
        self.lynorm = nn.LayerNorm(out_feats, elementwise_affine=True) if use_lynorm else (lambda x: x)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.linear.weight.size(1))
        self.linear.weight.data.uniform_(-stdv, stdv)
        if self.linear.bias is not None:
            self.linear.bias.data.uniform_(-stdv, stdv)

    def forward(self, g, h):
        if not self.use_pp or not self.training:
            h = ops.sage_concat(g, h)       # [h ‖ (A h) / in_deg] in one kernel
        if self.dropout:
            h = self.dropout(h)
        h = ops.linear(h, self.linear.weight, self.linear.bias)
        h = self.lynorm(h)
        if self.activation:
            h = self.activation(h)
        return h

    def concat(self, h, ah, norm):
        return torch.cat((h, ah * norm), dim=1)

    def get_norm(self, g):
        return g.inv_in_degree().unsqueeze(1)


class ISTSAGELayer(nn.Module):
    """cluster_gcn/modules.py:191-243: Linear(2*in, out), U(+-1/sqrt(2*in)) init for
    weight and bias, LayerNorm(out, affine=False) when use_lynorm."""

    def __init__(self, in_feats, out_feats, dropout, use_lynorm, activation=None):
        super().__init__()
        self.linear = nn.Linear(2 * in_feats, out_feats)
        self.activation = activation
        self.init_layer()
        self.dropout = nn.Dropout(p=dropout) if dropout else 0.
        self.lynorm = nn.LayerNorm(out_feats, elementwise_affine=False) if use_lynorm else (lambda x: x)
        self._drop_stream = ops.new_dropout_stream()      # this layer's slot in the fused dropout clock

    def init_layer(self):
        stdv = 1. / math.sqrt(self.linear.weight.size(1))
        self.linear.weight.data.uniform_(-stdv, stdv)
        self.linear.bias.data.uniform_(-stdv, stdv)

    def _p_drop(self):
        return float(self.dropout.p) if (self.dropout and self.training) else 0.0

    def prepare_input(self, g, h, out=None, balanced=True, background=0):
        """z = dropout([h ‖ (A h) / in_deg]) (+ its 3xTF32 low half) for this layer, outside
        autograd — the pipelined trainer runs it for the NEXT batch's input features while the
        current batch trains (they do not depend on the weights and need no gradient)."""
        return ops.sage_prepare(g, h, self._p_drop(), self._drop_stream, out=out, balanced=balanced,
                                background=background)

    def forward(self, g, h, pre=None):
        # pre: this layer's prepared input for (g, h) (prepare_input), if already computed
        W = self.linear.weight
        if (pre is None and not self.training and not torch.is_grad_enabled() and h.is_cuda
                and W.shape[0] < h.shape[1] and self.linear.bias is not None):
            # inference (evaluate(), utils.py:70-80): project first, aggregate at the output width
            h = ops.sage_project_first(g, h, W, self.linear.bias)
        elif h.is_cuda and ops.get_matmul_precision() != 'fp32':
            # tensor-core path: aggregation + concat + dropout (+ split) in K1's epilogue, the
            # dropout backward in the dz GEMM's epilogue, layer norm + ReLU in the projection's
            # epilogue when a tile holds the row (ops._SageLinear: the whole layer is one autograd node)
            ln = None
            if isinstance(self.lynorm, nn.LayerNorm) and (self.activation is None or _is_relu(self.activation)):
                ln = (self.lynorm.eps, self.activation is not None)
            h = ops.sage_linear(g, h, self.linear.weight, self.linear.bias, self._p_drop(), self._drop_stream,
                                pre=pre, ln=ln)
            if ln is not None:
                return h
        else:
            if pre is not None:
                h = pre.z
                if self.dropout and not pre.dropped:
                    h = self.dropout(h)
            else:
                h = ops.sage_concat(g, h)       # get_norm + update_all + `*norm` + cat fused
                if self.dropout:
                    h = self.dropout(h)
            # GIST swaps in weight slices of other widths; use the live tensors, and
            # normalise over the live output width
            h = ops.linear(h, self.linear.weight, self.linear.bias)
        if isinstance(self.lynorm, nn.LayerNorm):
            # LayerNorm over the live output width + ReLU in one kernel (fwd) / one (bwd)
            fuse = self.activation is None or _is_relu(self.activation)
            h = ops.layer_norm_act(h, self.lynorm.eps, relu=fuse and self.activation is not None)
            if self.activation and not fuse:
                h = self.activation(h)
            return h
        h = self.lynorm(h)
        if self.activation:
            h = self.activation(h)
        return h

    def get_norm(self, g):
        return g.inv_in_degree().unsqueeze(1)


class GraphSAGE(nn.Module):
    """cluster_gcn/modules.py:161-189."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout, use_pp):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(GraphSAGELayer(in_feats, n_hidden, activation=activation, dropout=dropout,
                                          use_pp=use_pp, use_lynorm=True))
        for _ in range(n_layers - 1):
            self.layers.append(GraphSAGELayer(n_hidden, n_hidden, activation=activation,
                                              dropout=dropout, use_pp=False, use_lynorm=True))
        self.layers.append(GraphSAGELayer(n_hidden, n_classes, activation=None, dropout=dropout,
                                          use_pp=False, use_lynorm=False))

    def forward(self, g):
        h = g.ndata['feat']
        for layer in self.layers:
            h = layer(g, h)
        return h


def _split_widths(in_feats, n_hidden, n_classes, n_layers, split_input, split_output, num_subnet):
    """(in, out, is_output) per layer — the width bookkeeping shared by both GCN
    containers (cluster_gcn/modules.py:260-308, gcn/gcn.py:27-56)."""
    k = num_subnet
    hk = int(n_hidden // k)
    first_in = int(in_feats // k) if split_input else in_feats
    first_out = n_hidden if (n_layers <= 1 and not split_output) else hk
    dims = [(first_in, first_out)]
    for i in range(n_layers - 1):
        dims.append((hk, n_hidden if (i == n_layers - 2 and not split_output) else hk))
    dims.append((hk if split_output else n_hidden, n_classes))
    return dims


class GCN(nn.Module):
    """The GraphSAGE-style IST container, cluster_gcn/modules.py:245-314.
    State-dict keys ``layers.{i}.linear.{weight,bias}``."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout,
                 use_layernorm=True, split_input=False, split_output=False, num_subnet=1,
                 use_aggregation=False):
        super().__init__()
        self.layers = nn.ModuleList()
        self.use_layernorm = use_layernorm
        self.split_input = split_input
        self.split_output = split_output
        if not use_aggregation:
            raise NotImplementedError('You must use graph sage')
        dims = _split_widths(in_feats, n_hidden, n_classes, n_layers, split_input, split_output,
                             num_subnet)
        for fin, fout in dims[:-1]:
            self.layers.append(ISTSAGELayer(fin, fout, dropout, use_layernorm, activation=activation))
        fin, fout = dims[-1]
        self.layers.append(ISTSAGELayer(fin, fout, dropout, False, activation=None))

    def forward(self, g):
        h = g.ndata['feat']
        pre0 = g._cache.get(SAGE_PRE0)      # layer-0 input prepared ahead of time (graphed.py)
        if self.training and h.is_cuda and ops.get_matmul_precision() != 'fp32':
            st = ops.dropout_state(h.device)
            if st.auto_tick:
                st.tick()                   # one dropout step per training forward
        if (h.is_cuda and ops.get_matmul_precision() == '3xtf32'
                and (torch.is_grad_enabled() or self.training)):
            # every layer's weight low half in one launch at the head of the step, instead of one
            # small launch in front of each layer's GEMM (consumed by ops._weight_lo)
            ops.presplit_weights([layer.linear.weight for layer in self.layers])
        try:
            for i, layer in enumerate(self.layers):
                h = layer(g, h, pre=pre0) if (i == 0 and pre0 is not None) else layer(g, h)
        finally:
            ops.clear_presplit()            # the low halves are valid for this forward only
        return h


class BaselineGCN(nn.Module):
    """cluster_gcn/modules.py:316-349: GraphConv stack + whole-tensor layer norm."""

    def __init__(self, in_feats, n_hidden, n_classes, n_layers, activation, dropout,
                 use_layernorm=True):
        super().__init__()
        self.layers = nn.ModuleList()
        self.use_layernorm = use_layernorm
        self.layers.append(GraphConv(in_feats, n_hidden, activation=activation))
        for _ in range(n_layers - 1):
            self.layers.append(GraphConv(n_hidden, n_hidden, activation=activation))
        self.layers.append(GraphConv(n_hidden, n_classes))
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, g):
        h = g.ndata['feat']
        for i, layer in enumerate(self.layers):
            if i != 0:
                h = self.dropout(h)
            h = layer(g, h)
            if i < len(self.layers) - 1 and self.use_layernorm:
                h = ops.tensor_layer_norm(h)    # F.layer_norm(h, h.shape): ONE mean/var over all n*d elements
        return h
