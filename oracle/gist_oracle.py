"""Functional CPU restatement (torch, fp32 or fp64) of the reference hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it follows.  The DGL-level semantics (copy_src/sum, in_degrees,
subgraph, GraphConv) follow SURVEY.md Appendix A because the DGL 0.5.3 wheel is
absent: **DGL semantics parity unpinned**.  Everything above DGL is pinned by
tests/golden/*.npz, which oracle/gen_golden.py produced by running the
reference's own modules (from /root/reference) on top of oracle/dgl_restated.

All functions are differentiable through torch autograd, so gradients of the
CUDA path are checked against autograd of this restatement.
"""
import random

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ graph ---
class OGraph:
    """COO multigraph: edge e is src[e] -> dst[e]; A_in[v,u] = #edges u->v."""

    def __init__(self, src, dst, n):
        self.src = torch.as_tensor(src).long().reshape(-1).cpu()
        self.dst = torch.as_tensor(dst).long().reshape(-1).cpu()
        self.n = int(n)

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.n)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.n)

    def canonical_csr(self):
        """in-CSR with columns ascending inside each row (the comparison form)."""
        key = self.dst * self.n + self.src
        key, _ = torch.sort(key)
        rowptr = torch.zeros(self.n + 1, dtype=torch.int64)
        rowptr[1:] = torch.cumsum(torch.bincount(self.dst, minlength=self.n), 0)
        return rowptr, key % self.n

    def subgraph(self, nids):
        """DGLGraph.subgraph (App. A): node i <-> nids[i], order kept; every edge with
        both endpoints selected survives (with multiplicity)."""
        nids = torch.as_tensor(np.asarray(nids)).long().reshape(-1)
        relabel = torch.full((self.n,), -1, dtype=torch.int64)
        relabel[nids] = torch.arange(len(nids))
        keep = (relabel[self.src] >= 0) & (relabel[self.dst] >= 0)
        return OGraph(relabel[self.src[keep]], relabel[self.dst[keep]], len(nids))


def copy_src_sum(g, h, chunk=1 << 22):
    """update_all(fn.copy_src, fn.sum) (modules.py:224-225): out[v] = sum_{u->v} h[u]."""
    out = torch.zeros((g.n,) + tuple(h.shape[1:]), dtype=h.dtype)
    E = g.src.shape[0]
    if E <= chunk:
        return out.index_add(0, g.dst, h[g.src])
    for s in range(0, E, chunk):
        out = out.index_add(0, g.dst[s:s + chunk], h[g.src[s:s + chunk]])
    return out


def sage_norm(g, dtype=torch.float32):
    """get_norm (modules.py:239-243): 1/in_degree as float, inf -> 0, shape [n,1]."""
    norm = 1. / g.in_degrees().to(dtype).unsqueeze(1)
    norm[torch.isinf(norm)] = 0
    return norm


# ----------------------------------------------------------------- layers ---
def ist_sage_layer(g, h, weight, bias, use_lynorm, activation=None, dropout_mask=None):
    """ISTSAGELayer.forward (modules.py:218-237).  dropout_mask, if given, is the
    already-scaled multiplicative mask nn.Dropout would have applied to [h ‖ ah]."""
    norm = sage_norm(g, h.dtype)
    ah = copy_src_sum(g, h) * norm
    z = torch.cat((h, ah), dim=1)
    if dropout_mask is not None:
        z = z * dropout_mask
    y = F.linear(z, weight, bias)
    if use_lynorm:
        y = F.layer_norm(y, (y.shape[-1],), None, None, 1e-5)
    if activation is not None:
        y = activation(y)
    return y


def sage_gcn_forward(g, feat, params, use_layernorm, activation=F.relu, dropout_masks=None):
    """GCN.forward of cluster_gcn/modules.py:310-314 over ISTSAGELayers: every layer but
    the last has (layer norm if use_layernorm) + activation; the last has neither
    (modules.py:299-308).  params = [(W_l, b_l)]."""
    h = feat
    L = len(params)
    for l, (w, b) in enumerate(params):
        last = l == L - 1
        h = ist_sage_layer(g, h, w, b, use_layernorm and not last, None if last else activation,
                           None if dropout_masks is None else dropout_masks[l])
    return h


def graphsage_layer(g, h, weight, bias, ln_weight=None, ln_bias=None, use_lynorm=True, activation=None,
                    aggregate=True, dropout_mask=None):
    """GraphSAGELayer.forward (modules.py:131-148): as ISTSAGELayer but the layer norm is AFFINE
    (nn.LayerNorm(out, elementwise_affine=True), :125), the bias is optional (:112) and with
    ``use_pp`` in training mode the aggregation + concat are skipped (``aggregate=False``, :133):
    the input is then already [h ‖ ah] wide."""
    if aggregate:
        norm = sage_norm(g, h.dtype)
        h = torch.cat((h, copy_src_sum(g, h) * norm), dim=1)
    if dropout_mask is not None:
        h = h * dropout_mask
    y = F.linear(h, weight, bias)
    if use_lynorm:
        y = F.layer_norm(y, (y.shape[-1],), ln_weight, ln_bias, 1e-5)
    if activation is not None:
        y = activation(y)
    return y


def graphsage_forward(g, feat, params, activation=F.relu):
    """GraphSAGE.forward (modules.py:161-189), eval mode / use_pp False: every layer but the last has
    the affine layer norm + activation, the last has neither.
    params = [(W, b, ln_weight | None, ln_bias | None)]."""
    h = feat
    L = len(params)
    for l, (w, b, lw, lb) in enumerate(params):
        last = l == L - 1
        h = graphsage_layer(g, h, w, b, lw, lb, not last, None if last else activation)
    return h


def graph_conv(g, feat, weight, bias, activation=None, norm='both'):
    """DGL 0.5.x GraphConv.forward (App. A): src-norm on the input, W first iff
    in > out (strict), dst-norm, bias, activation.  weight is [in, out]."""
    if norm == 'both':
        degs = g.out_degrees().to(feat.dtype).clamp(min=1)
        feat = feat * torch.pow(degs, -0.5).unsqueeze(1)
    if weight.shape[0] > weight.shape[1]:
        rst = copy_src_sum(g, feat @ weight)
    else:
        rst = copy_src_sum(g, feat) @ weight
    if norm != 'none':
        degs = g.in_degrees().to(feat.dtype).clamp(min=1)
        nrm = torch.pow(degs, -0.5) if norm == 'both' else 1.0 / degs
        rst = rst * nrm.unsqueeze(1)
    if bias is not None:
        rst = rst + bias
    if activation is not None:
        rst = activation(rst)
    return rst


def graphconv_gcn_forward(g, feat, params, use_layernorm, activation=F.relu, dropout_masks=None):
    """gcn/gcn.py:59-67 (also BaselineGCN, modules.py:341-349): dropout before every
    layer but the first, whole-tensor layer norm after every layer but the last."""
    h = feat
    L = len(params)
    for i, (w, b) in enumerate(params):
        if i != 0 and dropout_masks is not None and dropout_masks[i] is not None:
            h = h * dropout_masks[i]
        h = graph_conv(g, h, w, b, activation if i < L - 1 else None)
        if i < L - 1 and use_layernorm:
            h = F.layer_norm(h, h.shape)
    return h


# -------------------------------------------------------------------- GAT ---
def gat_layer(g, h, fc_weight, attn_weight, negative_slope=0.01):
    """GATLayer.forward (modules.py:57-65): z = fc(h); e_uv = leaky_relu(attn_fc([z_u ‖ z_v]))
    (:40-44, F.leaky_relu default slope 0.01); alpha = softmax over the in-edges of v (:53);
    h_v = sum_u alpha_uv z_u (:55).  Nodes without in-edges receive no message and keep the
    zero row DGL's degree bucketing leaves (DGL-recall).  Multi-edges count once each."""
    z = F.linear(h, fc_weight)
    D = z.shape[1]
    a = attn_weight.reshape(-1)
    el, er = z @ a[:D], z @ a[D:]
    e = F.leaky_relu(el[g.src] + er[g.dst], negative_slope)
    m = torch.full((g.n,), -float('inf'), dtype=e.dtype).scatter_reduce(0, g.dst, e.detach(), 'amax',
                                                                         include_self=True)
    w = torch.exp(e - m[g.dst])
    den = torch.zeros(g.n, dtype=e.dtype).index_add(0, g.dst, w)
    alpha = w / den[g.dst]
    return torch.zeros((g.n, D), dtype=z.dtype).index_add(0, g.dst, alpha.unsqueeze(1) * z[g.src])


def multi_head_gat_layer(g, h, heads, negative_slope=0.01):
    """MultiHeadGATLayer.forward (modules.py:74-76) with the INTENDED semantics: the mean over
    the head axis.  The committed code calls torch.mean(torch.stack(head_outs)) without a dim,
    which collapses everything to a scalar and makes the next layer fail."""
    return torch.stack([gat_layer(g, h, fw, aw, negative_slope) for fw, aw in heads]).mean(dim=0)


def gat_forward(g, feat, layers, negative_slope=0.01):
    """GAT.forward (modules.py:93-98): ELU after every layer, including the last.
    layers = [[(fc_weight, attn_weight) per head] per layer]."""
    h = feat
    for heads in layers:
        h = F.elu(multi_head_gat_layer(g, h, heads, negative_slope))
    return h


# ------------------------------------------------------------ cluster iter ---
def batch_node_ids(par_li, i, psize, batch_size):
    """partition_utils.py:20-25."""
    sel = [par_li[s] for s in range(i * batch_size, (i + 1) * batch_size) if s < psize]
    return np.concatenate(sel).reshape(-1).astype(np.int64)


# -------------------------------------------------------------- GIST algebra ---
def create_partition(num_subnet, size):
    """cluster_gcn_ist_distrib.py:51-65, stated position-wise: shuffled position i
    goes to sub-network i % m.  Consumes Python's global `random` stream."""
    perm = list(range(size))
    random.shuffle(perm)
    buckets = [[] for _ in range(num_subnet)]
    for pos, feat in enumerate(perm):
        buckets[pos % num_subnet].append(feat)
    out = []
    for b in buckets:
        idx = torch.tensor(b, dtype=torch.int64)
        out.append((idx, torch.cat((idx, idx + size))))
    return out


def sage_dispatch(base_params, parts, site):
    """Sub-model parameters of `site` cut from the full model
    (cluster_gcn_ist_distrib.py:204-226 / :291-313).  base_params = [(W,b)] with W in
    nn.Linear layout [out, 2*in]; parts[l][site] = (idx, full_idx)."""
    L = len(parts)
    sub = []
    for l, (W, b) in enumerate(base_params):
        if l == 0:
            idx, _ = parts[0][site]
            sub.append((W[idx, :].clone(), b[idx].clone()))
        elif l == L:
            _, full = parts[L - 1][site]
            sub.append((W[:, full].clone(), b.clone()))
        else:
            _, full_prev = parts[l - 1][site]
            nxt, _ = parts[l][site]
            sub.append((W[:, full_prev][nxt, :].clone(), b[nxt].clone()))
    return sub


def sage_sync(base_params, parts, subs):
    """Inverse scatter of every site's trained slices into the full model and mean of
    the shared last bias (cluster_gcn_ist_distrib.py:100-195).  Returns new base."""
    L = len(parts)
    m = len(subs)
    base = [(W.clone(), b.clone()) for (W, b) in base_params]
    for site in range(m):
        for l in range(L + 1):
            W, b = base[l]
            sw, sb = subs[site][l]
            if l == 0:
                idx, _ = parts[0][site]
                W[idx, :] = sw
                b[idx] = sb
            elif l == L:
                _, full = parts[L - 1][site]
                W[:, full] = sw
            else:
                _, full_prev = parts[l - 1][site]
                nxt, _ = parts[l][site]
                cols = W[:, full_prev]
                cols[nxt, :] = sw
                W[:, full_prev] = cols
                b[nxt] = sb
    base[L] = (base[L][0], sum(subs[s][L][1] for s in range(m)) / m)
    return base


def graphconv_split(main, feats_idx, subnet_id, n_layers, split_input, split_output):
    """gcn/train_ist.py:176-191 — GraphConv layout W[in, out]; main is a dict
    'layers.{i}.weight' / 'layers.{i}.bias'."""
    sub = dict(main)
    if split_input:
        idx = feats_idx[0][subnet_id]
        sub['layers.0.weight'] = main['layers.0.weight'][idx, :]
    for i in range(1, n_layers + 1):
        if i == n_layers and not split_output:
            continue
        idx = feats_idx[i][subnet_id]
        sub['layers.%d.weight' % (i - 1)] = sub['layers.%d.weight' % (i - 1)][:, idx]
        sub['layers.%d.bias' % (i - 1)] = main['layers.%d.bias' % (i - 1)][idx]
        sub['layers.%d.weight' % i] = main['layers.%d.weight' % i][idx, :]
    return sub


def graphconv_merge(main, feats_idx, sub_dicts, n_layers, split_input, split_output):
    """gcn/train_ist.py:241-285; returns the updated full state dict.  Faithful to the
    reference including what it does NOT do (last-layer bias is never merged; in
    the reference update_dict aliases main_dict's tensors, so writes are in place)."""
    upd = {k: v.clone() for k, v in main.items()}
    m = len(sub_dicts)
    w0 = 'layers.0.weight'
    if split_input:
        if n_layers <= 1 and not split_output:
            for idx, sd in zip(feats_idx[0], sub_dicts):
                upd[w0][idx, :] = sd[w0]
        else:
            for i, sd in enumerate(sub_dicts):
                cur, nxt = feats_idx[0][i], feats_idx[1][i]
                rows = upd[w0][cur, :]
                rows[:, nxt] = sd[w0]
                upd[w0][cur, :] = rows
    else:
        if n_layers <= 1 and not split_output:
            upd[w0] = sum(sd[w0] for sd in sub_dicts) / m
        else:
            for i, sd in enumerate(sub_dicts):
                upd[w0][:, feats_idx[1][i]] = sd[w0]
    for i in range(1, n_layers + 1):
        bk, wk = 'layers.%d.bias' % (i - 1), 'layers.%d.weight' % i
        if i == n_layers:
            if not split_output:
                upd[bk] = sum(sd[bk] for sd in sub_dicts) / m
                upd[wk] = sum(sd[wk] for sd in sub_dicts) / m
            else:
                for idx, sd in zip(feats_idx[i], sub_dicts):
                    upd[bk][idx] = sd[bk]
                    upd[wk][idx, :] = sd[wk]
        else:
            if i >= n_layers - 1 and not split_output:
                for idx, sd in zip(feats_idx[i], sub_dicts):
                    upd[bk][idx] = sd[bk]
                    upd[wk][idx, :] = sd[wk]
            else:
                for s, sd in enumerate(sub_dicts):
                    cur, nxt = feats_idx[i][s], feats_idx[i + 1][s]
                    upd[bk][cur] = sd[bk]
                    rows = upd[wk][cur, :]
                    rows[:, nxt] = sd[wk]
                    upd[wk][cur, :] = rows
    return upd
