"""GIST split / merge for the GraphConv-layout GCN (gcn/train_ist.py:148-286).

State-dict keys ``layers.{i}.weight`` ([in, out] layout) and ``layers.{i}.bias``.
Partitions are contiguous chunks of a torch-RNG permutation
(``torch.chunk(torch.randperm(n), m)``, train_ist.py:154-164) — a different
scheme from create_partition's round-robin deal.  Rows of W follow the previous
layer's partition, columns this layer's.  The 2-D slices run through the K5
CUDA kernels (``ops.slice_gather`` / ``ops.slice_scatter_``).

Faithful to the reference including its omissions: the last layer's bias is
never merged (train_ist.py:264-285 touches ``layers.{n_layers}.bias`` nowhere),
and tensors that are not split are averaged over the sub-models.
"""
import torch

from . import ops


def sample_feature_partitions(in_feats, n_hidden, n_layers, num_subnet, split_input, split_output):
    """train_ist.py:150-166; consumes torch's global RNG in the same order."""
    feats_idx = [torch.chunk(torch.randperm(in_feats), num_subnet) if split_input else None]
    for _ in range(1, n_layers):
        feats_idx.append(torch.chunk(torch.randperm(n_hidden), num_subnet))
    feats_idx.append(torch.chunk(torch.randperm(n_hidden), num_subnet) if split_output else None)
    return feats_idx


def _dev(idx, like):
    return None if idx is None else idx.to(like.device)


def split_state_dict(main, feats_idx, subnet_id, n_layers, split_input, split_output,
                     slice_ops=None):
    """Sub-model state dict of `subnet_id` (train_ist.py:176-191)."""
    gather, _ = slice_ops or (ops.slice_gather, ops.slice_scatter_)
    sub = dict(main)
    rows = {0: feats_idx[0][subnet_id] if split_input else None}
    cols = {}
    for i in range(1, n_layers + 1):
        if i == n_layers and not split_output:
            continue
        idx = feats_idx[i][subnet_id]
        cols[i - 1] = idx
        rows[i] = idx
    for l in range(n_layers + 1):
        r, c = rows.get(l), cols.get(l)
        if r is None and c is None:
            continue
        W = main['layers.%d.weight' % l]
        sub['layers.%d.weight' % l] = gather(W, _dev(r, W), _dev(c, W))
        if c is not None:
            b = main['layers.%d.bias' % l]
            sub['layers.%d.bias' % l] = gather(b, None, _dev(c, b))
    return sub


def merge_state_dicts(main, feats_idx, sub_dicts, n_layers, split_input, split_output,
                      slice_ops=None):
    """Merged full state dict (train_ist.py:241-285)."""
    _, scatter_ = slice_ops or (ops.slice_gather, ops.slice_scatter_)
    upd = {k: v.clone() for k, v in main.items()}
    m = len(sub_dicts)

    def mean(key):
        acc = sub_dicts[0][key].clone()
        for sd in sub_dicts[1:]:
            acc += sd[key]
        return acc / m

    def part_of(i):
        """feats_idx[i] partitions the inputs of layer i (= outputs of layer i-1)."""
        return feats_idx[i] if 0 <= i <= n_layers else None

    for l in range(n_layers + 1):
        wk, bk = 'layers.%d.weight' % l, 'layers.%d.bias' % l
        rpart = part_of(l)                                   # rows of W_l
        cpart = part_of(l + 1) if l < n_layers else None     # cols of W_l; classes never split
        if rpart is None and cpart is None:
            upd[wk] = mean(wk)
        else:
            for s, sd in enumerate(sub_dicts):
                W = upd[wk]
                scatter_(W, sd[wk], _dev(None if rpart is None else rpart[s], W),
                         _dev(None if cpart is None else cpart[s], W))
        if l < n_layers:
            if cpart is None:
                upd[bk] = mean(bk)          # train_ist.py:267 (outputs of layer L-1 not split)
            else:
                for s, sd in enumerate(sub_dicts):
                    b = upd[bk]
                    scatter_(b, sd[bk], None, _dev(cpart[s], b))
        # l == n_layers: the last bias is never merged by the reference
    return upd
