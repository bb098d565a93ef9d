"""GIST partition / dispatch / sync for the GAT model (config 5).

Index algebra of cluster_gcn/cluster_gcn_ist_distrib_gat.py:96-391, per attention head:

  layer 0        fc.weight[idx0, :]                   attn_fc.weight[:, full_idx0]      (:214-221)
  layer 0<l<L    fc.weight[:, prev_idx][next_idx, :]  attn_fc.weight[:, full_next]      (:229-236)
  layer L (out)  fc.weight[:, idx_{L-1}]              attn_fc.weight whole — shared by every
                                                      sub-model, averaged at sync        (:222-228, :99-102)

with (idx, full_idx) = create_partition(m, n_hidden), full_idx = cat(idx, idx + n_hidden) (the
a_l and a_r halves of the attention vector).  The reference file cannot run as committed
(``self.ags`` typo at :293/:384; it indexes ``n_heads`` heads in the one-head output layer;
``GAT(n_layers, …)`` builds n_layers layers while the wrapper walks n_layers + 1 — SURVEY.md
§2.4), so this module implements the algebra above for the heads that exist and builds
``GAT(n_layers + 1, …)``, which coincides with the reference for the only depth its sweep uses
(n_layers = 1, script/reddit/run_gat_distrib_sweep.py:11-15).

Data movement is the B200 design of gist_b200/ist.py: every rank keeps a full-model replica in
HBM, dispatch is a local K5 slice gather, sync is ONE packed all-gather + local K5 scatters.
"""
import torch
import torch.distributed as dist

from . import ops
from .ist import create_partition, _dist_ready
from .modules import GAT


class DistributedGATWrapper(torch.nn.Module):
    """args needs: rank, num_subnet, n_hidden, n_layers, n_heads."""

    def __init__(self, args, g, in_feats, n_classes, device, slice_ops=None):
        super().__init__()
        self._gather, self._scatter_ = slice_ops or (ops.slice_gather, ops.slice_scatter_)
        self.inplace_dispatch = False
        self.args, self.g, self.in_feats, self.n_classes, self.device = args, g, in_feats, n_classes, device
        assert args.n_hidden % args.num_subnet == 0
        mk = lambda hid: GAT(args.n_layers + 1, in_feats, hid, n_classes, args.n_heads)   # noqa: E731
        if args.rank == 0:
            self.base_model = mk(args.n_hidden).to(device)
        else:
            with torch.random.fork_rng(devices=[]):     # the reference builds nothing here
                self.base_model = mk(args.n_hidden).to(device)
        self.sub_model = mk(args.n_hidden // args.num_subnet).to(device)
        self.current_partition = None
        if _dist_ready() and args.num_subnet > 1:
            flat = torch.cat([p.data.reshape(-1) for p in self.base_model.parameters()])
            dist.broadcast(flat, src=0)
            off = 0
            for p in self.base_model.parameters():
                p.data.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()

    def sample_partitions(self):
        return [create_partition(self.args.num_subnet, self.args.n_hidden)
                for _ in range(self.args.n_layers)]

    def _to_dev(self, parts):
        return [[(i.to(self.device), f.to(self.device)) for (i, f) in layer] for layer in parts]

    def _plan(self, parts, site):
        """[(base tensor, sub tensor, row idx, col idx, shared)] for every parameter, in the fixed
        order used for packing."""
        L = len(parts)
        plan = []
        for l in range(L + 1):
            for hb, hs in zip(self.base_model.layers[l].heads, self.sub_model.layers[l].heads):
                if l == 0:
                    idx, full = parts[0][site]
                    plan.append((hb.fc.weight, hs.fc.weight, idx, None, False))
                    plan.append((hb.attn_fc.weight, hs.attn_fc.weight, None, full, False))
                elif l == L:
                    idx, _ = parts[L - 1][site]
                    plan.append((hb.fc.weight, hs.fc.weight, None, idx, False))
                    plan.append((hb.attn_fc.weight, hs.attn_fc.weight, None, None, True))
                else:
                    prev, _ = parts[l - 1][site]
                    nxt, full_next = parts[l][site]
                    plan.append((hb.fc.weight, hs.fc.weight, nxt, prev, False))
                    plan.append((hb.attn_fc.weight, hs.attn_fc.weight, None, full_next, False))
        return plan

    def _dispatch_local(self, parts):
        with torch.no_grad():
            for base, sub, ridx, cidx, shared in self._plan(parts, self.args.rank):
                if shared:
                    if self.inplace_dispatch:
                        sub.data.copy_(base.data)
                    else:
                        sub.data = base.data.clone()
                elif self.inplace_dispatch:
                    self._gather(base.data, ridx, cidx, out=sub.data)
                else:
                    sub.data = self._gather(base.data, ridx, cidx)

    def ini_sync_dispatch_model(self):
        parts = self._to_dev(self.sample_partitions())
        self._dispatch_local(parts)
        self.current_partition = parts

    dispatch_model = ini_sync_dispatch_model

    def sync_model(self):
        m = self.args.num_subnet
        parts = self.current_partition
        with torch.no_grad():
            mine = self._plan(parts, self.args.rank)
            flat = torch.cat([sub.data.reshape(-1) for _, sub, _, _, _ in mine])
            if _dist_ready() and m > 1:
                out = torch.empty(m * flat.numel(), dtype=flat.dtype, device=flat.device)
                dist.all_gather_into_tensor(out, flat)          # ONE collective for all slices
                gathered = out.view(m, flat.numel())
            else:
                assert m == 1, 'num_subnet > 1 needs an initialised process group'
                gathered = flat.unsqueeze(0)
            shared_acc = {}
            for site in range(m):
                off = 0
                for k, (base, sub, ridx, cidx, shared) in enumerate(self._plan(parts, site)):
                    numel = sub.numel()
                    piece = gathered[site, off:off + numel].view(sub.shape)
                    off += numel
                    if shared:      # averaged over the sub-models, in rank order (bit-reproducible)
                        shared_acc[k] = piece.clone() if k not in shared_acc else shared_acc[k] + piece
                    else:
                        self._scatter_(base.data, piece, ridx, cidx)
            for k, (base, sub, _, _, shared) in enumerate(mine):
                if shared:
                    base.data = shared_acc[k] / m
                    if self.inplace_dispatch:
                        sub.data.copy_(base.data)
                    else:
                        sub.data = base.data.clone()
