"""GIST (independent sub-network training) partition / dispatch / sync.

Mirrors cluster_gcn/cluster_gcn_ist_distrib.py:51-367 (and its byte-identical
copy in cluster_gcn_ist_ultra_wide.py).  What changed is the data movement:

  reference                                   here (8x B200, NVSwitch)
  ---------                                   -----------------------
  rank 0 is a parameter server holding the    every rank holds a replica of the
  only full model (on the CPU in ultra-wide)  full model in HBM (fits: 180 GB)
  dispatch = (2(L+1)-1)(m-1) two-rank         dispatch = local K5 slice gather,
  broadcasts, a new communicator per tensor   no communication at all
  sync = the same number of broadcasts back   sync = ONE all-gather of a packed
  + all_reduce of the last bias               buffer (all slices + last bias),
                                              then a local K5 scatter of every
                                              rank's slice into the replica

The partition is never communicated in either design: every rank draws it from
the same Python ``random`` stream (same seed, same call order).
"""
import os
import random

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .modules import GCN


def create_partition(num_subnet, size):
    """cluster_gcn_ist_distrib.py:51-65: shuffle range(size) with Python's
    ``random``; sub-network j owns the shuffled positions j, j+m, j+2m, ...
    Returns [(idx, full_idx)] with full_idx = cat(idx, idx + size) (the same
    columns in the `ah` half of the SAGE concat)."""
    perm = list(range(size))
    random.shuffle(perm)
    out = []
    for j in range(num_subnet):
        idx = torch.tensor(perm[j::num_subnet], dtype=torch.int64)
        out.append((idx, torch.cat((idx, idx + size))))
    return out


def _dist_ready():
    return dist.is_available() and dist.is_initialized()


class DistributedGNNWrapper(torch.nn.Module):
    """Full ("base") model replica + this rank's sub-model.

    args needs: rank, num_subnet, n_hidden, n_layers, dropout, use_layernorm."""

    def __init__(self, args, g, in_feats, n_classes, device, slice_ops=None, base_init='cpu'):
        """base_init: 'cpu' (default) builds the full model on the host and moves it, exactly as the
        reference does (…distrib.py:81-83: same seed -> bit-identical initial parameters);
        'device' draws rank 0's initial values with the device generator instead (same
        distributions) — for the ultra-wide configurations, where the host-side uniform_() of an
        8.6 GB layer dominates set-up.  Replicas on ranks > 0 never draw values: they are
        allocated on the device and filled by the broadcast from rank 0."""
        super().__init__()
        assert base_init in ('cpu', 'device')
        # set to a list to record CUDA-event timestamps of every sync / dispatch (bench.py)
        self.profile = None
        # (gather, scatter_) executors for the 2-D slices; the product path is the CUDA
        # K5 kernels.  Tests of the host-side plan / pack / all-gather logic on CPU
        # (gloo) inject a checker here; nothing in the package does.
        self._gather, self._scatter_ = slice_ops or (ops.slice_gather, ops.slice_scatter_)
        self._cuda_slices = slice_ops is None       # product path: all slices of a round in ONE launch (ops.slice_multi)
        self._symm = None                           # peer-memory sync state (see _peer_buffers)
        # True: dispatch writes the slices INTO the sub-model's existing parameter storage
        # (same values; keeps addresses stable for CUDA-graph replay) instead of rebinding
        # `.data` to fresh tensors as the reference does.
        self.inplace_dispatch = False
        self.args = args
        self.g = g
        self.in_feats = in_feats
        self.n_classes = n_classes
        self.device = device
        mk_base = lambda: GCN(in_feats, args.n_hidden, n_classes, args.n_layers, F.relu,  # noqa: E731
                              args.dropout, args.use_layernorm, False, False, 1, True)
        if args.rank == 0 and base_init == 'cpu':
            self.base_model = mk_base().to(device)
        elif args.rank == 0:
            with torch.device(device):
                self.base_model = mk_base()
        else:
            # replica: allocated on the device without drawing anything from this rank's RNG
            # streams (the reference builds nothing here); filled from rank 0 below
            with torch.device('meta'):
                self.base_model = mk_base()
            self.base_model = self.base_model.to_empty(device=device)
        self.sub_model = GCN(in_feats, args.n_hidden, n_classes, args.n_layers, F.relu,
                             args.dropout, args.use_layernorm, False, True, args.num_subnet,
                             True).to(device)
        self.current_partition = None
        if _dist_ready() and args.num_subnet > 1:
            for p in self.base_model.parameters():      # in place, tensor by tensor: no packed 8.6 GB copy
                dist.broadcast(p.data, src=0)
        elif args.rank != 0:
            raise RuntimeError('rank %d of a %d-way GIST job needs an initialised process group '
                               '(the full-model replica is filled from rank 0)' % (args.rank, args.num_subnet))

    # ---------------------------------------------------------- partition --
    def sample_partitions(self):
        return [create_partition(self.args.num_subnet, self.args.n_hidden)
                for _ in range(self.args.n_layers)]

    def _to_dev(self, parts):
        """Index tensors of a partition on the device: ONE packed host->device copy from a persistent
        pinned buffer (2 * L * m small synchronous copies stall the step pipeline at every round
        boundary), then views."""
        dev = torch.device(self.device)
        if dev.type != 'cuda':
            return [[(i.to(dev), f.to(dev)) for (i, f) in layer] for layer in parts]
        flat = [t for layer in parts for pair in layer for t in pair]
        total = sum(t.numel() for t in flat)
        if getattr(self, '_idx_pin', None) is None or self._idx_pin.numel() < total:
            self._idx_pin = torch.empty(total, dtype=torch.int64).pin_memory()
            self._idx_pin_ev = None
        if self._idx_pin_ev is not None:
            self._idx_pin_ev.synchronize()           # the previous round's copy has left the buffer
        off = 0
        for t in flat:
            self._idx_pin[off:off + t.numel()].copy_(t)
            off += t.numel()
        d = self._idx_pin[:total].to(dev, non_blocking=True)
        self._idx_pin_ev = torch.cuda.Event()
        self._idx_pin_ev.record()
        out, off = [], 0
        for layer in parts:
            row = []
            for (i, f) in layer:
                row.append((d[off:off + i.numel()], d[off + i.numel():off + i.numel() + f.numel()]))
                off += i.numel() + f.numel()
            out.append(row)
        return out

    # ------------------------------------------------------------ dispatch --
    def _slice_for(self, layer_idx, site, parts):
        """(row index, col index) of base layer `layer_idx` owned by `site`."""
        L = len(parts)
        if layer_idx == 0:
            return parts[0][site][0], None
        if layer_idx == L:
            return None, parts[L - 1][site][1]
        return parts[layer_idx][site][0], parts[layer_idx - 1][site][1]

    def _dispatch_local(self, parts):
        site = self.args.rank
        L = len(parts)
        with torch.no_grad():
            if self._cuda_slices:
                # product path: every slice of the round in one launch (K5 multi-gather)
                jobs = []
                for l in range(L + 1):
                    ridx, cidx = self._slice_for(l, site, parts)
                    base = self.base_model.layers[l].linear
                    sub = self.sub_model.layers[l].linear
                    wr = ridx.shape[0] if ridx is not None else base.weight.shape[0]
                    wc = cidx.shape[0] if cidx is not None else base.weight.shape[1]
                    if not (self.inplace_dispatch and tuple(sub.weight.shape) == (wr, wc)):
                        sub.weight.data = torch.empty((wr, wc), dtype=torch.float32, device=base.weight.device)
                    jobs.append((base.weight.data, ridx, cidx, sub.weight.data))
                    if l == L:
                        if self.inplace_dispatch:
                            sub.bias.data.copy_(base.bias.data)
                        else:
                            sub.bias.data = base.bias.data.clone()      # shared, full (…distrib.py:217-219)
                    else:
                        if not (self.inplace_dispatch and tuple(sub.bias.shape) == (wr,)):
                            sub.bias.data = torch.empty(wr, dtype=torch.float32, device=base.bias.device)
                        jobs.append((base.bias.data, None, ridx, sub.bias.data))
                ops.slice_multi(jobs, scatter=False)
                return
            for l in range(L + 1):
                ridx, cidx = self._slice_for(l, site, parts)
                base = self.base_model.layers[l].linear
                sub = self.sub_model.layers[l].linear
                if self.inplace_dispatch:
                    self._gather(base.weight.data, ridx, cidx, out=sub.weight.data)
                    if l == L:
                        sub.bias.data.copy_(base.bias.data)
                    else:
                        self._gather(base.bias.data, None, ridx, out=sub.bias.data)
                    continue
                sub.weight.data = self._gather(base.weight.data, ridx, cidx)
                if l == L:
                    sub.bias.data = base.bias.data.clone()      # shared, full (…distrib.py:217-219)
                else:
                    sub.bias.data = self._gather(base.bias.data, None, ridx)

    def ini_sync_dispatch_model(self):
        parts = self._to_dev(self.sample_partitions())
        self._dispatch_local(parts)
        self.current_partition = parts

    def dispatch_model(self):
        ev = self._stamp()
        parts = self._to_dev(self.sample_partitions())
        self._dispatch_local(parts)
        self.current_partition = parts
        self._stamp(ev, 'dispatch')

    # ------------------------------------------------------------- timing --
    def _stamp(self, ev=None, what=None, **extra):
        """Event pair around a round-boundary operation on the current stream (profile mode only)."""
        if self.profile is None or not torch.cuda.is_available():
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if ev is None:
            return e
        self.profile.append(dict(what=what, ev0=ev, ev1=e, **extra))
        return e

    # ---------------------------------------------------------------- sync --
    def _pack(self, out=None):
        ts = [t.data.reshape(-1) for lyr in self.sub_model.layers for t in (lyr.linear.weight, lyr.linear.bias)]
        return torch.cat(ts, out=out) if out is not None else torch.cat(ts)

    def _peer_buffers(self, numel):
        """Peer-memory sync — EXPERIMENTAL, opt-in with GIST_SYNC=peer (default: the NCCL all-gather path).
        Every rank packs its trained slices into a SYMMETRIC buffer (torch.distributed._symmetric_memory)
        that all ranks of the node map over NVLink; the merge kernel then reads each site's slices straight
        from that site's HBM while it scatters them into the local replica — all-gather + scatter as ONE
        kernel, no gathered staging buffer.  Status (round 2, 2 x B200): bit-exact against the reference
        golden (tests/test_gpu_dist_nccl.py, sync_mode='peer') and 0.27 ms per Reddit-shape round, but the
        ultra-wide configuration (139 MB per rank) did not complete in 25 minutes on the one attempt made,
        so it is not the default.  Returns (handle, [one flat view per rank]) or None (-> all-gather)."""
        m = self.args.num_subnet
        if (m <= 1 or not _dist_ready() or not self._cuda_slices or torch.device(self.device).type != 'cuda'
                or os.environ.get('GIST_SYNC', 'allgather') != 'peer'):
            return None
        if self._symm is not None and self._symm[0] == numel:
            return self._symm[1], self._symm[2]
        if self._symm is False:
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            if hasattr(symm_mem, 'enable_symm_mem_for_group'):
                try:
                    symm_mem.enable_symm_mem_for_group(dist.group.WORLD.group_name)
                except Exception:
                    pass
            buf = symm_mem.empty(numel, dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            views = [buf if r == self.args.rank else hdl.get_buffer(r, (numel,), torch.float32) for r in range(m)]
            self._symm = (numel, hdl, views)
            return hdl, views
        except Exception as ex:          # no NVLink peer mapping on this box: stay on the collective
            self._symm = False
            self._symm_error = repr(ex)[:300]
            return None

    def sync_model(self):
        m = self.args.num_subnet
        parts = self.current_partition
        L = len(parts)
        with torch.no_grad():
            ev = self._stamp()
            numel = sum(t.numel() for lyr in self.sub_model.layers for t in (lyr.linear.weight, lyr.linear.bias))
            peer = self._peer_buffers(numel)
            if peer is not None:
                hdl, views = peer
                self._pack(out=views[self.args.rank])
                ev = self._stamp(ev, 'sync_pack', bytes=numel * 4) or ev
                hdl.barrier(channel=0)                       # every site's slices are in its buffer
                self._merge(views, parts)                    # reads peers over NVLink, scatters locally
                hdl.barrier(channel=1)                       # nobody re-packs while a peer still reads
                self._stamp(ev, 'sync_peer_merge', bytes_sent=numel * 4, bytes_received=(m - 1) * numel * 4)
            else:
                flat = self._pack()
                ev = self._stamp(ev, 'sync_pack', bytes=flat.numel() * 4) or ev
                if _dist_ready() and m > 1:
                    out = getattr(self, '_gather_buf', None)     # persistent: a round must not pay a cudaMalloc
                    if out is None or out.numel() != m * flat.numel() or out.device != flat.device:
                        out = self._gather_buf = torch.empty(m * flat.numel(), dtype=flat.dtype, device=flat.device)
                    dist.all_gather_into_tensor(out, flat)       # ONE collective for all slices
                    gathered = out.view(m, flat.numel())
                else:
                    assert m == 1, 'num_subnet > 1 needs an initialised process group'
                    gathered = flat.unsqueeze(0)
                ev = self._stamp(ev, 'sync_all_gather', bytes_sent=flat.numel() * 4,
                                 bytes_received=(m - 1) * flat.numel() * 4) or ev
                self._merge(gathered, parts)
                self._stamp(ev, 'sync_scatter', bytes=m * flat.numel() * 4)
            # the reference all-reduces the last bias IN PLACE on every rank's sub-model
            if self.inplace_dispatch:
                self.sub_model.layers[L].linear.bias.data.copy_(self.base_model.layers[L].linear.bias.data)
            else:
                self.sub_model.layers[L].linear.bias.data = self.base_model.layers[L].linear.bias.data.clone()

    @staticmethod
    def _stream_merge(dst, w, ridx, cidx):
        """Row-streaming merge for this slice?  GIST_MERGE=rows / scatter forces it on (where it applies) / off;
        default: slices of >= 2^20 elements landing on >= 1/16 of the columns of rows of >= 32 KB — the sites'
        row sets are disjoint (create_partition splits a permutation), which the kernel relies on."""
        mode = os.environ.get('GIST_MERGE', 'auto')
        if mode == 'scatter' or not ops.slice_scatter_rows_ok(dst, ridx, cidx):
            return False
        if mode == 'rows':
            return True
        return w.numel() >= (1 << 20) and dst.shape[1] >= 8192 and 16 * cidx.numel() >= dst.shape[1]

    def _merge(self, gathered, parts):
        """Scatter every site's packed slices into the local full-model replica.  ``gathered``: [m, numel]
        tensor or a list of m flat views (peer buffers)."""
        m = len(gathered)
        L = len(parts)
        shapes = [(tuple(lyr.linear.weight.shape), tuple(lyr.linear.bias.shape))
                  for lyr in self.sub_model.layers]
        last_bias = None
        jobs, row_jobs = [], []
        for site in range(m):
            off = 0
            row = gathered[site]
            for l in range(L + 1):
                (wr, wc), (bn,) = shapes[l]
                w = row[off:off + wr * wc].view(wr, wc); off += wr * wc
                b = row[off:off + bn]; off += bn
                ridx, cidx = self._slice_for(l, site, parts)
                base = self.base_model.layers[l].linear
                if self._cuda_slices and self._stream_merge(base.weight.data, w, ridx, cidx):
                    # ultra-wide layer: whole sectors of the site's rows in ascending order (one launch for all
                    # sites) instead of one 32-byte sector per scattered element
                    inv = ops.index_invert(cidx, base.weight.shape[1])
                    row_jobs.append((w, ridx, inv, base.weight.data))
                elif self._cuda_slices:
                    jobs.append((w, ridx, cidx, base.weight.data))
                else:
                    self._scatter_(base.weight.data, w, ridx, cidx)
                if l == L:
                    last_bias = b.clone() if last_bias is None else last_bias + b   # rank order
                elif self._cuda_slices:
                    jobs.append((b, None, ridx, base.bias.data))
                else:
                    self._scatter_(base.bias.data, b, None, ridx)
        if jobs:
            ops.slice_multi(jobs, scatter=True)          # all sites x all tensors: one launch
        if row_jobs:
            ops.slice_scatter_rows_(row_jobs)
        self.base_model.layers[L].linear.bias.data = last_bias / m
