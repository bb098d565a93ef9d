"""CPU, world_size 2 over gloo: DistributedGATWrapper (GIST for the GAT model, config 5) against
a direct restatement of the reference's index algebra
(cluster_gcn/cluster_gcn_ist_distrib_gat.py:99-150, :204-236).  The reference wrapper itself
cannot run as committed, so there is no golden from it."""
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expected_after_sync(base0, subs, parts, L, heads_per_layer):
    """Apply the reference's assignments with plain torch indexing, site by site."""
    base = {k: v.clone() for k, v in base0.items()}
    m = len(subs)
    for site in range(m):
        sd = subs[site]
        for l in range(L + 1):
            for h in range(heads_per_layer[l]):
                fk, ak = 'layers.%d.heads.%d.fc.weight' % (l, h), 'layers.%d.heads.%d.attn_fc.weight' % (l, h)
                if l == 0:
                    idx, full = parts[0][site]
                    base[fk][idx, :] = sd[fk]                               # …gat.py:111-112
                    base[ak][:, full] = sd[ak]                              # :113-114
                elif l == L:
                    idx, _ = parts[L - 1][site]
                    base[fk][:, idx] = sd[fk]                               # :119-120
                else:
                    prev, _ = parts[l - 1][site]
                    nxt, full_next = parts[l][site]
                    rows = base[fk][:, prev]
                    rows[nxt, :] = sd[fk]                                   # :129-132
                    base[fk][:, prev] = rows
                    base[ak][:, full_next] = sd[ak]                         # :135-136
    for h in range(heads_per_layer[L]):                                     # all_reduce mean, :99-102
        ak = 'layers.%d.heads.%d.attn_fc.weight' % (L, h)
        acc = subs[0][ak].clone()
        for s in subs[1:]:
            acc = acc + s[ak]
        base[ak] = acc / m
    return base


def _worker(rank, m, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        from gist_b200.ist_gat import DistributedGATWrapper
        from tests.test_host_logic import cpu_slice_ops
        L, heads, hid, fin, ncls, seed = 2, 2, 8, 5, 3, 7
        torch.manual_seed(seed)
        np.random.seed(seed)
        random.seed(seed)
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=m)
        args = SimpleNamespace(rank=rank, num_subnet=m, n_hidden=hid, n_layers=L, n_heads=heads)
        w = DistributedGATWrapper(args, None, fin, ncls, torch.device('cpu'), slice_ops=cpu_slice_ops())
        base0 = {k: v.clone() for k, v in w.base_model.state_dict().items()}
        w.ini_sync_dispatch_model()
        parts = w.current_partition
        errs = []
        # dispatch: sub-model == slices of the replica
        sd = w.sub_model.state_dict()
        idx0, full0 = parts[0][rank]
        if not torch.equal(sd['layers.0.heads.1.fc.weight'], base0['layers.0.heads.1.fc.weight'][idx0, :]):
            errs.append('dispatch fc0')
        if not torch.equal(sd['layers.0.heads.0.attn_fc.weight'], base0['layers.0.heads.0.attn_fc.weight'][:, full0]):
            errs.append('dispatch attn0')
        prev, _ = parts[0][rank]
        nxt, fnext = parts[1][rank]
        if not torch.equal(sd['layers.1.heads.0.fc.weight'], base0['layers.1.heads.0.fc.weight'][:, prev][nxt, :]):
            errs.append('dispatch fc1')
        if not torch.equal(sd['layers.2.heads.0.fc.weight'], base0['layers.2.heads.0.fc.weight'][:, parts[1][rank][0]]):
            errs.append('dispatch fcL')
        if tuple(sd['layers.2.heads.0.attn_fc.weight'].shape) != (1, 2 * ncls):
            errs.append('shared attn shape')
        # "train": deterministic per-rank perturbation of every sub tensor
        with torch.no_grad():
            for i, p in enumerate(w.sub_model.parameters()):
                p.data = p.data * 1.5 + 0.01 * (rank + 1) * (i + 1)
        mine = {k: v.clone() for k, v in w.sub_model.state_dict().items()}
        gathered = [None] * m
        dist.all_gather_object(gathered, mine)
        w.sync_model()
        exp = _expected_after_sync(base0, gathered, parts, L, [heads, heads, 1])
        for k, v in w.base_model.state_dict().items():
            if not torch.allclose(v, exp[k], rtol=0, atol=1e-7):
                errs.append('sync ' + k)
        q.put((rank, errs))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, ['EXC ' + traceback.format_exc()]))


def test_gat_wrapper_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    m = 2
    procs = [ctx.Process(target=_worker, args=(r, m, 29731, q)) for r in range(m)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(m))
    for p in procs:
        p.join(60)
    assert all(len(v) == 0 for v in res.values()), res
