"""CPU, world_size 2 and 4 over gloo: the N>1 path of DistributedGNNWrapper (partition
draws, local dispatch, packed all-gather, merge) against the golden recorded from the
reference's DistributedGNNWrapper running on real gloo processes.  The CUDA K5 slice
kernels are replaced by an injected torch-indexing checker (tests/test_host_logic.py)."""
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _worker(rank, m, port, ci, q, backend='gloo', sync_mode=None):
    try:
        if sync_mode is not None:
            os.environ['GIST_SYNC'] = sync_mode          # 'peer' (NVLink peer memory) | 'allgather' (NCCL)
        _worker_body(rank, m, port, ci, q, backend)
    except Exception as e:      # surface the failure instead of hanging the parent
        import traceback
        q.put((rank, ['EXC ' + traceback.format_exc()]))


def _worker_body(rank, m, port, ci, q, backend='gloo'):
    """backend 'gloo': CPU tensors, injected slice checker.  backend 'nccl' (tests/test_gpu_dist_nccl.py):
    one GPU per rank, the real K5 slice kernels and the NCCL all-gather."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gist_b200.ist import DistributedGNNWrapper
    from tests.test_host_logic import cpu_slice_ops
    G = np.load(os.path.join(GOLD, 'wrapper.npz'))
    p = 'w%d_' % ci
    m_, fin, hid, ncls, L, seed = (int(v) for v in G[p + 'cfg'])
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    if backend == 'nccl':
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=m,
                                device_id=dev)
        slice_ops = None
    else:
        dev = torch.device('cpu')
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=m)
        slice_ops = cpu_slice_ops()
    args = SimpleNamespace(rank=rank, num_subnet=m, n_hidden=hid, n_layers=L, dropout=0.0, use_layernorm=True)
    w = DistributedGNNWrapper(args, None, fin, ncls, dev, slice_ops=slice_ops)
    errs = []

    def eq(name, got, exact=True):
        ref = G[p + name]
        got = got.detach().cpu().numpy()
        ok = np.array_equal(got, ref) if exact else np.allclose(got, ref, rtol=1e-6, atol=1e-7)
        if not ok:
            errs.append(name)

    # every rank holds the full-model replica, equal to the reference's rank-0 init
    for l, lyr in enumerate(w.base_model.layers):
        eq('r0_base0.layers.%d.linear.weight' % l, lyr.linear.weight)
        eq('r0_base0.layers.%d.linear.bias' % l, lyr.linear.bias)
    w.ini_sync_dispatch_model()
    for l in range(L):
        eq('r%d_part0.%d' % (rank, l), torch.stack([q_[0] for q_ in w.current_partition[l]]))
    for l, lyr in enumerate(w.sub_model.layers):
        eq('r%d_sub0.layers.%d.linear.weight' % (rank, l), lyr.linear.weight)
        eq('r%d_sub0.layers.%d.linear.bias' % (rank, l), lyr.linear.bias)
    with torch.no_grad():
        for li, lyr in enumerate(w.sub_model.layers):
            lyr.linear.weight.data = lyr.linear.weight.data * 1.25 + 0.01 * (rank + 1) * (li + 1)
            lyr.linear.bias.data = lyr.linear.bias.data - 0.125 * (rank + 1)
    dist.barrier()
    w.sync_model()
    if backend == 'nccl' and os.environ.get('GIST_SYNC') == 'peer' and not w._symm:
        errs.append('peer-memory path was requested but not used: %s' % getattr(w, '_symm_error', '?'))
    eq('r%d_lastbias_after_sync' % rank, w.sub_model.layers[-1].linear.bias, exact=False)
    for l, lyr in enumerate(w.base_model.layers):        # EVERY rank's replica == reference rank 0
        eq('r0_base1.layers.%d.linear.weight' % l, lyr.linear.weight)
        eq('r0_base1.layers.%d.linear.bias' % l, lyr.linear.bias, exact=(l < L))
    dist.barrier()
    w.dispatch_model()
    for l in range(L):
        eq('r%d_part1.%d' % (rank, l), torch.stack([q_[0] for q_ in w.current_partition[l]]))
    for l, lyr in enumerate(w.sub_model.layers):
        eq('r%d_sub1.layers.%d.linear.weight' % (rank, l), lyr.linear.weight)
        eq('r%d_sub1.layers.%d.linear.bias' % (rank, l), lyr.linear.bias, exact=(l < L))
    q.put((rank, errs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('ci,m', [(0, 2), (1, 4)])
def test_wrapper_over_gloo_matches_reference_processes(ci, m):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29710 + ci
    procs = [ctx.Process(target=_worker, args=(r, m, port, ci, q)) for r in range(m)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(m))
    for p in procs:
        p.join(60)
    assert all(len(v) == 0 for v in res.values()), res
