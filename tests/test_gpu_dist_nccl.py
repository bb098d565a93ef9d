"""GPU, world_size 2 (and 4 when the box has them) over NCCL: DistributedGNNWrapper end to end on
real devices — partition draws, local K5 dispatch, then either the peer-memory merge (one kernel reading
every site's slices over NVLink) or ONE packed NCCL all-gather + local K5 merge — against
the golden recorded from the reference's own DistributedGNNWrapper running on real processes
(tests/golden/wrapper.npz).  Every rank's full-model replica must equal the reference's rank-0
model bit for bit after the sync.  Skipped on boxes with fewer devices than ranks."""
import pytest
import torch

from tests.test_dist_gloo import _worker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('sync_mode', ['peer', 'allgather'])
@pytest.mark.parametrize('ci,m', [(0, 2), (1, 4)])
def test_wrapper_over_nccl_matches_reference_processes(ci, m, sync_mode):
    if torch.cuda.device_count() < m:
        pytest.skip('needs %d CUDA devices, found %d' % (m, torch.cuda.device_count()))
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29740 + ci + (10 if sync_mode == 'peer' else 0)
    procs = [ctx.Process(target=_worker, args=(r, m, port, ci, q, 'nccl', sync_mode)) for r in range(m)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(m))
    for p in procs:
        p.join(60)
    assert all(len(v) == 0 for v in res.values()), res
