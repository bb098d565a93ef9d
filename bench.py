#!/usr/bin/env python
"""bench.py — Reddit-shape GIST GraphSAGE training throughput (epochs/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # CPU port of the reference path

For N > 1 launch under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One STEP = one cluster-batch training step of the rank's sub-GCN (device batch build ->
forward -> CE -> backward -> Adam), with the GIST sync + re-dispatch every
--iter-per-site steps inside the timed region.  One EPOCH = psize // batch_size steps.
With m = N sub-GCNs the reference runs n_epochs // m local epochs per rank
(cluster_gcn_ist_distrib.py:385), so one local pass on every rank counts as m epochs:
value = N * K / steps_per_epoch / seconds  (whole job).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

def metric_name(shape, model='sage'):
    return '%s_shape_gist_%s_epochs_per_s' % (shape, 'gat' if model == 'gat' else 'graphsage')


CONFIG_OF = {'reddit': 'configs[2]: Cluster-GCN GraphSAGE on a Reddit-shaped',
             'amazon2m': 'configs[3]: ultra-wide GIST GraphSAGE on an Amazon2M-shaped',
             'pubmed': 'PubMed-shaped', 'cora': 'Cora-shaped'}


def workload_string(a, n_nodes, n_edges, in_feats, psize, world):
    """config.workload — ONE string for both arms (the driver compares them)."""
    return ('%s synthetic graph (%d nodes, %d directed edges, %d feats, %d parts, batch %d parts), hidden %d, '
            '%d layers, GIST m=%d sub-GCNs (one per GPU), iter_per_site %d' % (
                CONFIG_OF[a.shape], n_nodes, n_edges, in_feats, psize, a.batch_size, a.n_hidden, a.n_layers + 1,
                world, a.iter_per_site))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=150)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='gist', choices=['gist', 'reference'])
    ap.add_argument('--shape', default='reddit', choices=['reddit', 'amazon2m', 'pubmed', 'cora'],
                    help='reddit = the headline config; amazon2m with --n-hidden 32768 --psize 15000 --gpus 8 is '
                         'the ultra-wide config 4 (on one GPU use --n-hidden 4096: the per-rank slice of m = 8)')
    ap.add_argument('--scale', type=float, default=1.0, help='<1 shrinks the graph (debug only; reported)')
    ap.add_argument('--model', default='sage', choices=['sage', 'gat'],
                    help='sage: cluster_gcn_ist_distrib.py (the headline); gat: cluster_gcn_ist_distrib_gat.py '
                         '(config 5; use --n-hidden 512 --n-heads 4 --n-layers 1)')
    ap.add_argument('--n-heads', type=int, default=4)
    ap.add_argument('--n-hidden', type=int, default=256)
    ap.add_argument('--n-layers', type=int, default=2)
    ap.add_argument('--psize', type=int, default=None, help='number of parts; default: the shape\'s (reddit 1500, amazon2m 15000)')
    ap.add_argument('--batch-size', type=int, default=20)
    ap.add_argument('--dropout', type=float, default=0.2)
    ap.add_argument('--lr', type=float, default=1e-2)
    ap.add_argument('--weight-decay', type=float, default=5e-4)
    ap.add_argument('--iter-per-site', type=int, default=None,
                    help='local steps between GIST syncs; default min(100, max(1, steps // 2)) so that EVERY '
                         'timed region holds at least one sync_model() + dispatch_model() round boundary')
    ap.add_argument('--first-epoch-rule', action='store_true',
                    help='apply the reference\'s "no re-dispatch during epoch 0" rule (…distrib.py:401) from the '
                         'first benchmarked step; default: steady state (rounds as in epochs >= 1: sync + dispatch)')
    ap.add_argument('--host-barriers', action='store_true',
                    help='call dist.barrier() before every dispatch / sync as the reference does (…distrib.py:402, 422); '
                         'default off: the packed all-gather is the synchronisation point, and a host-side barrier '
                         'drains the asynchronous step queue at every round boundary')
    ap.add_argument('--base-init', default=None, choices=['cpu', 'device'],
                    help='where rank 0 draws the full model\'s initial values; default cpu (reference-identical) '
                         'below hidden 8192, device above')
    ap.add_argument('--soak-s', type=float, default=1.0,
                    help='seconds of extra, untimed steps of the same loop run under the clock sampler after the '
                         'timed region (a 5 ms timed region cannot hold a 50 ms nvidia-smi sample)')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--cpu-steps', type=int, default=6, help='CPU-baseline sample size (steps)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eval-spmm', action='store_true')
    ap.add_argument('--no-timeline', action='store_true', help='skip the CUPTI timeline of the replayed step')
    ap.add_argument('--matmul', default='3xtf32', choices=['fp32', 'tf32', '3xtf32'],
                    help='nn.Linear contractions: the tcgen05 kernel (K4) in fp32-accurate 3xTF32 mode '
                         '(default; meets the 1e-5 parity tolerance), single-pass TF32 (~1e-3), or cuBLAS '
                         'fp32 sgemm through torch')
    ap.add_argument('--mode', default='graph', choices=['graph', 'eager'],
                    help='graph: whole training step captured in a CUDA graph; eager: op-by-op')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='graph mode: build batch k inside step k instead of overlapping the build of batch '
                         'k+1 with the training of batch k')
    ap.add_argument('--ncu', default='', choices=['', 'steps', 'fullgraph', 'fullgraph-all'],
                    help='bracket that region with cudaProfilerStart/Stop (ncu --profile-from-start off)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def tensor_peak_tf32():
    """TF32 dense tensor peak = half the bf16 figure (sustained: the GEMMs run inside a long step)."""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get('bf16_tflops_sustained', d['bf16_tflops'])) / 2, \
            'measured (MEASURED_PEAKS.json bf16_tflops_sustained / 2: kind::tf32 runs at half the bf16 rate)'
    return 1590.0 / 2, 'fallback (B200_PROFILING.md 1.59 PF/s bf16 / 2)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.index = index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        time.sleep(0.08)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, names = [], ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out['sm_max_mhz'] = float(r[2])
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == 'active':
                    reasons.add(nm)
        if sm:
            out['sm_mhz'] = float(np.median(sm))
        out['reasons'] = sorted(reasons)
        out['samples'] = len(sm)
        return out


def spmm_bytes(nnz, n_dst, n_src, d, scaled):
    """Algorithmic bytes of one SpMM launch (SURVEY.md §8d / BASELINE.md §4) and the
    compulsory lower bound."""
    alg = 4 * nnz * d + 4 * n_dst * d + 4 * nnz + 4 * (n_dst + 1) + 4 * n_dst * (1 if scaled else 0)
    comp = 4 * n_src * d + 4 * n_dst * d + 4 * nnz + 4 * (n_dst + 1)
    return alg, comp


def replay_timeline(loop, steps, record):
    """Run `steps` more steps of `loop` (all ranks, so round boundaries stay collective-safe); on the
    recording rank under torch.profiler (CUPTI).  Returns per-kernel-class busy time per step, its
    share of the step's wall time and of all kernel time, and kernel launches per step."""
    import torch
    if not record:
        for _ in range(steps):
            loop.next_step()
        torch.cuda.synchronize()
        return None
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                loop.next_step()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
               and 'memcpy' not in e.name.lower() and 'memset' not in e.name.lower()]
    except Exception as ex:          # profiling is diagnostic: never fail the bench on it
        return {'error': repr(ex)[:200], 'classes': {}}
    if not evs:
        return None
    t0 = min(e.time_range.start for e in evs)
    t1 = max(e.time_range.end for e in evs)
    wall = (t1 - t0) / steps

    def klass(name):
        n = name.lower()
        if 'spmm' in n:
            return 'spmm'
        if 'gemm_tf32' in n or 'splitk' in n:
            return 'gemm'
        if 'batch_' in n or 'chunk_' in n or 'gather_' in n or 'scan_' in n or 'seg_' in n:
            return 'batch_build'
        if 'gat_' in n:
            return 'gat'
        return 'other'
    cls = {}
    for e in evs:
        c = cls.setdefault(klass(e.name), [0, 0.0])
        c[0] += 1
        c[1] += e.time_range.end - e.time_range.start
    total = sum(c[1] for c in cls.values())
    return {'steps': steps, 'wall_us_per_step': round(wall, 2), 'kernels_per_step': round(len(evs) / steps, 1),
            'kernel_busy_us_per_step': round(total / steps, 2),
            'classes': {k: {'launches_per_step': round(c[0] / steps, 1), 'busy_us_per_step': round(c[1] / steps, 2),
                            'share_of_wall': round(c[1] / steps / wall, 4), 'share_of_kernel_time': round(c[1] / total, 4)}
                        for k, c in sorted(cls.items())},
            'note': 'branches of the graph overlap, so shares of wall can sum to more than 1; measured under the '
                    'profiler (shares only, not bench values)'}


def sync_summary(prof, world, timed_ms, steps):
    """GIST round boundaries inside the timed region: CUDA-event pairs recorded by
    DistributedGNNWrapper (sync = pack -> ONE all-gather -> local scatter; dispatch = local gathers)."""
    by = {}
    for r in prof:
        e = by.setdefault(r['what'], {'n': 0, 'ms': 0.0, 'rec': r})
        e['n'] += 1
        e['ms'] += r['ev0'].elapsed_time(r['ev1'])
    key = 'sync_peer_merge' if 'sync_peer_merge' in by else 'sync_all_gather'
    if key not in by:
        return {'rounds_in_timed_region': 0}
    peer = key == 'sync_peer_merge'
    rounds = by[key]['n']
    mean = lambda k: (by[k]['ms'] / by[k]['n']) if k in by else 0.0     # noqa: E731
    ag = by[key]['rec']
    ag_ms = mean(key)
    sync_ms = mean('sync_pack') + ag_ms + mean('sync_scatter')
    out = {
        'mode': 'peer memory: one kernel reads every site\'s packed slices over NVLink and scatters them into the '
                'local replica (no collective, no staging buffer)' if peer else
                'NCCL: one all_gather_into_tensor of the packed slices, then one local scatter launch',
        'rounds_in_timed_region': rounds, 'dispatches_in_timed_region': by.get('dispatch', {'n': 0})['n'],
        'ms': round(sync_ms, 4), 'pack_ms': round(mean('sync_pack'), 4),
        ('peer_merge_ms' if peer else 'all_gather_ms'): round(ag_ms, 4),
        'scatter_ms': round(mean('sync_scatter'), 4), 'dispatch_ms': round(mean('dispatch'), 4),
        'bytes_per_rank': ag['bytes_sent'], 'bytes_received_per_rank': ag['bytes_received'],
        'share_of_timed_region': round(sum(by.get(k, {'ms': 0})['ms'] for k in
                                           ('sync_pack', 'sync_all_gather', 'sync_peer_merge', 'sync_scatter', 'dispatch'))
                                       / max(timed_ms, 1e-9), 4),
        'what': 'per round: pack this rank\'s trained slices, exchange + merge them into the local full-model '
                'replica, dispatch = one K5 multi-gather launch with the next partition (no communication).  '
                'Event pairs on the training stream, rank 0.',
    }
    if world > 1 and ag_ms > 0:
        out['GBps'] = round(ag['bytes_received'] / 1e9 / (ag_ms / 1e3), 2)
        out['nvlink_line_rate_GBps'] = 900.0
        out['frac_of_line_rate'] = round(out['GBps'] / 900.0, 4)
        out['GBps_note'] = ('bytes received per rank / %s time (ingress per GPU; NVLink 5: 900 GB/s per direction)'
                            % ('merge-kernel (two device barriers included)' if peer else 'all-gather'))
    return out


# ------------------------------------------------------------------ gist arm --
def run_gist(a):
    import torch
    import torch.distributed as dist
    import gist_b200 as gb
    from gist_b200 import _lib, ops, synth
    from gist_b200.train import make_optimizer, train_step
    from types import SimpleNamespace

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == a.gpus, 'launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)' % (a.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert a.warmup >= 3, 'timing rules: warm-up >= 3 steps'
    if a.iter_per_site is None:
        a.iter_per_site = min(100, max(1, a.steps // 2))
    if a.base_init is None:
        a.base_init = 'device' if a.n_hidden >= 8192 else 'cpu'
    ops.set_matmul_precision(a.matmul)
    assert a.n_hidden % world == 0

    # same seeds on every rank (…distrib.py:570-572): partitions and batch order are
    # derived locally, never communicated
    torch.manual_seed(a.seed)
    np.random.seed(a.seed)
    random.seed(a.seed)

    t_setup = time.time()
    ds = synth.make(a.shape, seed=0, device=dev, scale=a.scale)
    g = synth.to_gist_graph(ds)
    n_nodes, n_edges = ds.num_nodes, int(ds.src.shape[0])
    in_feats, n_classes = ds.feat.shape[1], ds.num_classes
    train_nid = torch.nonzero(ds.train_mask).reshape(-1).cpu().numpy().astype(np.int64)
    psize = a.psize if (a.scale == 1.0 and a.psize) else int(ds.part.max().item()) + 1
    del ds
    wargs = SimpleNamespace(rank=rank, num_subnet=world, n_hidden=a.n_hidden, n_layers=a.n_layers,
                            dropout=a.dropout, use_layernorm=True, n_heads=a.n_heads)
    Wrapper = gb.DistributedGATWrapper if a.model == 'gat' else gb.DistributedGNNWrapper

    def fresh(h2d):
        """ClusterIter + wrapper in the state the reference has when train() starts."""
        random.seed(a.seed)
        torch.manual_seed(a.seed)
        it = gb.ClusterIter('', g, psize, a.batch_size, train_nid, use_pp=False, h2d=h2d)
        w = Wrapper(wargs, g, in_feats, n_classes, dev, **({} if a.model == 'gat' else {'base_init': a.base_init}))
        w.ini_sync_dispatch_model()
        return it, w

    class GraphLoop:
        """Same loop with the step body replayed from a CUDA graph (gist_b200/graphed.py)."""

        def __init__(self, it, w, readback):
            from gist_b200.graphed import GraphedClusterTrainer
            self.it, self.w, self.readback = it, w, readback
            w.inplace_dispatch = True
            w.sub_model.train()
            self.tr = GraphedClusterTrainer(it, w.sub_model, a.lr, a.weight_decay, h2d=it.h2d,
                                            pipeline=not a.no_pipeline).capture()
            self.total_iter = 0
            self.running_loss = 0.0
            self.loss_acc = torch.zeros((), device=dev)
            self.nodes = 0
            self.round_host_s, self.rounds = 0.0, 0         # host wall time spent in sync + dispatch calls

        def next_step(self):
            if self.total_iter % a.iter_per_site == 0 and self.total_iter > 0:
                e = self.total_iter // len(self.it) + (0 if a.first_epoch_rule else 1)
                t0 = time.perf_counter()
                if e > 0:                                   # …distrib.py:401
                    if world > 1 and a.host_barriers:
                        dist.barrier()
                    self.w.dispatch_model()
                self.tr.reset_optimizer()                   # fresh Adam at every round (…distrib.py:405-407)
                self.round_host_s += time.perf_counter() - t0
            if self.readback:
                # every step: ids H2D from pinned memory, loss D2H into pinned memory; the host
                # consumes the value one step late so the readback never stalls the launch queue
                v = self.tr.step_logged()
                if v is not None:
                    self.running_loss += v
            else:
                self.loss_acc += self.tr.step()
            self.nodes += self.tr.n_pad
            self.total_iter += 1
            if self.total_iter % a.iter_per_site == 0:
                t0 = time.perf_counter()
                if world > 1 and a.host_barriers:
                    dist.barrier()
                self.w.sync_model()
                self.round_host_s += time.perf_counter() - t0
                self.rounds += 1

        def finish(self):
            if self.readback:
                v = self.tr.drain()
                if v is not None:
                    self.running_loss += v

    class Loop:
        """The reference's step loop (…distrib.py:394-427) unrolled into next_step()."""

        def __init__(self, it, w, readback):
            self.it, self.w, self.readback = it, w, readback
            self.e, self.total_iter, self.opt = 0, 0, None
            self.iter = iter(it)
            self.running_loss = 0.0
            self.loss_acc = torch.zeros((), device=dev)
            self.nodes = 0

        def next_step(self):
            try:
                cluster = next(self.iter)
            except StopIteration:
                self.e += 1
                self.iter = iter(self.it)
                cluster = next(self.iter)
            if self.total_iter % a.iter_per_site == 0:
                if self.e + (0 if a.first_epoch_rule else 1) > 0 and self.total_iter > 0:   # …distrib.py:401
                    if world > 1 and a.host_barriers:
                        dist.barrier()
                    self.w.dispatch_model()
                self.w.sub_model.train()
                self.opt = make_optimizer(self.w.sub_model.parameters(), a.lr, a.weight_decay)
            loss = train_step(self.w.sub_model, self.opt, cluster)
            self.nodes += cluster.number_of_nodes()
            if self.readback:
                self.running_loss += float(loss)          # D2H + sync every step, as the reference
            else:
                self.loss_acc += loss
            self.total_iter += 1
            if self.total_iter % a.iter_per_site == 0:
                if world > 1 and a.host_barriers:
                    dist.barrier()
                self.w.sync_model()

    def timed(loop, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        r0 = loop.tr.replays if hasattr(loop, 'tr') else 0
        e0.record()
        for _ in range(k):
            loop.next_step()
        if hasattr(loop, 'finish'):
            loop.finish()              # last pending loss readback lands inside the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        n_launch = _lib.launch_count() - l0
        if hasattr(loop, 'tr'):     # kernel nodes of this library replayed from the CUDA graph
            n_launch += (loop.tr.replays - r0) * loop.tr.gist_launches_per_step
        launches = torch.tensor([n_launch], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        return ms.item(), int(launches.item())

    steps_per_epoch = psize // a.batch_size

    # Round-boundary warm-up on a throwaway wrapper: the first sync / re-dispatch of a process pays
    # one-off costs (NCCL's first all-gather sets up its channels, the slice kernels' module loads)
    # that otherwise land inside whichever timed region reaches step `iter_per_site` first.
    _, w0 = fresh('epoch')
    w0.inplace_dispatch = True
    w0.sync_model()
    w0.dispatch_model()
    torch.cuda.synchronize()
    del w0

    # ---- device-resident arm: node-id lists already in HBM, no loss readback -------
    it, w = fresh('epoch')
    LoopT = GraphLoop if a.mode == 'graph' else Loop
    loop = LoopT(it, w, readback=False)
    # clocks are sampled from the warm-up, through the timed region, to the end of an untimed soak of
    # the SAME loop (>= --soak-s seconds): the timed region alone (steps x ~0.25 ms) is shorter than
    # one nvidia-smi sampling period
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        loop.next_step()
    if a.ncu == 'steps':
        torch.cuda.profiler.start()
    if hasattr(w, 'profile'):
        w.profile = []
    rh0, rn0 = getattr(loop, 'round_host_s', 0.0), getattr(loop, 'rounds', 0)
    ms, launches = timed(loop, a.steps)
    round_host_ms = ((getattr(loop, 'round_host_s', 0.0) - rh0) * 1e3 / max(getattr(loop, 'rounds', 0) - rn0, 1))
    sync_prof = getattr(w, 'profile', None)
    if hasattr(w, 'profile'):
        w.profile = None
    if a.ncu == 'steps':
        torch.cuda.profiler.stop()
    value = world * a.steps / steps_per_epoch / (ms / 1e3)
    final_loss = float(loop.loss_acc) / max(a.steps + a.warmup, 1)
    soak_steps = 0
    if a.soak_s > 0 and not a.ncu:
        soak_steps = int(a.soak_s * 1e3 / max(ms / a.steps, 1e-3)) + 1     # same count on every rank (ms is the max over ranks)
        for i in range(soak_steps):
            loop.next_step()
            if i % 64 == 63:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks['window'] = ('warm-up + timed region + %d untimed soak steps of the same loop (%.2f s)'
                            % (soak_steps, soak_steps * ms / a.steps / 1e3))
    sync = sync_summary(sync_prof, world, ms, a.steps) if sync_prof is not None else None
    if sync is not None:
        sync['host_ms_per_round'] = round(round_host_ms, 3)      # wall time of the host inside sync + barrier + dispatch calls

    # ---- instrumented pass: per-launch SpMM durations with CUDA events --------------
    # Eager steps on the same workload, each preceded by a device-side sleep long enough for
    # the host to enqueue the whole step first: the event pairs around every SpMM launch then
    # measure back-to-back GPU execution, not host launch gaps.
    it3, w3 = fresh('epoch')
    loop3 = Loop(it3, w3, readback=False)
    for _ in range(3):
        loop3.next_step()
    prof_steps = min(a.steps, 20)
    sleep_cycles = int(6e-3 * 1.9e9)
    ops.SPMM_PROFILE = []
    ops.GEMM_PROFILE = []
    ops.GAT_PROFILE = []
    # every kernel alone on one stream: with the weight-gradient branch on its side stream the
    # event pairs would time a launch together with whatever GEMM happens to run beside it
    overlap, ops.OVERLAP_WEIGHT_GRADS = ops.OVERLAP_WEIGHT_GRADS, False
    step_evs = []
    torch.cuda.synchronize()
    for _ in range(prof_steps):
        torch.cuda._sleep(sleep_cycles)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        loop3.next_step()
        s1.record()
        step_evs.append((s0, s1))
        torch.cuda.synchronize()
    ops.OVERLAP_WEIGHT_GRADS = overlap
    prof, ops.SPMM_PROFILE = ops.SPMM_PROFILE, None
    gprof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    aprof, ops.GAT_PROFILE = ops.GAT_PROFILE, None
    prof_ms = sum(x.elapsed_time(y) for x, y in step_evs)
    del loop3, it3, w3
    alg_b = comp_b = spmm_ms = 0.0
    by_d = {}
    nnz_cache = {}
    for r in prof:
        key = r['rowptr'].data_ptr()
        if key not in nnz_cache:
            nnz_cache[key] = int(r['rowptr'][r['n_dst']].item())
        nnz = nnz_cache[key]
        t = r['ev0'].elapsed_time(r['ev1'])
        ab, cb = spmm_bytes(nnz, r['n_dst'], r['n_src'], r['d'], r['scaled'])
        alg_b += ab; comp_b += cb; spmm_ms += t
        e = by_d.setdefault(r['d'], [0, 0.0, 0.0])
        e[0] += 1; e[1] += ab; e[2] += t
    peak, peak_src = peaks()
    n_l = max(len(prof), 1)
    # DRAM bytes per SpMM launch from the committed `ncu --set full` capture of the same command
    # (profiles/README.md); null when the workload is not the one that was captured
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r2c_spmm_step_traffic.json')
    if a.shape == 'reddit' and a.scale == 1.0 and a.n_hidden == 256 and world == 1 and os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj['traffic_bytes_per_launch'], 'profiles/r2c_spmm_step_traffic.json: ' + tj['source']
    achieved = alg_b / 1e9 / (spmm_ms / 1e3) if spmm_ms > 0 else 0.0
    roofline = {
        'bound': 'hbm', 'effective_bound': 'L2 / gather latency: a cluster batch\'s operand (<= 6 MB) is L2-resident, so the '
                                           'algorithmic gather bytes never reach HBM (see traffic and dram_frac)',
        'dram_frac': (round(traffic / (spmm_ms * 1e-3 / n_l) / 1e9 / peak, 4) if traffic and spmm_ms > 0 else None),
        'kernel': 'spmm_seg_kernel + spmm_csr_kernel (all %d SpMM launches/step, fwd + transpose)' % (len(prof) // max(prof_steps, 1)),
        'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s', 'frac': round(achieved / peak, 4),
        'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
        'bytes_per_launch': round(alg_b / n_l), 'compulsory_bytes_per_launch': round(comp_b / n_l),
        'us_per_launch': round(spmm_ms * 1e3 / n_l, 2),
        'share_of_step': round((spmm_ms / prof_steps) / (ms / a.steps), 4),
        'by_width': {str(d): {'launches_per_step': v[0] / prof_steps, 'GBps': round(v[1] / 1e9 / (v[2] / 1e3), 1),
                              'us': round(v[2] * 1e3 / v[0], 2)} for d, v in sorted(by_d.items())},
        'note': 'cluster batches (<=5 MB of features) are L2-resident: algorithmic gather bytes are served by '
                'L2, so achieved can exceed the HBM copy peak; see roofline_fullgraph for the HBM-bound case',
    }

    if a.model == 'gat' and aprof:
        # K6 forward (edge softmax + weighted gather in one pass): per launch one D-wide source row
        # and one score per edge, the destination's score / output row / log-sum-exp per row
        gb_, gms = 0.0, 0.0
        for r in aprof:
            key = r['rowptr'].data_ptr()
            if key not in nnz_cache:
                nnz_cache[key] = int(r['rowptr'][r['n']].item())
            nnz = nnz_cache[key]
            hh = r.get('heads', 1)       # a batched launch aggregates all heads: per-head bytes x heads
            gb_ += hh * (4.0 * nnz * r['D'] + 4.0 * nnz + 4.0 * nnz + 4.0 * r['n'] * r['D'] + 4.0 * (r['n'] + 1) + 8.0 * r['n'])
            gms += r['ev0'].elapsed_time(r['ev1'])
        roofline = {'bound': 'hbm', 'kernel': 'gat_aggregate_kernel (K6 forward, %d launches/step)' % (len(aprof) // max(prof_steps, 1)),
                    'achieved': round(gb_ / 1e9 / (gms / 1e3), 1), 'peak': peak, 'unit': 'GB/s',
                    'frac': round(gb_ / 1e9 / (gms / 1e3) / peak, 4), 'traffic': None, 'peak_source': peak_src,
                    'bytes_per_launch': round(gb_ / len(aprof)), 'us_per_launch': round(gms * 1e3 / len(aprof), 2),
                    'share_of_step': round((gms / prof_steps) / (ms / a.steps), 4),
                    'note': 'algorithmic bytes = 4*nnz*D (source rows) + 8*nnz (col, source score) + 4*n*D (out) + '
                            'rowptr + per-row score/lse; batches are L2-resident (see the SAGE note)'}

    # tensor roofline of the dense contractions (K4): useful FLOP = 2MNK per product; in 3xTF32
    # mode the tensor core executes three MMAs per useful one
    roofline_gemm = None
    if gprof:
        tpeak, tsrc = tensor_peak_tf32()
        fl = sum(2.0 * r['M'] * r['N'] * r['K'] for r in gprof)
        mma_fl = sum(2.0 * r['M'] * r['N'] * r['K'] * r['passes'] for r in gprof)
        gms = sum(r['ev0'].elapsed_time(r['ev1']) for r in gprof)
        big = max(gprof, key=lambda r: r['M'] * r['N'] * r['K'])
        bsel = [r for r in gprof if (r['M'], r['N'], r['K']) == (big['M'], big['N'], big['K'])]
        bms = sum(r['ev0'].elapsed_time(r['ev1']) for r in bsel) / len(bsel)
        bfl = 2.0 * big['M'] * big['N'] * big['K']
        roofline_gemm = {
            'bound': 'tensor', 'kernel': 'gemm_tf32_kernel (all %d launches/step)' % (len(gprof) // max(prof_steps, 1)),
            'achieved': round(fl / gms / 1e9, 1), 'peak': tpeak, 'unit': 'TFLOP/s',
            'frac': round(fl / gms / 1e9 / tpeak, 4), 'peak_source': tsrc,
            'mma_achieved': round(mma_fl / gms / 1e9, 1), 'mma_frac': round(mma_fl / gms / 1e9 / tpeak, 4),
            'share_of_step': round((gms / prof_steps) / (ms / a.steps), 4),
            'largest': {'M': big['M'], 'N': big['N'], 'K': big['K'], 'us': round(bms * 1e3, 1),
                        'TFLOPs': round(bfl / bms / 1e9, 1), 'mma_TFLOPs': round(bfl * big['passes'] / bms / 1e9, 1)},
            'note': 'achieved = useful fp32-equivalent FLOP (2MNK) / time; mma_* counts the %d TF32 MMA passes the '
                    'tensor core actually executes' % big['passes'],
        }

    # ---- end-to-end arm: node ids from pinned host memory + loss readback every step --
    # (Runs BEFORE the CUPTI timeline below: once torch.profiler has been active in a process, every later launch
    # and event call of the host pays the profiler's callbacks — measured on the same box, same code: e2e 0.2639
    # ms/step behind the timeline vs 0.2405 with --no-timeline, while the device-resident arm, timed earlier,
    # read 0.2341 / 0.2318.  The e2e arm is the one bounded by host time per step.)
    it2, w2 = fresh('step')
    loop2 = LoopT(it2, w2, readback=True)
    for _ in range(a.warmup):
        loop2.next_step()
    h2d_of = (lambda: loop2.tr.h2d_bytes) if a.mode == 'graph' else (lambda: it2.h2d_bytes)
    h2d0 = h2d_of()
    ms2, _ = timed(loop2, a.steps)
    e2e = {'value': round(world * a.steps / steps_per_epoch / (ms2 / 1e3), 4), 'unit': 'epochs/s',
           'h2d_bytes_per_step': int((h2d_of() - h2d0) / a.steps), 'd2h_bytes_per_step': 4,
           'ms_per_step': round(ms2 / a.steps, 4),
           'what': 'public API (ClusterIter -> GraphedClusterTrainer.step_logged): batch node ids copied from '
                   'pinned host memory every step, the loss copied to pinned host memory every step and summed '
                   'on the host (consumed one step late in graph mode, float(loss) in eager mode); graph + '
                   'features resident in HBM',
           'mean_loss_host': round(loop2.running_loss / max(a.steps + a.warmup, 1), 4)}

    # ---- kernel shares of the REPLAYED step (CUPTI timeline of a few more steps of the same loop) ----
    # roofline.share_of_step above divides eager single-stream kernel time by the overlapped graph
    # step; the replayed graph runs three branches concurrently, so the honest figures are each
    # kernel class's busy time over the step's wall time, from the timeline of the replay itself.
    timeline = None
    if a.mode == 'graph' and not a.no_timeline and not a.ncu:
        timeline = replay_timeline(loop, 8, rank == 0)
        if timeline is not None:
            roofline['share_of_step'] = timeline['classes'].get('spmm', {}).get('share_of_wall', roofline['share_of_step'])
            roofline['share_of_step_source'] = 'CUPTI timeline of the replayed graph: SpMM busy time / step wall time'
            if roofline_gemm is not None and 'gemm' in timeline['classes']:
                roofline_gemm['share_of_step'] = timeline['classes']['gemm']['share_of_wall']
                roofline_gemm['share_of_step_source'] = roofline['share_of_step_source'].replace('SpMM', 'GEMM')

    # ---- the HBM-bound case: full-graph SpMM that evaluate() runs (rank 0) -------------
    full = None
    if rank == 0 and not a.no_eval_spmm:
        full = {}
        n = g.number_of_nodes()
        for d in (in_feats, a.n_hidden):
            # operands laid out as the product lays feature matrices out in HBM (ops.pad_rows): rows
            # padded to a multiple of 4 floats, so d = 602 is gathered with 128-bit loads
            x = ops._padded_empty(n, d, dev).normal_()
            y = ops._padded_empty(n, d, dev)
            best = {}
            variants = [('auto', 0, x, y), ('vec128', 3 << _lib.SPMM_VEC_SHIFT, x, y), ('vec64', 2 << _lib.SPMM_VEC_SHIFT, x, y),
                        ('vec32', 1 << _lib.SPMM_VEC_SHIFT, x, y)]
            if d % 4:      # round 1's layout: contiguous rows, 64-bit gathers
                variants.append(('unpadded_rows', 0, x.contiguous(), torch.empty(n, d, device=dev)))
            for name, fl, xx, yy in variants:
                for _ in range(2):
                    ops.spmm_raw(g.rowptr, g.col_buffer, n, n, xx, yy, dst_scale=g.inv_in_degree(), flags=fl)
                torch.cuda.synchronize()
                ts = []
                if (a.ncu == 'fullgraph' and name == 'auto') or a.ncu == 'fullgraph-all':
                    torch.cuda.profiler.start()
                    ops.spmm_raw(g.rowptr, g.col_buffer, n, n, xx, yy, dst_scale=g.inv_in_degree(), flags=fl)
                    torch.cuda.synchronize()
                    torch.cuda.profiler.stop()
                for _ in range(5):
                    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    f0.record()
                    ops.spmm_raw(g.rowptr, g.col_buffer, n, n, xx, yy, dst_scale=g.inv_in_degree(), flags=fl)
                    f1.record()
                    torch.cuda.synchronize()
                    ts.append(f0.elapsed_time(f1))
                best[name] = float(np.mean(ts))
            del variants
            ab, cb = spmm_bytes(n_edges, n, n, d, True)
            t = best['auto']
            full['d%d' % d] = {'ms': round(t, 3), 'achieved': round(ab / 1e9 / (t / 1e3), 1), 'peak': peak,
                               'unit': 'GB/s', 'frac': round(ab / 1e9 / (t / 1e3) / peak, 4),
                               'algorithmic_bytes': ab, 'compulsory_bytes': cb,
                               'ms_by_variant': {k: round(v, 3) for k, v in best.items()},
                               'flush': 'operand (%.0f MB) larger than L2' % (4 * n * d / 1e6)}
            fp = os.path.join(ROOT, 'profiles', 'r2_spmm_fullgraph_traffic.json')
            if a.shape == 'reddit' and a.scale == 1.0 and os.path.exists(fp):
                # the same launch under `ncu --set full` (committed capture): what actually moved
                fl_ = json.load(open(fp))['launches']
                cap = fl_[0] if d == in_feats else (fl_[1] if len(fl_) > 1 else None)
                if cap:
                    dram = cap['dram_read_bytes'] + cap['dram_write_bytes']
                    full['d%d' % d].update({
                        'traffic': round(dram), 'traffic_over_compulsory': round(dram / cb, 1),
                        'dram_frac': round(dram / 1e9 / (cap['us'] / 1e6) / peak, 4),
                        'l2_to_sm_TBps': cap.get('xbar_to_l1_TBps'), 'l2_hit_pct': cap.get('l2_hit_pct'),
                        'traffic_source': 'profiles/r2_spmm_fullgraph_traffic.json (ncu --set full of this launch: %.2f ms under the profiler)' % (cap['us'] / 1e3),
                        'effective_bound': 'L2->SM crossbar / gather latency: frac counts every gathered byte (SURVEY 8d no-reuse model), '
                                           'dram_frac what reached HBM'})
            del x, y

    # ---- evaluate() on the full graph (a12; outside the reference's epoch timer) ---------
    # reference structure: two full forward passes (val, test), aggregate-first in every layer;
    # here: one pass for both masks, project-first wherever the output is narrower than the input
    ev = None
    if rank == 0 and not a.no_eval_spmm and getattr(w, 'base_model', None) is not None:
        from gist_b200.train import evaluate_masks
        labels, vm, tm = g.ndata['label'], g.ndata['val_mask'], g.ndata['test_mask']

        def timed_eval(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                fn()
            f1.record()
            torch.cuda.synchronize()
            return f0.elapsed_time(f1) / reps

        def ref_structure():
            w.base_model.eval()
            for m_ in (vm, tm):
                with torch.enable_grad():           # grad mode on = the aggregate-first training form
                    pred = w.base_model(g).argmax(dim=1)
                ((pred == labels) & m_).sum()
        ms_fast = timed_eval(lambda: evaluate_masks(w.base_model, g, labels, [vm, tm]))
        ms_ref = timed_eval(ref_structure)
        ev = {'ms_val_plus_test': round(ms_fast, 2), 'ms_reference_structure': round(ms_ref, 2),
              'what': 'full-graph inference of the full-width model for the val and test masks: one pass, '
                      'project-first layers (aggregate at the output width) vs two aggregate-first passes '
                      '(utils.py:70-80 called twice, cluster_gcn_ist_distrib.py:439-446)'}
        w.base_model.train()

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ---------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.model == 'sage':
        cpu = cpu_baseline(a, it.g, steps_per_epoch, it)

    if rank == 0:
        line = {
            'metric': metric_name(a.shape, a.model), 'value': round(value, 4), 'unit': 'epochs/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': round(ms / a.steps, 4),
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'per_rank_steps_per_s': round(a.steps / (ms / 1e3), 1),
            'epochs_per_s_one_pass_by_all_ranks': round(value / world, 4),
            'dtype': {'fp32': 'f32', '3xtf32': 'f32 (GEMMs: 3xTF32 split-operand tensor-core products, fp32-accurate)',
                      'tf32': 'f32 (aggregation, norms, loss, Adam) + tf32 tensor-core GEMM inputs'}[a.matmul],
            'data': 'synthetic',
            'config': {
                'workload': workload_string(a, n_nodes, n_edges, in_feats, psize, world),
                'reference_caller': ('cluster_gcn_ist_distrib_gat.py (GAT, %d heads: fused edge-softmax + weighted SpMM, K6)' % a.n_heads
                                     if a.model == 'gat' else 'cluster_gcn_ist_distrib.py (GraphSAGE aggregation path)'),
                'steps_per_epoch': steps_per_epoch, 'num_subnet': world, 'scale': a.scale, 'mode': a.mode,
                'pipeline': (a.mode == 'graph' and not a.no_pipeline),
                'epoch_accounting': 'value: GIST accounting — m ranks x one local pass = m epochs (reference: '
                                    'local_epochs = n_epochs // num_subnet, …distrib.py:385); '
                                    'epochs_per_s_one_pass_by_all_ranks = value / m is SURVEY 8(d)\'s literal '
                                    'definition (one pass executed concurrently by all m ranks = 1 epoch)',
                'scaling_note': 'a FIXED model (hidden %d) and a fixed epoch budget are split m = N ways: every rank '
                                'trains a hidden/m-wide sub-GCN over the full batch sequence for n_epochs/m local '
                                'epochs, so the job is fixed and per-GPU work shrinks with N (strong scaling of the '
                                'GIST job; neither per-GPU work nor per-GPU model is constant)' % a.n_hidden,
                'l2': 'inputs larger than L2: every step gathers a different ~%d-node batch from the %.0f MB '
                      'training feature matrix' % (loop.nodes // max(loop.total_iter, 1),
                                                   4.0 * it.g.number_of_nodes() * in_feats / 1e6),
                'gemm': {'tf32': 'hand-written tcgen05 kernel, TF32 inputs / fp32 accumulate in TMEM (K4)',
                         '3xtf32': 'hand-written tcgen05 kernel (K4), 3xTF32: operands split x = hi + lo, three TF32 '
                                   'MMAs per K step, fp32 accumulate in TMEM; ~1e-6 relative vs fp64',
                         'fp32': 'cuBLAS fp32 sgemm via torch'}[a.matmul],
                'loss_after': round(final_loss, 4),
                'step_options': _step_options(),
            },
            'roofline': roofline, 'roofline_gemm': roofline_gemm, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'sync': sync, 'replay_timeline': timeline,
            'clocks': clocks, 'roofline_fullgraph': full, 'evaluate': ev,
            'setup_s': round(time.time() - t_setup, 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _step_options():
    """The A/B switches the captured step ran with (environment-controlled; defaults = the measured best)."""
    from gist_b200 import _lib, graph as gg, graphed, ops
    return {'builder': gg.BUILDER, 'fused_tail': graphed.FUSED_TAIL, 'layernorm_in_splitk_pass_up_to_256': ops.FUSED_LN_WIDE,
            'segment_records': ops.SEG_META, 'queue_prefetch': ops.SEG_PREFETCH, 'ids_staged_ahead': graphed.STAGE_IDS_AHEAD,
            'programmatic_dependent_launch': _lib.get_pdl(), 'merge': os.environ.get('GIST_MERGE', 'auto')}


def cpu_baseline(a, train_g, steps_per_epoch, it, threads=None):
    """Oracle-side CPU port of the reference step, timed on this box's host cores on a
    bounded sample (a.cpu_steps steps) of the same batches."""
    import torch
    from oracle import cpu_reference as R
    if threads:
        torch.set_num_threads(threads)
    rowptr = train_g.rowptr.cpu().numpy().astype(np.int64)
    col = train_g.col.cpu().numpy().astype(np.int64)
    feat = train_g.ndata['feat'].cpu()
    label = train_g.ndata['label'].cpu()
    world = max(int(os.environ.get('WORLD_SIZE', '1')), 1)
    tr = R.CpuClusterTrainer(rowptr, col, feat, label, feat.shape[1], a.n_hidden, int(label.max()) + 1,
                             a.n_layers, num_subnet=world, dropout=a.dropout, lr=a.lr,
                             weight_decay=a.weight_decay, seed=a.seed)
    batches = [it.batch_node_ids(i)[1] for i in range(min(a.cpu_steps + 1, len(it)))]
    sec = R.time_steps(tr, batches, warmup=1)
    return {'value': round(1.0 / (sec * steps_per_epoch), 5), 'unit': 'epochs/s',
            'cores': torch.get_num_threads(), 'host_cpus': os.cpu_count(), 'kind': 'port',
            'ms_per_step': round(sec * 1e3, 2),
            'sample': '%d training steps of the same workload (CPU induced-subgraph per step + torch CSR SpMM '
                      '+ autograd + Adam), extrapolated to %d steps/epoch' % (len(batches) - 1, steps_per_epoch)}


# -------------------------------------------------------------- reference arm --
def run_reference(a):
    """The reference's own CPU implementation of the path cannot be installed (DGL 0.5.3
    wheel absent); this arm times the oracle's CPU port on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    if a.iter_per_site is None:
        a.iter_per_site = min(100, max(1, a.steps // 2))
    import torch
    import scipy.sparse as sp
    from gist_b200 import synth          # synthetic-shape generator only (plain torch ops)
    from oracle import cpu_reference as R
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank,
    # and the other ranks have exited, so rank 0 takes the box's cores back
    try:
        torch.set_num_threads(max(len(os.sched_getaffinity(0)), 1))
    except AttributeError:
        torch.set_num_threads(max(os.cpu_count() or 1, 1))
    random.seed(a.seed)
    torch.manual_seed(a.seed)
    use_cuda = torch.cuda.is_available()
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))) if use_cuda else torch.device('cpu')
    # Setup (not timed).  The same seeded generator as the gist arm, so both arms see the
    # same graph; everything after generation is host code: none of this repo's kernels,
    # graph object or sampler is on this arm's path.
    ds = synth.make(a.shape, seed=0, device=dev, scale=a.scale)
    n_nodes, n_edges, in_feats = ds.num_nodes, int(ds.src.shape[0]), ds.feat.shape[1]
    src, dst = ds.src.cpu().numpy(), ds.dst.cpu().numpy()
    A = sp.csr_matrix((np.ones(len(src), dtype=np.float32), (dst, src)), shape=(n_nodes, n_nodes))
    del src, dst
    train_nid = torch.nonzero(ds.train_mask).reshape(-1).cpu().numpy().astype(np.int64)
    psize = a.psize if (a.scale == 1.0 and a.psize) else int(ds.part.max().item()) + 1
    At = A[train_nid][:, train_nid].tocsr()          # sampler.py:34 training graph
    At.sort_indices()
    feat, label = ds.feat.cpu()[train_nid], ds.label.cpu()[train_nid]
    part = ds.part.cpu().numpy()[train_nid]
    del ds, A
    order = np.argsort(part, kind='stable')
    bounds = np.searchsorted(part[order], np.arange(psize + 1))
    par_li = [order[bounds[p]:bounds[p + 1]].astype(np.int64) for p in range(psize)]
    random.shuffle(par_li)                            # sampler.py:55
    steps_per_epoch = psize // a.batch_size
    tr = R.CpuClusterTrainer(At.indptr.astype(np.int64), At.indices.astype(np.int64), feat, label, in_feats,
                             a.n_hidden, int(label.max()) + 1, a.n_layers, num_subnet=world, dropout=a.dropout,
                             lr=a.lr, weight_decay=a.weight_decay, seed=a.seed)

    def batch_ids(i):
        return np.concatenate(par_li[i * a.batch_size:(i + 1) * a.batch_size]).astype(np.int64)
    # bounded sample: the CPU step is ~50-100 ms (Reddit shape), ~2 s at the ultra-wide widths
    cap = 400 if a.n_hidden < 2048 else 12
    k = min(a.steps, cap)
    wu = min(max(a.warmup, 1), cap)
    batches = [batch_ids(i % steps_per_epoch) for i in range(k + wu)]
    sec = R.time_steps(tr, batches, warmup=wu)
    # N sub-GCNs time-share the same host cores: N local passes take N x as long
    value = 1.0 / (sec * steps_per_epoch)
    line = {
        'impl': 'reference', 'metric': metric_name(a.shape), 'value': round(value, 5), 'unit': 'epochs/s', 'n_gpus': world,
        'steps': k, 'warmup': wu, 'ms_per_step': round(sec * 1e3, 3), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'epoch_accounting': 'same as the gist arm (GIST accounting): the m = N sub-GCNs (hidden/m wide) time-share this '
                            'host, so their m local passes — m epochs — take m x one pass: value = m / (m x pass time) '
                            '= 1 / pass time of ONE hidden/m-wide sub-GCN',
        'config': {'workload': workload_string(a, n_nodes, n_edges, in_feats, psize, world),
                   'reference_note': 'CPU port of the reference step (oracle/cpu_reference.py; DGL 0.5.3 is not installable): '
                                     'one hidden/m-wide sub-GCN on all host threads, no sync (single process)',
                   'steps_per_epoch': steps_per_epoch, 'scale': a.scale},
        'cpu_baseline': {'value': round(value, 5), 'unit': 'epochs/s', 'cores': torch.get_num_threads(),
                         'host_cpus': os.cpu_count(), 'kind': 'port',
                         'sample': '%d steps (bounded sample of the epoch), all host threads' % k},
        'e2e': {'value': round(value, 5), 'unit': 'epochs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gist(args)
