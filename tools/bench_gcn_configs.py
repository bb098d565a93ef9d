#!/usr/bin/env python
"""Measured lines for BASELINE.json configs[0] and configs[1] (the full-graph GraphConv trainers):

  configs[0]  gcn/train.py     2-layer GCN, Cora shape (2708 nodes, 10556 edges + self loops, 1433 feats, hidden 16)
  configs[1]  gcn/train_ist.py 3-layer GCN, PubMed shape (19717 nodes, 88648 edges + self loops, 500 -> 496 feats),
              hidden 256, 8 sub-GCNs, split_output, iter_per_site 5 (script/sweep.py)

For each: epochs/s of gist_b200's trainer on cuda:0 (CUDA events, evaluation excluded as in the reference's
timer) and of the oracle's CPU port on the host cores beside it (reported baseline, kind = "port").
One JSON line per config on stdout.  usage: bench_gcn_configs.py [--epochs K] [--warmup W] [--matmul MODE]"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gist_b200 import _lib, ops, synth                      # noqa: E402
from gist_b200.graph import GistGraph                       # noqa: E402
from gist_b200.train_gcn import GCNTrainer                  # noqa: E402
from gist_b200.train_ist import ISTGCNTrainer, add_self_loops, random_projection   # noqa: E402


def time_gpu(tr, k, w, ncu=False):
    for e in range(w):
        tr.train_epoch(e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    if ncu:
        torch.cuda.profiler.start()
    e0.record()
    loss = None
    for e in range(w, w + k):
        loss = tr.train_epoch(e)
    e1.record()
    torch.cuda.synchronize()
    if ncu:
        torch.cuda.profiler.stop()
    n_l = _lib.launch_count() - l0
    cap = getattr(tr, '_captured', None)
    if cap is not None:
        n_l += k * cap.gist_launches              # kernel nodes of this library replayed from the graph
    return e0.elapsed_time(e1) / k, float(loss), n_l


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--cpu-epochs', type=int, default=5)
    ap.add_argument('--matmul', default='3xtf32', choices=['fp32', 'tf32', '3xtf32'])
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--ncu', action='store_true', help='bracket the timed epochs with cudaProfilerStart/Stop')
    ap.add_argument('--only', type=int, default=None, help='run only this config (0 or 1)')
    ap.add_argument('--eager', action='store_true', help='op-by-op instead of one CUDA-graph replay per epoch')
    a = ap.parse_args()
    ops.set_matmul_precision(a.matmul)
    dev = torch.device('cuda', 0)
    from oracle import cpu_reference as R                   # CPU baseline leg only

    for cfg in ((0, 1) if a.only is None else (a.only,)):
        shape = 'cora' if cfg == 0 else 'pubmed'
        ds = synth.make(shape, seed=0)
        feat = ds.feat
        if cfg == 1:
            feat = random_projection(feat, 8, seed=0)       # train_ist.py:70-81 (500 -> 496)
        src, dst = add_self_loops(ds.src, ds.dst, ds.num_nodes)
        g = GistGraph.from_edges(src, dst, ds.num_nodes, device=dev)
        x, y, tm = feat.to(dev), ds.label.to(dev), ds.train_mask.to(dev)
        torch.manual_seed(0)
        if cfg == 0:
            args = SimpleNamespace(n_hidden=16, n_layers=1, dropout=0.5, lr=1e-3, weight_decay=5e-4, n_epochs=400,
                                   use_layernorm='True', lr_scheduler=False)
            tr = GCNTrainer(g, x, y, tm, ds.num_classes, args, dev, use_graph=not a.eager)
            what = 'configs[0]: gcn/train.py 2-layer GCN, Cora-shaped synthetic graph'
        else:
            args = SimpleNamespace(n_hidden=256, n_layers=2, num_subnet=8, iter_per_site=5, dropout=0.5, lr=1e-2,
                                   weight_decay=5e-4, n_epochs=10 ** 6, split_input='False', split_output='True',
                                   use_layernorm='True')
            tr = ISTGCNTrainer(g, x, y, tm, ds.num_classes, args, dev, use_graph=not a.eager)
            what = 'configs[1]: gcn/train_ist.py 3-layer GCN, PubMed-shaped synthetic graph, 8 sub-GCNs, iter_per_site 5'
        ms, loss, launches = time_gpu(tr, a.epochs, a.warmup, a.ncu)
        line = {'metric': '%s_shape_gcn_epochs_per_s' % shape, 'value': round(1e3 / ms, 2), 'unit': 'epochs/s',
                'n_gpus': 1, 'steps': a.epochs, 'warmup': a.warmup, 'ms_per_step': round(ms, 4),
                'higher_is_better': True, 'dtype': 'f32' if a.matmul == 'fp32' else 'f32 (GEMMs %s)' % a.matmul,
                'data': 'synthetic',
                'config': {'workload': '%s (%d nodes, %d directed edges incl. self loops, %d feats)' % (
                    what, ds.num_nodes, int(src.shape[0]), x.shape[1]), 'eval': 'excluded (as the reference timer)',
                           'mode': 'eager' if a.eager else 'graph (one replay per epoch)'},
                'loss_after': round(loss, 4), 'gpu_launches': launches}
        if not a.no_cpu:
            rp, cl = g.rowptr.cpu().numpy(), g.col.cpu().numpy()
            if cfg == 0:
                ct = R.CpuGCNTrainer(rp, cl, feat, ds.label, ds.train_mask, 16, ds.num_classes, 1, lr=1e-3)
            else:
                ct = R.CpuISTGCNTrainer(rp, cl, feat, ds.label, ds.train_mask, 256, ds.num_classes, 2, 8)
            sec = R.time_epochs(ct, a.cpu_epochs, warmup=1)
            line['cpu_baseline'] = {'value': round(1.0 / sec, 3), 'unit': 'epochs/s', 'cores': torch.get_num_threads(),
                                    'kind': 'port', 'sample': '%d epochs of the same workload (torch CSR SpMM + '
                                    'autograd + Adam on the host)' % a.cpu_epochs}
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
