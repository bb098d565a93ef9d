#!/usr/bin/env python
"""Where does the host time of one end-to-end step go?  (diagnostic, not a bench)"""
import os, random, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gist_b200 as gb
from gist_b200 import ops, synth
from gist_b200.graphed import GraphedClusterTrainer
from types import SimpleNamespace

dev = torch.device('cuda', 0)
ops.set_matmul_precision(sys.argv[1] if len(sys.argv) > 1 else '3xtf32')
random.seed(0); torch.manual_seed(0)
ds = synth.make('reddit', seed=0, device=dev, scale=1.0)
g = synth.to_gist_graph(ds)
train_nid = torch.nonzero(ds.train_mask).reshape(-1).cpu().numpy().astype(np.int64)
in_feats, ncls = ds.feat.shape[1], ds.num_classes
del ds
it = gb.ClusterIter('', g, 1500, 20, train_nid, use_pp=False, h2d='step')
w = gb.DistributedGNNWrapper(SimpleNamespace(rank=0, num_subnet=1, n_hidden=256, n_layers=2, dropout=0.2,
                                             use_layernorm=True), g, in_feats, ncls, dev)
w.ini_sync_dispatch_model(); w.inplace_dispatch = True; w.sub_model.train()
tr = GraphedClusterTrainer(it, w.sub_model, 1e-2, 5e-4, h2d='step').capture()
for _ in range(10):
    tr.step_logged()
torch.cuda.synchronize()
N = 300
T = np.zeros(N)
torch.cuda.synchronize()
t_all0 = time.perf_counter()
for k in range(N):
    t0 = time.perf_counter()
    tr.step_logged()
    T[k] = time.perf_counter() - t0
tr.drain()
torch.cuda.synchronize()
tot = time.perf_counter() - t_all0
c = T * 1e6
print('step_logged: wall per step %.1f us; host call mean %.1f p50 %.1f p90 %.1f max %.1f us' % (
    tot / N * 1e6, c.mean(), np.median(c), np.percentile(c, 90), c.max()))
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(N):
    tr.step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('step (no readback): host enqueue %.1f us/step, wall %.1f us/step' % ((t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
