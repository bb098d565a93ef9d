#!/usr/bin/env python
"""Kernel timeline of a few replays of the captured training step (CUPTI via torch.profiler):
writes gpurun_out/timeline_<tag>.csv with one row per kernel (step, stream, start_us, dur_us, name).
Diagnostic for the step's critical path; numbers under the profiler are not bench values."""
import csv, os, random, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gist_b200 as gb
from gist_b200 import ops, synth
from gist_b200.graphed import GraphedClusterTrainer
from types import SimpleNamespace
from torch.profiler import profile, ProfilerActivity

prec = sys.argv[1] if len(sys.argv) > 1 else '3xtf32'
pipeline = (sys.argv[2] != 'nopipe') if len(sys.argv) > 2 else True
hidden = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device('cuda', 0)
ops.set_matmul_precision(prec)
random.seed(0); torch.manual_seed(0)
ds = synth.make('reddit', seed=0, device=dev, scale=1.0)
g = synth.to_gist_graph(ds)
train_nid = torch.nonzero(ds.train_mask).reshape(-1).cpu().numpy().astype(np.int64)
in_feats, ncls = ds.feat.shape[1], ds.num_classes
del ds
it = gb.ClusterIter('', g, 1500, 20, train_nid, use_pp=False, h2d='epoch')
w = gb.DistributedGNNWrapper(SimpleNamespace(rank=0, num_subnet=1, n_hidden=hidden, n_layers=2, dropout=0.2,
                                             use_layernorm=True), g, in_feats, ncls, dev)
w.ini_sync_dispatch_model(); w.inplace_dispatch = True; w.sub_model.train()
tr = GraphedClusterTrainer(it, w.sub_model, 1e-2, 5e-4, h2d='epoch', pipeline=pipeline).capture()
for _ in range(20):
    tr.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        tr.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
tag = '%s_%s_h%d' % (prec, 'pipe' if pipeline else 'nopipe', hidden)
with open(os.path.join(ROOT, 'gpurun_out', 'timeline_%s.csv' % tag), 'w', newline='') as f:
    wr = csv.writer(f)
    wr.writerow(['start_us', 'dur_us', 'end_us', 'stream', 'name'])
    for e in evs:
        s = e.time_range.start - t0
        d = e.time_range.end - e.time_range.start
        stream = getattr(e, 'stream', None)
        wr.writerow(['%.2f' % s, '%.2f' % d, '%.2f' % (s + d), stream, e.name[:90]])
print('wrote', len(evs), 'kernel events, span %.1f us' % (evs[-1].time_range.end - t0))
