#!/usr/bin/env python
"""GIST merge (sync_model's scatter of all m sites' slices into the local full-model replica) at the
ultra-wide size of configuration 4, on ONE GPU: the all-gather lands in a local buffer anyway, so m packed
slice sets in local memory reproduce the merge exactly.  Times the 4-byte scatter kernel (GIST_MERGE=scatter)
and the row-streaming kernel (GIST_MERGE=rows) with CUDA events and checks that both leave the same replica.

    python tools/merge_bench.py [hidden=32768] [m=8] [n_layers=2] [in_feats=100] [n_classes=47]

One JSON line per mode; DRAM bytes are the whole-sector model of profiles/README.md (touched sectors x 32 B
read + written)."""
import json
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gist_b200 import _lib  # noqa: E402
from gist_b200.ist import DistributedGNNWrapper  # noqa: E402

hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_layers = int(sys.argv[3]) if len(sys.argv) > 3 else 2
in_feats = int(sys.argv[4]) if len(sys.argv) > 4 else 100
n_classes = int(sys.argv[5]) if len(sys.argv) > 5 else 47
reps = 3
dev = torch.device('cuda', 0)
random.seed(0); torch.manual_seed(0); np.random.seed(0)
args = SimpleNamespace(rank=0, num_subnet=m, n_hidden=hidden, n_layers=n_layers, dropout=0.0, use_layernorm=True)
w = DistributedGNNWrapper(args, None, in_feats, n_classes, dev, base_init='device')
parts = w._to_dev(w.sample_partitions())
numel = sum(t.numel() for lyr in w.sub_model.layers for t in (lyr.linear.weight, lyr.linear.bias))
gathered = torch.randn(m, numel, device=dev)
torch.cuda.synchronize()
big = max(w.base_model.parameters(), key=lambda p: p.numel())
touched_elems = m * max(t.numel() for lyr in w.sub_model.layers for t in (lyr.linear.weight,))
results = {}
check = None
for mode in ('scatter', 'rows', 'scatter', 'rows'):
    os.environ['GIST_MERGE'] = mode
    times = []
    l0 = _lib.launch_count()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            w._merge(gathered, parts)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    launches = (_lib.launch_count() - l0) // reps
    # the replica after a merge is independent of what it held at the merged positions before
    digest = torch.stack([p.detach().double().sum() for p in w.base_model.parameters()]).cpu()
    if check is None:
        check = digest
        ref_big = big.detach().clone() if big.numel() * 4 < 40e9 else None
    same = bool(torch.equal(digest, check)) and (ref_big is None or bool(torch.equal(big.detach(), ref_big)))
    line = dict(what='gist_merge', mode=mode, hidden=hidden, m=m, n_layers=n_layers, ms=round(min(times), 4),
                ms_all=[round(t, 4) for t in times], launches=launches, bytes_merged=int(m * numel * 4),
                largest_replica_GB=round(big.numel() * 4 / 1e9, 3), equal_to_first_mode=same,
                GBps_payload=round(m * numel * 4 / 1e6 / min(times), 1))
    results.setdefault(mode, line)
    print(json.dumps(line), flush=True)
if 'rows' in results and 'scatter' in results:
    print(json.dumps(dict(what='gist_merge_speedup', scatter_ms=results['scatter']['ms'], rows_ms=results['rows']['ms'],
                          speedup=round(results['scatter']['ms'] / results['rows']['ms'], 2))), flush=True)
