#!/usr/bin/env python
"""Time the extended-epilogue SpMM (z = dropout([h | mean-agg(h)]) + 3xTF32 low half) and its
transpose on one cluster batch of a synthetic shape, per kernel variant (diagnostic, not a bench).
usage: spmm_batch_bench.py [shape=amazon2m] [d=4096]"""
import os, random, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gist_b200 as gb
from gist_b200 import _lib, ops, synth

shape = sys.argv[1] if len(sys.argv) > 1 else 'amazon2m'
d = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device('cuda', 0)
random.seed(0); torch.manual_seed(0)
ds = synth.make(shape, seed=0, device=dev, scale=1.0)
g = synth.to_gist_graph(ds)
train_nid = torch.nonzero(ds.train_mask).reshape(-1).cpu().numpy().astype(np.int64)
psize = int(ds.part.max().item()) + 1
del ds
it = gb.ClusterIter('', g, psize, 20, train_nid, use_pp=False)
sg = next(iter(it))
n = sg.number_of_nodes()
nnz = int(sg.rowptr[n].item())
print('batch: n=%d nnz=%d d=%d' % (n, nnz, d))
h = torch.randn(n, d, device=dev)
z = ops._padded_empty(n, 2 * d, dev); z_lo = ops._padded_empty(n, 2 * d, dev)
saved = torch.empty(1, dtype=torch.int64, device=dev)
desc = ops.dropout_state(dev).desc(0.2, 3, step_saved=saved)
dz = torch.randn(n, 2 * d, device=dev); dh = torch.empty(n, d, device=dev)
colptr, row = sg.csc()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


B = _lib.SPMM_BG_SHIFT
for name, fl in (('auto', 0), ('narrow', _lib.SPMM_NARROW), ('wide', _lib.SPMM_WIDE),
                 ('narrow+bg3', _lib.SPMM_NARROW | 3 << B), ('narrow+bg6', _lib.SPMM_NARROW | 6 << B),
                 ('wide+bg2', _lib.SPMM_WIDE | 2 << B), ('wide+bg4', _lib.SPMM_WIDE | 4 << B),
                 ('wide+bg8', _lib.SPMM_WIDE | 8 << B)):
    fwd = lambda: ops.spmm_raw(sg.rowptr, sg.col_buffer, n, n, h, z[:, d:], dst_scale=sg.inv_in_degree(),
                               self_out=z[:, :d], out_lo=z_lo[:, d:], self_lo=z_lo[:, :d], drop=desc,
                               drop_col0_out=d, drop_col0_self=0, flags=fl)
    bwd = lambda: ops.spmm_raw(colptr, row, n, n, dz[:, d:], dh, src_scale=sg.inv_in_degree(), addend=dz[:, :d],
                               flags=fl)
    print('%-11s fwd(EX) %.1f us   transpose %.1f us' % (name, timeit(fwd), timeit(bwd)))
