#!/usr/bin/env python
"""How far does an fp32 evaluation of the REFERENCE arithmetic itself sit from fp64?  Runs the oracle
restatement of the SAGE GCN (cluster_gcn/modules.py) in fp32 and in fp64 on the module-test
configurations and prints the norm-wise relative spread of logits and gradients: the floor any
fp32 implementation (the reference's own included) has against the fp64 oracle.  CPU only.
With --gpu also runs this repo's CUDA modules on the same inputs (both GEMM back ends)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gist_oracle as O  # noqa: E402
from tests.util import random_graph  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    gpu = '--gpu' in sys.argv
    cfgs = [(602, 256, 41, 2, True, 800, 12000), (100, 64, 47, 3, True, 800, 12000), (50, 32, 5, 1, False, 800, 12000),
            (602, 256, 41, 2, True, 2600, 110001)]
    for fin, hid, ncls, L, ln, n, nnz in cfgs:
        src, dst = random_graph(n, nnz, seed=L + nnz % 7)
        og = O.OGraph(src, dst, n)
        torch.manual_seed(0)
        dims = [(hid, 2 * fin)] + [(hid, 2 * hid)] * (L - 1) + [(ncls, 2 * hid)]
        params32 = []
        for o, i in dims:
            s = 1.0 / (i ** 0.5)
            params32.append((torch.empty(o, i).uniform_(-s, s), torch.empty(o).uniform_(-s, s)))
        x = torch.randn(n, fin)
        y = torch.randint(0, ncls, (n,))
        res = {}
        outs = {}
        for name, dt in (('fp64', torch.float64), ('fp32_cpu_oracle', torch.float32)):
            ps = [(w.to(dt).requires_grad_(True), b.to(dt).requires_grad_(True)) for w, b in params32]
            out = O.sage_gcn_forward(og, x.to(dt), ps, ln)
            F.cross_entropy(out, y).backward()
            outs[name] = (out, ps)
        ref_out, ref_ps = outs['fp64']
        o32, p32 = outs['fp32_cpu_oracle']
        res['fp32_cpu_oracle'] = dict(logits=rel(o32, ref_out), dW=max(rel(a[0].grad, b[0].grad) for a, b in zip(p32, ref_ps)),
                                      db=max(rel(a[1].grad, b[1].grad) for a, b in zip(p32, ref_ps)))
        if gpu:
            from gist_b200 import GistGraph, SageGCN, ops
            g = GistGraph.from_edges(src, dst, n, device='cuda')
            g.ndata['feat'] = x.cuda()
            for mode in ('3xtf32', 'fp32'):
                ops.set_matmul_precision(mode)
                model = SageGCN(fin, hid, ncls, L, F.relu, 0.0, ln, False, False, 1, True).cuda()
                with torch.no_grad():
                    for l, (w, b) in zip(model.layers, params32):
                        l.linear.weight.copy_(w)
                        l.linear.bias.copy_(b)
                out = model(g)
                F.cross_entropy(out, y.cuda()).backward()
                res['gist_' + mode] = dict(logits=rel(out, ref_out),
                                           dW=max(rel(l.linear.weight.grad, r[0].grad) for l, r in zip(model.layers, ref_ps)),
                                           db=max(rel(l.linear.bias.grad, r[1].grad) for l, r in zip(model.layers, ref_ps)))
            ops.set_matmul_precision(ops.DEFAULT_MATMUL_PRECISION)
        print(json.dumps({'cfg': dict(fin=fin, hid=hid, ncls=ncls, L=L, ln=ln, n=n, nnz=nnz),
                          **{k: {kk: float('%.3g' % vv) for kk, vv in v.items()} for k, v in res.items()}}))


if __name__ == '__main__':
    main()
