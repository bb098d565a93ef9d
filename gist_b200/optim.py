"""torch.optim.Adam semantics (the optimizer every GIST trainer builds:
cluster_gcn_ist_distrib.py:405-407, gcn/train_ist.py:210, gcn/train.py:89) with the whole
update of all parameter tensors in ONE kernel launch (csrc/fused.cu, adam_multi_kernel).

torch's fused Adam spends ~40 us on the six small tensors of a Reddit sub-GCN (one CTA per
64 K elements); here the chip-wide grid takes a few microseconds.  The step counter lives on
the device so the update can be captured in a CUDA graph and replayed.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, require_cuda, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._counter = None
        self._steps = {}        # id(group) -> device float step counter
        # fused tail of a training step (gist_adam_multi_ex_f32), set by the trainer that owns the step:
        self.lo_map = {}        # parameter -> persistent contiguous tensor receiving tf32_lo(updated parameter)
        self.tick = None        # device int64 incremented by the launch (the dropout clock of the next step)

    def reset_state(self):
        """Zero moments and step counters IN PLACE (a 'fresh' optimizer whose buffers keep their
        addresses: what a captured CUDA graph needs at every GIST re-dispatch)."""
        for st in self.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
        for v in self._steps.values():
            v.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        lib = _lib.load()
        live = [g for g in self.param_groups if any(p.grad is not None for p in g['params'])]
        for group in live:
            ps = [p for p in group['params'] if p.grad is not None]
            dev = ps[0].device
            require_cuda(*ps)
            if self._counter is None:
                self._counter = torch.zeros(1, dtype=torch.int32, device=dev)
            if id(group) not in self._steps:
                self._steps[id(group)] = torch.zeros((), dtype=torch.float32, device=dev)
            step = self._steps[id(group)]
            n = len(ps)
            P, G, M, V, LO = ((ctypes.c_void_p * n)() for _ in range(5))
            N = (ctypes.c_int64 * n)()
            keep = []
            for i, p in enumerate(ps):
                st = self.state[p]
                if 'exp_avg' not in st or st['exp_avg'].shape != p.shape:
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                assert p.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32
                P[i], G[i] = p.data_ptr(), g.data_ptr()
                M[i], V[i] = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
                N[i] = p.numel()
                lo = self.lo_map.get(p)
                if lo is not None:
                    assert lo.is_contiguous() and lo.numel() == p.numel() and lo.dtype == torch.float32
                    LO[i] = lo.data_ptr()
            b1, b2 = group['betas']
            last = group is live[-1]
            check(lib.gist_adam_multi_ex_f32(n, P, G, M, V, N, group['lr'], b1, b2, group['eps'],
                                             group['weight_decay'], _lib.ptr(step), _lib.ptr(self._counter),
                                             LO if self.lo_map else None,
                                             _lib.ptr(self.tick) if last else None, stream_ptr(dev)),
                  'adam_multi_ex_f32')
        return None
