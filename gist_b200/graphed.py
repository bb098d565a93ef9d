"""Whole-step CUDA-graph execution of the cluster-batch training step.

A Reddit cluster batch is ~2 k nodes, so the reference's step (and this repo's
eager step) is bound by kernel launches and Python, not by the GPU.  Here ONE
training step — device batch build (K3) -> L+1 SAGE layers forward (K1 + GEMM)
-> cross entropy -> backward (K2 + GEMMs) -> Adam — is captured once into a CUDA
graph and replayed per step.  Batches have different node / edge counts, so the
captured shapes are the maxima over any grouping of parts and every batch is
padded with the builder's `-1` sentinel: pad rows are isolated, have zero
features and a False train mask, hence contribute exactly 0 to the loss and to
every gradient.  Results equal the eager step's (tests/test_gpu_graphed.py).

Callers: bench.py and the trainers; mirrors cluster_gcn_ist_distrib.py:398-417.
"""
import torch

from .train import make_optimizer, masked_cross_entropy

_KEYS = ('feat', 'label', 'train_mask')


class GraphedClusterTrainer:
    def __init__(self, cluster_iter, model, lr, weight_decay, h2d='epoch'):
        assert h2d in ('epoch', 'step')
        self.it, self.model, self.h2d = cluster_iter, model, h2d
        g = cluster_iter.g
        self.dev = g.device
        self.n_pad = cluster_iter.max_batch_nodes()
        self.cap = max(cluster_iter.max_batch_edges(), 1)
        self.nids = torch.full((self.n_pad,), -1, dtype=torch.int64, device=self.dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.opt = make_optimizer(model.parameters(), lr, weight_decay)
        self.graph = None
        self._epoch_dev = None          # [steps, n_pad] device ids (h2d='epoch')
        self._epoch_host = None         # pinned host ids (h2d='step')
        self.i = 0
        self.h2d_bytes = 0
        self.replays = 0
        self.gist_launches_per_step = 0
        g.is_symmetric()                # decided once, outside capture (it syncs)
        self._load_epoch()

    # ------------------------------------------------------------------ data --
    def _load_epoch(self):
        ids = self.it.padded_epoch_ids(self.n_pad)
        if self.h2d == 'epoch':
            self._epoch_dev = ids.to(self.dev)
            self.h2d_bytes += ids.numel() * 8
        else:
            self._epoch_host = ids.pin_memory()
        self.i = 0

    def _stage_ids(self):
        if self.i >= len(self.it):
            self.it.end_epoch()
            self._load_epoch()
        if self.h2d == 'epoch':
            self.nids.copy_(self._epoch_dev[self.i])
        else:
            self.nids.copy_(self._epoch_host[self.i], non_blocking=True)
            self.h2d_bytes += self.n_pad * 8
        self.i += 1

    # ------------------------------------------------------------------ step --
    def _body(self):
        cluster = self.it.g.subgraph(self.nids, col_capacity=self.cap, ndata_keys=_KEYS)
        self.opt.zero_grad(set_to_none=True)
        pred = self.model(cluster)
        loss = masked_cross_entropy(pred, cluster.ndata['label'], cluster.ndata['train_mask'])
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def capture(self):
        """Warm up on a real batch, capture, then restore parameters / optimizer state so the
        capture itself does not advance training."""
        params = list(self.model.parameters())
        saved = [p.detach().clone() for p in params]
        self.model.train()
        self.nids.copy_(self._epoch_dev[0] if self.h2d == 'epoch' else self._epoch_host[0].to(self.dev))
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(3):
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        from . import _lib
        l0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.gist_launches_per_step = _lib.launch_count() - l0   # this library's kernel nodes per replay
        with torch.no_grad():
            for p, q in zip(params, saved):
                p.copy_(q)
        self.reset_optimizer()
        return self

    def reset_optimizer(self):
        """A fresh Adam, as the reference builds at every dispatch (…distrib.py:405-407), but in
        place: moments and step counters are zeroed so the captured graph's pointers stay valid."""
        self.opt.reset_state()

    def step(self):
        """One training step; returns the 0-d device loss tensor (valid until the next step)."""
        if self.graph is None:
            self.capture()
        self._stage_ids()
        self.graph.replay()
        self.replays += 1
        return self.loss
