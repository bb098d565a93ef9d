"""CUDA-graph capture of a full-graph training step (the GraphConv trainers of gcn/train.py and
gcn/train_ist.py).  Their step touches the same tensors every epoch — the whole graph, the whole
feature matrix, parameters updated in place — so the step is static and the reference's per-epoch
cost (and this repo's eager one) is Python + kernel launches: ~40 launches per sub-network step on
a 20 k-node graph.  One replay per epoch removes that.

Mirrors the capture discipline of graphed.py: warm up on a side stream, capture into a private
pool, then restore parameters and optimizer state so the capture itself does not advance training.
"""
import torch

from . import _lib


class CapturedStep:
    """``fn()`` runs one complete step (zero_grad -> forward -> loss -> backward -> optimizer step)
    over persistent tensors and returns nothing; results it wants to expose must be copied into
    tensors the caller owns.  ``optimizers``: gist_b200.optim.Adam instances stepped inside fn."""

    def __init__(self, fn, params, optimizers, device, warmup=3):
        self.fn, self.params, self.opts, self.dev = fn, list(params), list(optimizers), torch.device(device)
        self.graph = None
        self.gist_launches = 0
        self._capture(warmup)

    def _opt_state(self):
        out = []
        for o in self.opts:
            for st in o.state.values():
                out += [v for v in st.values() if torch.is_tensor(v)]
            out += list(o._steps.values())
        return out

    def _capture(self, warmup):
        main = torch.cuda.current_stream(self.dev)
        saved_p = [p.detach().clone() for p in self.params]
        had_state = self._opt_state()
        saved_s = [t.clone() for t in had_state]
        # the warm-up steps and the capture run nn.Dropout: put the device generator back afterwards,
        # so a captured run consumes the same random stream as the eager one (train_ist re-captures
        # at every lr change)
        rng = torch.cuda.get_rng_state(self.dev)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(main)
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.fn()
        main.wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(g):
            self.fn()
        self.gist_launches = _lib.launch_count() - l0
        self.graph = g
        with torch.no_grad():
            for p, q in zip(self.params, saved_p):
                p.copy_(q)
            for o in self.opts:
                o.reset_state()                     # state created during warm-up: back to a fresh optimizer
            for t, q in zip(had_state, saved_s):    # state that existed before: back to what it was
                t.copy_(q)
        torch.cuda.synchronize(self.dev)
        torch.cuda.set_rng_state(rng, self.dev)

    def replay(self):
        self.graph.replay()
