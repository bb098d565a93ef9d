"""The vanilla full-graph GCN trainer of gcn/train.py (config 1: 2-layer GCN, Cora shape) on
the sm_100a kernels: same argument names, same loop (gcn/train.py:86-121)."""
import time
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import ops
from .gcn import GCN
from .graph import GistGraph
from .optim import Adam
from .train import loss_and_backward
from .train_ist import _flag, add_self_loops, evaluate


class GCNTrainer:
    """model / optimizer state of gcn/train.py::main between epochs."""

    def __init__(self, g, features, labels, train_mask, n_classes, args, device=None, use_graph=False):
        assert isinstance(g, GistGraph)
        self.args = args
        self.use_graph = bool(use_graph)        # replay the (static) full-graph step from a CUDA graph
        self._captured, self._captured_lr = None, None
        self.device = torch.device(device) if device is not None else features.device
        self.features, self.labels, self.train_mask = features, labels, train_mask.bool()
        self.model = GCN(g, features.shape[1], args.n_hidden, n_classes, args.n_layers, F.relu, args.dropout,
                         _flag(args.use_layernorm)).to(self.device)                 # train.py:80-83
        self.optimizer = Adam(self.model.parameters(), lr=args.lr, weight_decay=args.weight_decay)   # :87
        self._loss = torch.zeros((), dtype=torch.float32, device=self.device)

    def train_epoch(self, epoch):
        a = self.args
        if getattr(a, 'lr_scheduler', False):                                         # train.py:93-99
            if epoch == int(0.5 * a.n_epochs) or epoch == int(0.75 * a.n_epochs):
                for pg in self.optimizer.param_groups:
                    pg['lr'] = pg['lr'] / 10
        self.model.train()
        if self.use_graph:
            lr = self.optimizer.param_groups[0]['lr']
            if self._captured is None or self._captured_lr != lr:       # lr is a kernel argument of Adam
                from .graph_capture import CapturedStep
                self._captured = CapturedStep(self._step, list(self.model.parameters()), [self.optimizer],
                                              self.device)
                self._captured_lr = lr
            self._captured.replay()
        else:
            self._step()
        return self._loss

    def _step(self):
        self.optimizer.zero_grad(set_to_none=True)
        logits = self.model(self.features)
        loss = loss_and_backward(logits, self.labels, self.train_mask)     # train.py:106-108
        self.optimizer.step()
        self._loss.copy_(loss)


def main(args, data, device='cuda', log=print, eval_every=1, use_graph=False):
    """gcn/train.py::main on an already-loaded dataset (fields as gist_b200.train_ist.main)."""
    device = torch.device(device)
    features = torch.as_tensor(data.features, dtype=torch.float32)
    labels = torch.as_tensor(data.labels).long().to(device)
    masks = [torch.as_tensor(m).bool().to(device) for m in (data.train_mask, data.val_mask, data.test_mask)]
    g = getattr(data, 'graph', None)
    if not isinstance(g, GistGraph):
        src, dst = torch.as_tensor(data.src).long(), torch.as_tensor(data.dst).long()
        n = features.shape[0]
        if _flag(getattr(args, 'self_loop', True)):
            src, dst = add_self_loops(src, dst, n)
        g = GistGraph.from_edges(src, dst, n, device=device)
    features = features.to(device)
    tr = GCNTrainer(g, features, labels, masks[0], data.num_labels, args, device, use_graph=use_graph)
    dur, record = [], []
    for epoch in range(args.n_epochs):
        if epoch >= 3:
            torch.cuda.synchronize(device)
            t0 = time.time()
        loss = tr.train_epoch(epoch)
        if epoch >= 3:
            torch.cuda.synchronize(device)
            dur.append(time.time() - t0)
        if eval_every and (epoch % eval_every == 0 or epoch == args.n_epochs - 1):
            record.append([evaluate(tr.model, features, labels, masks[1]),
                           evaluate(tr.model, features, labels, masks[2])])
            if log:
                log('Epoch {:05d} | Loss {:.4f} | Val Accuracy {:.4f} | Test Accuracy {:.4f}'.format(
                    epoch, float(loss), record[-1][0], record[-1][1]))
    return SimpleNamespace(record=record, trainer=tr, dur=dur)
