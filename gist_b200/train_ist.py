"""The single-device GIST trainer of gcn/train_ist.py (config 2: 3-layer GCN, PubMed shape,
8 sub-GCNs on one GPU) as a class, same argument names and the same order of RNG draws.

Reference loop (gcn/train_ist.py:140-300), per epoch:
  * every ``iter_per_site`` epochs: snapshot the full model's state dict, draw the feature
    partitions (``torch.chunk(torch.randperm(n), m)`` per split layer, :150-166), and for every
    sub-network build a fresh narrow GCN (whose constructor consumes the global RNG for its
    Xavier init, exactly as here), load its slice (:176-191), and give it a fresh Adam with the
    step-decayed learning rate (:193-210);
  * every epoch: one full-graph training step per sub-network, in sub-network order (:216-229);
  * at the end of the round (or of training): merge the slices back (:241-285).
Evaluation of the full model (:16-25) runs after every epoch, outside the reference's timer.

The GraphConv forward/backward, whole-tensor layer norm, masked cross entropy, Adam and the
2-D slice gathers/scatters are the sm_100a kernels of csrc/; there is no CPU path.
"""
import time
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import ist_graphconv as IG
from . import ops
from .gcn import GCN
from .graph import GistGraph
from .optim import Adam
from .train import loss_and_backward


def _flag(v):
    """The reference passes booleans as the strings 'True' / 'False' (train_ist.py:42-59)."""
    if isinstance(v, str):
        assert v in ('True', 'False'), ['Only True or False, get ', v]
        return v == 'True'
    return bool(v)


def add_self_loops(src, dst, n):
    """train_ist.py:110-112: drop existing self-loops, then add one per node."""
    keep = src != dst
    loops = torch.arange(n, dtype=src.dtype, device=src.device)
    return torch.cat((src[keep], loops)), torch.cat((dst[keep], loops))


def random_projection(features, num_subnet, seed=None):
    """train_ist.py:70-81: densify with sklearn's GaussianRandomProjection to the largest
    multiple of num_subnet not above the input width (host-side preprocessing, as the reference)."""
    from sklearn import random_projection as rp
    n_components = int(features.shape[-1] / num_subnet) * num_subnet
    tr = rp.GaussianRandomProjection(n_components=n_components, random_state=seed)
    return torch.FloatTensor(tr.fit_transform(features.cpu().numpy() if torch.is_tensor(features) else features))


@torch.no_grad()
def evaluate(model, features, labels, mask):
    """train_ist.py:16-25 without the boolean-index gather."""
    model.eval()
    logits = model(features)
    m = mask.bool()
    pred = logits.argmax(dim=1)
    return float(((pred == labels) & m).sum().item()) / max(int(m.sum().item()), 1)


class ISTGCNTrainer:
    """State of train_ist.main() between epochs.  ``args`` carries the reference's argparse
    names: n_hidden, n_layers, num_subnet, iter_per_site, dropout, lr, weight_decay, n_epochs,
    split_input, split_output, use_layernorm."""

    def __init__(self, g, features, labels, train_mask, n_classes, args, device=None, use_graph=False):
        assert isinstance(g, GistGraph)
        self.args = args
        # use_graph: the m sub-network steps of one epoch are captured ONCE into a CUDA graph (the
        # step is static: whole graph, whole feature matrix) and replayed per epoch; dispatch then
        # loads the new slices into the SAME sub-model / optimizer tensors instead of building new ones
        self.use_graph = bool(use_graph)
        self._captured, self._captured_lr, self._streams = None, None, None
        self.device = torch.device(device) if device is not None else features.device
        self.g = g
        self.features, self.labels = features, labels
        self.train_mask = train_mask.bool()
        self.in_feats, self.n_classes = features.shape[1], n_classes
        self.split_input, self.split_output = _flag(args.split_input), _flag(args.split_output)
        self.use_layernorm = _flag(args.use_layernorm)
        assert (args.n_hidden % args.num_subnet) == 0                    # train_ist.py:62
        if self.split_input:
            assert self.in_feats % args.num_subnet == 0                  # train_ist.py:83
        # train_ist.py:126-130: the full model is built on the host (RNG order) and moved
        self.model = GCN(g, self.in_feats, args.n_hidden, n_classes, args.n_layers, F.relu, args.dropout,
                         self.use_layernorm).to(self.device)
        self.sub_models, self.opt_list, self.model_inputs = [], [], []
        self.main_dict = self.feats_idx = None
        self.last_loss = None
        self._loss = torch.zeros(args.num_subnet, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ one epoch --
    def _lr(self, epoch):
        a = self.args
        lr = a.lr
        if epoch >= int(a.n_epochs * 0.5):
            lr /= 10
        if epoch >= int(a.n_epochs * 0.75):
            lr /= 10
        return lr

    def _dispatch(self, epoch):
        a = self.args
        self.main_dict = {k: v.detach() for k, v in self.model.state_dict().items()}
        self.feats_idx = IG.sample_feature_partitions(self.in_feats, a.n_hidden, a.n_layers, a.num_subnet,
                                                      self.split_input, self.split_output)
        persistent = self.use_graph and len(self.sub_models) == a.num_subnet
        if not persistent:
            self.sub_models, self.opt_list, self.model_inputs = [], [], []
        for s in range(a.num_subnet):
            # the reference builds a fresh narrow GCN per sub-network here; its Xavier init draws
            # from the global RNG before the slice overwrites it, so the draw is kept either way
            sub = GCN(self.g, self.in_feats, a.n_hidden, self.n_classes, a.n_layers, F.relu, a.dropout,
                      self.use_layernorm, self.split_input, self.split_output, a.num_subnet)
            sd = IG.split_state_dict(self.main_dict, self.feats_idx, s, a.n_layers, self.split_input,
                                     self.split_output)
            idx = self.feats_idx[0][s].to(self.device) if self.split_input else None
            if persistent:
                self.sub_models[s].load_state_dict(sd)                  # in place: captured pointers stay valid
                self.opt_list[s].reset_state()                          # a fresh Adam (train_ist.py:207-209)
                for pg in self.opt_list[s].param_groups:
                    pg['lr'] = self._lr(epoch)
                if idx is not None:
                    ops.slice_gather(self.features, None, idx, out=self.model_inputs[s])
                continue
            sub = sub.to(self.device)
            sub.load_state_dict(sd)
            self.sub_models.append(sub)
            self.opt_list.append(Adam(sub.parameters(), lr=self._lr(epoch), weight_decay=a.weight_decay))
            self.model_inputs.append(ops.slice_gather(self.features, None, idx) if idx is not None   # features[:, idx]
                                     else self.features)

    def _merge(self):
        a = self.args
        subs = [{k: v.detach() for k, v in m.state_dict().items()} for m in self.sub_models]
        upd = IG.merge_state_dicts(self.main_dict, self.feats_idx, subs, a.n_layers, self.split_input,
                                   self.split_output)
        self.model.load_state_dict(upd)

    def train_epoch(self, epoch):
        """One iteration of the reference's epoch loop (train_ist.py:140-286), evaluation excluded.
        Returns the last sub-network's loss as a 0-d device tensor (no host sync)."""
        a = self.args
        self.model.eval()                                               # train_ist.py:144
        if epoch % a.iter_per_site == 0:
            self._dispatch(epoch)
        if self.use_graph:
            lr = self.opt_list[0].param_groups[0]['lr']
            if self._captured is None or self._captured_lr != lr:       # lr is a kernel argument of Adam
                from .graph_capture import CapturedStep
                params = [p for sub in self.sub_models for p in sub.parameters()]
                self._captured = CapturedStep(self._sub_steps, params, self.opt_list, self.device)
                self._captured_lr = lr
            self._captured.replay()
        else:
            self._sub_steps()
        if (epoch + 1) % a.iter_per_site == 0 or epoch == a.n_epochs - 1:
            self._merge()
        self.last_loss = self._loss[a.num_subnet - 1]
        return self.last_loss

    def _sub_steps(self):
        """train_ist.py:212-229 for every sub-network.  The m sub-networks of one epoch share only
        read-only inputs (graph, features, labels) — parameters, gradients and optimizer state are
        disjoint — so in graph mode each runs on its own stream: m parallel branches of the captured
        graph instead of a chain of m x ~60 small kernels.  Results are independent of the overlap."""
        m = self.args.num_subnet
        if not self.use_graph:
            for s in range(m):
                self._sub_step(s)
            return
        main = torch.cuda.current_stream(self.device)
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(m)]
        for s in range(m):
            st = self._streams[s]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                self._sub_step(s)
        for st in self._streams:
            main.wait_stream(st)

    def _sub_step(self, s):
        sub, opt = self.sub_models[s], self.opt_list[s]
        opt.zero_grad(set_to_none=True)
        sub.train()
        logits = sub(self.model_inputs[s])
        loss = loss_and_backward(logits, self.labels, self.train_mask)     # train_ist.py:220-222
        opt.step()
        self._loss[s].copy_(loss)


def main(args, data, device='cuda', log=print, eval_every=1, use_graph=False):
    """train_ist.main(args) on an already-loaded dataset (``data``: graph edges ``src``/``dst`` or
    a GistGraph ``graph``, ``features``, ``labels``, ``train_mask``/``val_mask``/``test_mask``,
    ``num_labels`` — the fields of DGL's citation datasets the reference reads, :66-92).
    Returns the per-epoch [val, test] accuracy record and the trainer."""
    device = torch.device(device)
    use_ist = _flag(getattr(args, 'use_ist', True))
    if not use_ist:
        raise NotImplementedError('Should train with IST')             # train_ist.py:288
    features = data.features
    if _flag(getattr(args, 'use_random_proj', False)):
        features = random_projection(features, args.num_subnet)
    features = torch.as_tensor(features, dtype=torch.float32)
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
    tr = ISTGCNTrainer(g, features, labels, masks[0], data.num_labels, args, device, use_graph=use_graph)
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
            acc_val = evaluate(tr.model, features, labels, masks[1])
            acc_test = evaluate(tr.model, features, labels, masks[2])
            record.append([acc_val, acc_test])
            if log:
                log('Epoch {:05d} | Time(s) {:.4f} | Loss {:.4f} | Val Accuracy {:.4f} | Test Accuracy {:.4f} |'
                    'ETputs(KTEPS) {:.2f}'.format(epoch, sum(dur) / max(len(dur), 1), float(loss), acc_val, acc_test,
                                                  g.number_of_edges() / max(sum(dur) / max(len(dur), 1), 1e-9) / 1000))
    return SimpleNamespace(record=record, trainer=tr, dur=dur)
