"""CUDA-graph replayed step == eager step (padding contributes exactly nothing)."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup(seed, h2d, dropout=0.0):
    from gist_b200 import ClusterIter, SageGCN, synth
    ds = synth.make('reddit', seed=0, device='cpu', scale=0.01, feat_dim=64)
    g = synth.to_gist_graph(ds, device='cuda')
    train_nid = np.nonzero(ds.train_mask.numpy())[0].astype(np.int64)
    psize = int(ds.part.max()) + 1
    random.seed(seed)
    torch.manual_seed(seed)
    it = ClusterIter('', g, psize, 3, train_nid, use_pp=False, h2d=h2d, rng=random.Random(seed))
    model = SageGCN(64, 32, ds.num_classes, 2, F.relu, dropout, True, False, False, 1, True).cuda()
    return it, model


@pytest.mark.parametrize('pipeline', [False, True])
@pytest.mark.parametrize('h2d', ['epoch', 'step'])
def test_graphed_matches_eager(h2d, pipeline):
    from gist_b200.graphed import GraphedClusterTrainer
    from gist_b200.train import train_step
    it_e, model_e = _setup(5, 'step')
    it_g, model_g = _setup(5, h2d)
    model_g.load_state_dict(model_e.state_dict())
    opt = torch.optim.Adam(model_e.parameters(), lr=1e-2, weight_decay=5e-4)
    tr = GraphedClusterTrainer(it_g, model_g, 1e-2, 5e-4, h2d=h2d, pipeline=pipeline).capture()
    for p, q in zip(model_g.parameters(), model_e.parameters()):
        assert torch.equal(p, q)                      # capture did not advance training
    model_e.train()
    steps = 0
    for epoch in range(2):
        for cluster in it_e:
            le = float(train_step(model_e, opt, cluster))
            lg = float(tr.step())
            assert abs(le - lg) <= 1e-4 * max(abs(le), 1.0), (steps, le, lg)
            steps += 1
    assert steps == 2 * len(it_e)
    for p, q in zip(model_g.parameters(), model_e.parameters()):
        assert (p - q).abs().max().item() <= 2e-4 * max(q.abs().max().item(), 1.0)
    # the scratch relabel map is clean after replays
    assert (it_g.g._node_map() == -1).all()


def _run_trainer(monkeypatch, fused_tail, pdl, steps, poke_at=None, hidden_seed=5):
    """Losses and final parameters of a pipelined graph trainer with dropout; ``poke_at``: step before which a
    weight is rewritten the way a GIST dispatch does it (K5 scatter through the raw pointer)."""
    from gist_b200 import _lib, graphed, ops
    monkeypatch.setattr(graphed, 'FUSED_TAIL', fused_tail)
    _lib.set_pdl(pdl)
    try:
        ops._drop_stream_counter[0] = 0         # same dropout stream ids (hence masks) for every variant's layers
        it, model = _setup(hidden_seed, 'epoch', dropout=0.2)
        clock = ops.dropout_state(torch.device('cuda', 0))
        clock.step.zero_()
        tr = graphed.GraphedClusterTrainer(it, model, 1e-2, 5e-4, h2d='epoch', pipeline=True).capture()
        assert tr.fused_tail == fused_tail
        clock.step.fill_(100)           # same clock for every variant from here on
        losses = []
        for k in range(steps):
            if k == poke_at:
                W = model.layers[1].linear.weight
                ops.slice_scatter_(W.data, torch.full((W.shape[0], 3), 0.01, device='cuda'), None,
                                   torch.tensor([0, 5, 9], device='cuda'))
                tr.reset_optimizer()
            losses.append(tr.step().clone())
        torch.cuda.synchronize()
        return torch.stack(losses).cpu(), [p.detach().clone().cpu() for p in model.parameters()], tr
    finally:
        _lib.set_pdl(False)


def test_fused_tail_equals_separate_launches(monkeypatch):
    """Adam + weight low halves + clock tick in one launch, loss written by the cross-entropy kernel: the same
    losses and parameters, bit for bit, as the separate copy / tick / split nodes — across an external weight
    rewrite (the persistent low halves are re-split) — with fewer kernel nodes per step."""
    la, pa, ta = _run_trainer(monkeypatch, True, False, 9, poke_at=4)
    lb, pb, tb = _run_trainer(monkeypatch, False, False, 9, poke_at=4)
    assert torch.equal(la, lb), (la, lb)
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)
    assert ta.gist_launches_per_step <= tb.gist_launches_per_step - 2


@pytest.mark.parametrize('fused_tail', [True, False])
def test_programmatic_dependent_launch_changes_nothing(monkeypatch, fused_tail):
    """GIST_PDL: programmatic edges between the kernels of the captured step — same results bit for bit."""
    la, pa, _ = _run_trainer(monkeypatch, fused_tail, True, 8)
    lb, pb, _ = _run_trainer(monkeypatch, fused_tail, False, 8)
    assert torch.equal(la, lb), (la, lb)
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)
