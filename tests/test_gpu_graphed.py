"""CUDA-graph replayed step == eager step (padding contributes exactly nothing)."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup(seed, h2d):
    from gist_b200 import ClusterIter, SageGCN, synth
    ds = synth.make('reddit', seed=0, device='cpu', scale=0.01, feat_dim=64)
    g = synth.to_gist_graph(ds, device='cuda')
    train_nid = np.nonzero(ds.train_mask.numpy())[0].astype(np.int64)
    psize = int(ds.part.max()) + 1
    random.seed(seed)
    torch.manual_seed(seed)
    it = ClusterIter('', g, psize, 3, train_nid, use_pp=False, h2d=h2d, rng=random.Random(seed))
    model = SageGCN(64, 32, ds.num_classes, 2, F.relu, 0.0, True, False, False, 1, True).cuda()
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
