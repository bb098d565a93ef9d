"""get_data (cluster_gcn_ist_distrib.py:484-518) end to end on files in the reference's Amazon2M
format: load -> standardise -> device graph -> ClusterIter -> one training step."""
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.test_datasets import _write_amazon

pytestmark = pytest.mark.gpu


def test_get_data_feeds_the_cluster_trainer(tmp_path):
    import gist_b200 as gb
    from gist_b200 import datasets
    from gist_b200.train import make_optimizer, train_step
    from oracle import datasets_oracle as O
    d = str(tmp_path)
    _write_amazon(d, n=600, m=6000, seed=5, list_labels=False, str_ids=False)
    ref = O.amazon_process(d)
    random.seed(0)
    torch.manual_seed(0)
    args = SimpleNamespace(dataset='amazon2m', psize=12, batch_size=3, normalize=True, use_pp=False)
    # dn='' semantics are ClusterIter's; get_data passes the dataset name, so keep its cache out of ../data
    import gist_b200.sampler as S
    orig = S._partition.cache_path
    S._partition.cache_path = lambda dn, psize, cache_dir: str(tmp_path / ('%s_%d.npy' % (dn, psize)))
    try:
        (g, it, train_mask, val_mask, test_mask, labels, train_nid, in_feats, n_classes,
         n_edges) = datasets.get_data(args, torch.device('cuda'), raw_dir=d)
    finally:
        S._partition.cache_path = orig
    assert g.device.type == 'cuda' and in_feats == 7 and n_classes == ref['num_classes']
    assert n_edges == len(ref['src'])
    assert np.array_equal(train_nid, np.nonzero(ref['train_mask'])[0])
    assert val_mask.is_cuda and test_mask.is_cuda and not labels.is_cuda
    assert len(it) == 4
    # the loader standardised once (AmazonDataset.py:87-92) and get_data again on the training rows
    # (…distrib.py:493-499): training-row moments are (0, 1) up to the constant column
    f = g.ndata['feat'][torch.from_numpy(ref['train_mask']).cuda()]
    assert f.mean(0).abs().max().item() < 1e-4
    model = gb.SageGCN(in_feats, 16, n_classes, 1, F.relu, 0.0, True, False, False, 1, True).cuda()
    opt = make_optimizer(model.parameters(), 1e-2, 0.0)
    seen = 0
    losses = []
    for cluster in it:
        assert cluster.ndata['train_mask'].all()              # every batch row is a training node
        losses.append(float(train_step(model, opt, cluster)))
        seen += cluster.number_of_nodes()
    assert seen == len(train_nid) and all(np.isfinite(losses))
