"""dgl.transform.metis_partition stand-in.  METIS is third-party and seeded-
random: the node->part ASSIGNMENT is an input of this restatement (parity on
it is unpinned and not attempted).  The assignment is read from
``g.ndata['_part']`` if present, else contiguous equal blocks of node ids."""
import numpy as np
import torch


def metis_partition(g, k, extra_cached_hops=0):
    n = g.number_of_nodes()
    if '_part' in g.ndata:
        part = g.ndata['_part'].long()
    else:
        part = torch.from_numpy((np.arange(n) * k // max(n, 1)).astype(np.int64))
    out = {}
    for p in range(k):
        nids = torch.nonzero(part == p).reshape(-1)
        out[p] = g.subgraph(nids)
    return out
