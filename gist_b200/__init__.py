"""gist_b200 — B200-native (sm_100a) implementation of GIST's aggregation hot path
behind the reference's PyTorch module API.  See DESIGN.md."""
from . import function  # noqa: F401
from .graph import GistGraph, GistError, NID  # noqa: F401
from .modules import (GraphConv, GraphSAGELayer, ISTSAGELayer, GraphSAGE, GCN as SageGCN,  # noqa: F401
                      BaselineGCN, GATLayer, MultiHeadGATLayer, GAT)
from .gcn import GCN  # noqa: F401
from .sampler import ClusterIter, get_partition_list, get_subgraph  # noqa: F401
from .ist import create_partition, DistributedGNNWrapper  # noqa: F401
from .ist_gat import DistributedGATWrapper  # noqa: F401

__all__ = ['function', 'GistGraph', 'GistError', 'NID', 'GraphConv', 'GraphSAGELayer', 'ISTSAGELayer',
           'GraphSAGE', 'SageGCN', 'BaselineGCN', 'GATLayer', 'MultiHeadGATLayer', 'GAT', 'GCN', 'ClusterIter', 'get_partition_list',
           'get_subgraph', 'create_partition', 'DistributedGNNWrapper', 'DistributedGATWrapper']
