"""Drop-in for gcn/gcn.py: the GraphConv-based GCN used by gcn/train.py and
gcn/train_ist.py (graph captured in the constructor, ``forward(features)``,
state-dict keys ``layers.{i}.{weight,bias}``)."""
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .modules import GraphConv, _split_widths


class GCN(nn.Module):
    def __init__(self, g, in_feats, n_hidden, n_classes, n_layers, activation, dropout,
                 use_layernorm=True, split_input=False, split_output=False, num_subnet=1):
        super().__init__()
        self.g = g
        self.layers = nn.ModuleList()
        self.use_layernorm = use_layernorm
        self.split_input = split_input
        self.split_output = split_output
        dims = _split_widths(in_feats, n_hidden, n_classes, n_layers, split_input, split_output,
                             num_subnet)
        for fin, fout in dims[:-1]:
            self.layers.append(GraphConv(fin, fout, activation=activation))
        self.layers.append(GraphConv(*dims[-1]))
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, features):
        h = features
        for i, layer in enumerate(self.layers):
            if i != 0:
                h = self.dropout(h)
            h = layer(self.g, h)
            if i < len(self.layers) - 1 and self.use_layernorm:
                h = ops.tensor_layer_norm(h)   # F.layer_norm(h, h.shape): ONE mean/var over all n*d elements (gcn/gcn.py:65-66)
        return h
