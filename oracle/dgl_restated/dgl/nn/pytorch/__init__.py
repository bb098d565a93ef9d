"""dgl.nn.pytorch.GraphConv (0.5.x, norm='both') restated — SURVEY.md Appendix A."""
import torch
from torch import nn


class GraphConv(nn.Module):
    def __init__(self, in_feats, out_feats, norm='both', weight=True, bias=True,
                 activation=None, allow_zero_in_degree=False):
        super().__init__()
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats)) if weight else None
        self.bias = nn.Parameter(torch.Tensor(out_feats)) if bias else None
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, feat):
        from ... import DGLError, function as fn
        graph = graph.local_var()
        if not self._allow_zero_in_degree and (graph.in_degrees() == 0).any():
            raise DGLError('There are 0-in-degree nodes in the graph, output for those nodes '
                           'will be invalid. Add self-loops or set allow_zero_in_degree.')
        if self._norm == 'both':
            degs = graph.out_degrees().to(feat.device).float().clamp(min=1)
            norm = torch.pow(degs, -0.5)
            feat = feat * norm.reshape((-1,) + (1,) * (feat.dim() - 1))
        weight = self.weight
        if self._in_feats > self._out_feats:
            # mult W first to reduce the feature size for aggregation
            if weight is not None:
                feat = torch.matmul(feat, weight)
            graph.ndata['h'] = feat
            graph.update_all(fn.copy_src(src='h', out='m'), fn.sum(msg='m', out='h'))
            rst = graph.ndata['h']
        else:
            graph.ndata['h'] = feat
            graph.update_all(fn.copy_src(src='h', out='m'), fn.sum(msg='m', out='h'))
            rst = graph.ndata['h']
            if weight is not None:
                rst = torch.matmul(rst, weight)
        if self._norm != 'none':
            degs = graph.in_degrees().to(feat.device).float().clamp(min=1)
            norm = torch.pow(degs, -0.5) if self._norm == 'both' else 1.0 / degs
            rst = rst * norm.reshape((-1,) + (1,) * (feat.dim() - 1))
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst
