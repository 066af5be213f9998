"""Layer stacks -- reference ``layers/processor.py:9-32``.  The GCN/GAT/SAGE ablation processors
(:35-84) are outside the hot path and not provided."""
import torch
import torch.nn as nn

from .. import ops
from ..graph import GraphIndex
from .gated_gcn import GatedGCN, SymGatedGCN, state_format


class _Processor(nn.Module):
    _layer = SymGatedGCN

    def __init__(self, num_layers, hidden_features, normalization, dropout=None):
        super().__init__()
        self.convs = nn.ModuleList([
            self._layer(hidden_features, hidden_features, normalization, dropout) for _ in range(num_layers)
        ])

    def forward_positions(self, gi, h, e_pos):
        ws = {}
        for conv in self.convs:
            h, e_pos = conv.forward_positions(gi, h, e_pos, ws)
        return h, e_pos

    def forward_positions16(self, gi, h32, h16, e16):
        ws = {}
        for conv in self.convs:
            h32, h16, e16 = conv.forward_positions16(gi, h32, h16, e16, ws)
        return h32, h16, e16

    def forward(self, graph, h, e):
        gi = GraphIndex.from_graph(graph)
        out_dev = h.device
        h_d = h.detach().to(device=gi.device, dtype=torch.float32).contiguous().clone()
        e_d = e.detach().to(device=gi.device, dtype=torch.float32).contiguous()
        if state_format(h_d.shape[1]) == 'split16' and gi.E > 0:
            e16 = ops.split_rows(e_d, gi.in_eid[:gi.E])
            h_d, _, e16 = self.forward_positions16(gi, h_d, ops.split_rows(h_d), e16)
            return h_d.to(out_dev), ops.merge_rows(e16, gi.in_eid[:gi.E]).to(out_dev)
        e_pos = ops.gather_rows(e_d, gi.in_eid[:gi.E])
        h_d, e_pos = self.forward_positions(gi, h_d, e_pos)
        return h_d.to(out_dev), ops.scatter_rows(e_pos, gi.in_eid[:gi.E]).to(out_dev)


class SymGatedGCN_processor(_Processor):
    _layer = SymGatedGCN


class GatedGCN_processor(_Processor):
    _layer = GatedGCN
