"""Whole-model wiring -- reference ``models/full_graph.py`` (SymGatedGCNModel :9-30,
GatedGCNModel :33-53).  Same constructor signatures, sub-module names and ``state_dict`` keys;
``model(graph, x, e) -> (E, 1)`` fp32 logits in the graph's edge-id order, on the device of ``x``.

In ``model.train()`` mode the forward runs under autograd through ``gnnome_b200.autograd`` (batch-statistics
BatchNorm, hand-written forward / adjoint kernels for the graph primitives); what follows describes ``eval`` mode.

Inputs may live on the CPU (the reference's ``inference.py:388`` forces ``device='cpu'``): they are
moved to the current CUDA device, the graph is staged once (``GraphIndex``, cached on the graph
object) and all edge state stays in dst-sorted position order until the scores are scattered back.
The reference's ``print(x.shape)`` debugging side effect (:25) is not reproduced."""
import torch
import torch.nn as nn

from .. import layers
from ..graph import GraphIndex
from ..layers.encoders import encode_rows, encode_rows2
from ..layers.gated_gcn import state_format


def _to_dev(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class SymGatedGCNModel(nn.Module):
    def __init__(self, node_features, edge_features, hidden_features, hidden_ne_features, num_layers,
                 hidden_edge_scores, normalization, dropout=None):
        super().__init__()
        self.linear1_node = nn.Linear(node_features, hidden_ne_features, bias=True)
        self.linear2_node = nn.Linear(hidden_ne_features, hidden_features, bias=True)
        self.linear1_edge = nn.Linear(edge_features, hidden_ne_features, bias=True)
        self.linear2_edge = nn.Linear(hidden_ne_features, hidden_features, bias=True)
        self.gnn = layers.SymGatedGCN_processor(num_layers, hidden_features, normalization, dropout=dropout)
        self.predictor = layers.ScorePredictor(hidden_features, hidden_edge_scores)
        self.relu = nn.ReLU()

    def forward(self, graph, x, e):
        if self.training or self.gnn.convs[0].normalization == 'layer':
            # train.py: under autograd, batch-statistics BatchNorm; LayerNorm models run on the same primitives
            from ..autograd import model_forward
            return model_forward(self, graph, x, e)
        gi = GraphIndex.from_graph(graph)
        out_dev = x.device
        x_d, e_d = _to_dev(x, gi.device), _to_dev(e, gi.device)
        if state_format(self.linear2_node.out_features) == 'split16' and gi.E > 0:
            h16, h32 = encode_rows2(x_d, None, self.linear1_node, self.linear2_node, gi.N, want32=True)   # :26
            e16, _ = encode_rows2(e_d, gi.in_eid, self.linear1_edge, self.linear2_edge, gi.E)              # :27
            h32, h16, e16 = self.gnn.forward_positions16(gi, h32, h16, e16)                              # :28
            return self.predictor.score_positions16(gi, self.predictor.node_rows16(h16), e16).to(out_dev)  # :29
        h = encode_rows(x_d, None, self.linear1_node, self.linear2_node, gi.N)             # :26
        e_pos = encode_rows(e_d, gi.in_eid, self.linear1_edge, self.linear2_edge, gi.E)    # :27
        h, e_pos = self.gnn.forward_positions(gi, h, e_pos)                               # :28
        return self.predictor.forward_positions(gi, h, e_pos).to(out_dev)                 # :29


class GatedGCNModel(nn.Module):
    def __init__(self, node_features, edge_features, hidden_features, hidden_ne_features, num_layers,
                 hidden_edge_scores, normalization, dropout=None, directed=True):
        super().__init__()
        self.directed = directed
        self.node_encoder = layers.NodeEncoder(node_features, hidden_ne_features, hidden_features)
        self.edge_encoder = layers.EdgeEncoder(edge_features, hidden_ne_features, hidden_features)
        self.gnn = layers.GatedGCN_processor(num_layers, hidden_features, normalization, dropout=dropout)
        self.predictor = layers.ScorePredictor(hidden_features, hidden_edge_scores)

    def forward(self, graph, x, e):
        if self.directed and (self.training or self.gnn.convs[0].normalization == 'layer'):
            from ..autograd import model_forward
            return model_forward(self, graph, x, e)
        if self.training:
            raise NotImplementedError('training GatedGCNModel(directed=False) runs through gnnome_b200.train_dist.ShardedTrainer '
                                      '(one rank is enough on one GPU); model(graph, x, e) covers eval mode')
        gi = GraphIndex.from_graph(graph)
        out_dev = x.device
        x_d, e_d = _to_dev(x, gi.device), _to_dev(e, gi.device)
        if self.directed and state_format(self.node_encoder.linear2.out_features) == 'split16' and gi.E > 0:
            ne, ee = self.node_encoder, self.edge_encoder
            h16, h32 = encode_rows2(x_d, None, ne.linear1, ne.linear2, gi.N, want32=True)
            e16, _ = encode_rows2(e_d, gi.in_eid, ee.linear1, ee.linear2, gi.E)
            h32, h16, e16 = self.gnn.forward_positions16(gi, h32, h16, e16)
            return self.predictor.score_positions16(gi, self.predictor.node_rows16(h16), e16).to(out_dev)
        gi2 = None
        if not self.directed:
            # dgl.add_reverse_edges (:48): edge ids [0,E) = originals, [E,2E) = reversed copies that
            # carry the same input features (:49).  The doubled graph gets its own index.
            gi2 = getattr(gi, '_undirected', None)
            if gi2 is None:
                gi2 = gi._undirected = GraphIndex(torch.cat((gi.src, gi.dst)), torch.cat((gi.dst, gi.src)),
                                                  gi.N, gi.device)
                gi2._feat_idx = (gi2.in_eid[:gi2.E] % max(gi.E, 1)).to(torch.int32).contiguous()
                inv = torch.empty(gi2.E, dtype=torch.int64, device=gi.device)
                inv[gi2.in_eid[:gi2.E].long()] = torch.arange(gi2.E, device=gi.device)
                gi2._orig_rows = inv[gi.in_eid[:gi.E].long()].to(torch.int32).contiguous()
            if state_format(self.node_encoder.linear2.out_features) == 'split16' and gi.E > 0:
                # the same wiring on the split16 / tcgen05 kernels: the edge state of the doubled graph as fp16 (hi, lo)
                # images, the rows of the original edges (:51 e[:E]) gathered as 4-byte words in gi's position order
                from .. import ops
                ne, ee = self.node_encoder, self.edge_encoder
                H = ne.linear2.out_features
                h16, h32 = encode_rows2(x_d, None, ne.linear1, ne.linear2, gi.N, want32=True)
                e16 = encode_rows2(e_d, gi2._feat_idx, ee.linear1, ee.linear2, gi2.E)[0]
                h32, h16, e16 = self.gnn.forward_positions16(gi2, h32, h16, e16)
                words = ops.gather_rows(e16.view(gi2.E, -1).view(torch.float32), gi2._orig_rows)
                e16o = words.view(torch.float16).view(gi.E, 2, H)
                return self.predictor.score_positions16(gi, self.predictor.node_rows16(h16), e16o).to(out_dev)
        h = encode_rows(x_d, None, self.node_encoder.linear1, self.node_encoder.linear2, gi.N)
        if self.directed:                                                                  # :45-46
            e_pos = encode_rows(e_d, gi.in_eid, self.edge_encoder.linear1, self.edge_encoder.linear2, gi.E)
            h, e_pos = self.gnn.forward_positions(gi, h, e_pos)
        else:
            e2 = encode_rows(e_d, gi2._feat_idx, self.edge_encoder.linear1, self.edge_encoder.linear2, gi2.E)
            h, e2 = self.gnn.forward_positions(gi2, h, e2)
            from .. import ops
            e_pos = ops.gather_rows(e2, gi2._orig_rows)                                    # :51 e[:E]
        return self.predictor.forward_positions(gi, h, e_pos).to(out_dev)                 # :52
