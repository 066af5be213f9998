"""Edge-score MLP -- reference ``layers/score_predictor.py:5-24``."""
import torch
import torch.nn as nn

from .. import ops
from ..graph import GraphIndex


class ScorePredictor(nn.Module):
    def __init__(self, in_features, hidden_edge_scores):
        super().__init__()
        self.W1 = nn.Linear(3 * in_features, hidden_edge_scores)
        self.W2 = nn.Linear(hidden_edge_scores, 32)
        self.W3 = nn.Linear(32, 1)
        self.in_features, self.hidden_edge_scores = in_features, hidden_edge_scores

    def node_rows(self, x):
        """S[n] = [x_n W1s^T | x_n W1d^T + b1]  ([N][2 hs]): the node halves of W1 (:13-14), once per node."""
        H, hs = self.in_features, self.hidden_edge_scores
        dev = dict(device=x.device, dtype=torch.float32)
        W1 = self.W1.weight.detach().to(**dev)                      # [hs][3H] = [W1s | W1d | W1e]
        Ws_t = torch.cat((W1[:, :H], W1[:, H:2 * H]), dim=0).t().contiguous()   # [H][2hs]
        bias = torch.cat((torch.zeros(hs, **dev), self.W1.bias.detach().to(**dev)))
        return ops.node_linear(x, Ws_t, bias)

    def score_positions(self, gi: GraphIndex, S, e_pos, scores=None):
        """Per-edge part: S rows of both endpoints + e W1e^T -> relu -> W2 -> relu -> W3 (:14-16)."""
        H, hs = self.in_features, self.hidden_edge_scores
        dev = dict(device=e_pos.device, dtype=torch.float32)
        W1e_t = self.W1.weight.detach().to(**dev)[:, 2 * H:].t().contiguous()   # [H][hs]
        if scores is None:
            scores = torch.empty((gi.E, 1), **dev)
        ops.score_forward(gi, H, hs, S, W1e_t, self.W2.weight.detach().to(**dev).contiguous(),
                          self.W2.bias.detach().to(**dev), self.W3.weight.detach().to(**dev).reshape(-1).contiguous(),
                          self.W3.bias.detach().to(**dev), e_pos, scores)
        return scores

    def _node_weights(self, dev):
        H, hs = self.in_features, self.hidden_edge_scores
        W1 = self.W1.weight.detach().to(**dev)
        Wn = torch.cat((W1[:, :H], W1[:, H:2 * H]), dim=0).contiguous()          # [2hs][H] (nn.Linear layout)
        bias = torch.cat((torch.zeros(hs, **dev), self.W1.bias.detach().to(**dev)))
        return Wn, bias

    def node_rows16(self, x16):
        """``node_rows`` from the split16 images of x, on the tensor cores."""
        dev = dict(device=x16.device, dtype=torch.float32)
        Wn, bias = self._node_weights(dev)
        return ops.node_linear_tc2(x16, ops.pack_linear_tc(Wn), bias, Wn.shape[0])

    def score_positions16(self, gi: GraphIndex, S, e16, scores=None, tensor_cores=True):
        """``score_positions`` with the edge state given as split16 images; ``tensor_cores=False`` runs the
        CUDA-core cross-check kernel."""
        H, hs = self.in_features, self.hidden_edge_scores
        dev = dict(device=e16.device, dtype=torch.float32)
        W1e = self.W1.weight.detach().to(**dev)[:, 2 * H:].contiguous()         # [hs][H]
        if scores is None:
            scores = torch.empty((gi.E, 1), **dev)
        tail = (self.W2.weight.detach().to(**dev).contiguous(), self.W2.bias.detach().to(**dev),
                self.W3.weight.detach().to(**dev).reshape(-1).contiguous(), self.W3.bias.detach().to(**dev), e16, scores)
        if tensor_cores:
            ops.score_forward_tc2(gi, H, hs, S, ops.pack_linear_tc(W1e), *tail)
        else:
            ops.score_forward2(gi, H, hs, S, W1e.t().contiguous(), *tail)
        return scores

    def forward_positions(self, gi: GraphIndex, x, e_pos):
        """Scores in ORIGINAL edge-id order, shape (E, 1), from position-ordered edge rows."""
        return self.score_positions(gi, self.node_rows(x), e_pos)

    def forward(self, graph, x, e):
        gi = GraphIndex.from_graph(graph)
        out_dev = x.device
        x_d = x.detach().to(device=gi.device, dtype=torch.float32).contiguous()
        e_d = e.detach().to(device=gi.device, dtype=torch.float32).contiguous()
        e_pos = ops.gather_rows(e_d, gi.in_eid[:gi.E])
        return self.forward_positions(gi, x_d, e_pos).to(out_dev)
