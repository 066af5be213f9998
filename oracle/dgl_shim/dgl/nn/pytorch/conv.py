"""Import-only stubs: ``layers/processor.py:1`` imports these for its ablation processors
(GCN/GAT/SAGE), which are outside the hot path."""
import torch.nn as nn


class _Unsupported(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError('ablation baselines are outside the oracle scope')


class GraphConv(_Unsupported):
    pass


class GATConv(_Unsupported):
    pass


class SAGEConv(_Unsupported):
    pass
