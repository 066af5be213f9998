"""Two-layer input MLPs -- reference ``layers/node_encoder.py:5-34`` / ``layers/edge_encoder.py:4-33``."""
import torch
import torch.nn as nn

from .. import ops


def encode_rows(x, idx, lin1, lin2, rows):
    dev = dict(device=x.device, dtype=torch.float32)
    return ops.encode(x, idx, lin1.weight.detach().to(**dev).contiguous(), lin1.bias.detach().to(**dev),
                      lin2.weight.detach().to(**dev).t().contiguous(), lin2.bias.detach().to(**dev), rows)


def encode_rows2(x, idx, lin1, lin2, rows, want16=True, want32=False):
    """Encoder output as split16 images and / or fp32 rows: ``(out16, out32)``."""
    dev = dict(device=x.device, dtype=torch.float32)
    return ops.encode2(x, idx, lin1.weight.detach().to(**dev).contiguous(), lin1.bias.detach().to(**dev),
                       lin2.weight.detach().to(**dev).t().contiguous(), lin2.bias.detach().to(**dev), rows,
                       want16=want16, want32=want32)


class _Encoder(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, bias=True):
        super().__init__()
        if not bias:
            raise NotImplementedError('bias=False encoders are not supported by the CUDA path')
        self.linear1 = nn.Linear(in_channels, hidden_channels, bias=bias)
        self.linear2 = nn.Linear(hidden_channels, out_channels, bias=bias)
        self.relu = nn.ReLU()

    def forward(self, x):
        from ..graph import _cuda_device
        dev = x.device if x.is_cuda else _cuda_device()
        x_d = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        return encode_rows(x_d, None, self.linear1, self.linear2, x_d.shape[0]).to(x.device)


class NodeEncoder(_Encoder):
    pass


class EdgeEncoder(_Encoder):
    pass
