"""Tensor-level wrappers over the C ABI (one function per ``gnb_*`` entry point).

PyTorch is plumbing here: it owns the device memory and the stream; every byte of arithmetic
happens in ``libgnnome_b200.so``."""
import torch

from . import _lib
from .graph import GraphIndex, current_stream_ptr


def _f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f'{name} must be a contiguous float32 CUDA tensor (got {t.dtype}, {t.device})')
    return t.data_ptr()


def _opt(t):
    return None if t is None else t.data_ptr()


def encode(x, idx, W1, b1, W2t, b2, rows, out=None):
    """out[r] = W2 relu(W1 x[idx[r]] + b1) + b2   (idx may be None)."""
    lib = _lib.load()
    hid, in_f = W1.shape
    H = W2t.shape[1]
    if out is None:
        out = torch.empty((rows, H), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.gnb_encode(_f32(x, 'x'), _opt(idx), rows, in_f, hid, H, _f32(W1, 'W1'), _f32(b1, 'b1'),
                                  _f32(W2t, 'W2t'), _f32(b2, 'b2'), _f32(out, 'out'),
                                  current_stream_ptr(x.device)), 'gnb_encode')
    return out


def node_linear(a, Wt, bias, out=None):
    """out = a @ Wt + bias with Wt k-major [K][M]."""
    lib = _lib.load()
    rows, K = a.shape
    M = Wt.shape[1]
    if out is None:
        out = torch.empty((rows, M), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.gnb_node_linear(_f32(a, 'a'), rows, K, _f32(Wt, 'Wt'), _f32(bias, 'bias'), M,
                                       _f32(out, 'out'), out.stride(0), current_stream_ptr(a.device)),
                   'gnb_node_linear')
    return out


def edge_forward(gi: GraphIndex, H, P, We_t, scale_e, shift_e, e, F, carry, flags):
    lib = _lib.load()
    with torch.cuda.device(e.device):
        _lib.check(lib.gnb_edge_forward(gi.ref(), H, _f32(P, 'P'), P.stride(0), _f32(We_t, 'We_t'),
                                        _f32(scale_e, 'scale_e'), _f32(shift_e, 'shift_e'), _f32(e, 'e'),
                                        _f32(F, 'F'), _f32(carry, 'carry'), flags,
                                        current_stream_ptr(e.device)), 'gnb_edge_forward')


def node_update(gi: GraphIndex, H, P, e, F, carry, h_in, scale_h, shift_h, h_out, flags):
    lib = _lib.load()
    with torch.cuda.device(h_in.device):
        _lib.check(lib.gnb_node_update(gi.ref(), H, _f32(P, 'P'), P.stride(0), _f32(e, 'e'), _f32(F, 'F'),
                                       _f32(carry, 'carry'), _f32(h_in, 'h_in'), _f32(scale_h, 'scale_h'),
                                       _f32(shift_h, 'shift_h'), _f32(h_out, 'h_out'), flags,
                                       current_stream_ptr(h_in.device)), 'gnb_node_update')


def score_forward(gi: GraphIndex, H, hs, S, W1e_t, W2, b2, W3, b3, e, scores):
    lib = _lib.load()
    with torch.cuda.device(e.device):
        _lib.check(lib.gnb_score_forward(gi.ref(), H, hs, _f32(S, 'S'), _f32(W1e_t, 'W1e_t'), _f32(W2, 'W2'),
                                         _f32(b2, 'b2'), _f32(W3, 'W3'), _f32(b3, 'b3'), _f32(e, 'e'),
                                         _f32(scores, 'scores'), current_stream_ptr(e.device)),
                   'gnb_score_forward')


def gather_rows(x, idx, out=None):
    lib = _lib.load()
    rows, W = idx.numel(), x.shape[1]
    if out is None:
        out = torch.empty((rows, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.gnb_gather_rows(_f32(x, 'x'), idx.data_ptr(), rows, W, _f32(out, 'out'),
                                       current_stream_ptr(x.device)), 'gnb_gather_rows')
    return out


def scatter_rows(x, idx, out=None):
    lib = _lib.load()
    rows, W = idx.numel(), x.shape[1]
    if out is None:
        out = torch.empty((rows, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.gnb_scatter_rows(_f32(x, 'x'), idx.data_ptr(), rows, W, _f32(out, 'out'),
                                        current_stream_ptr(x.device)), 'gnb_scatter_rows')
    return out
