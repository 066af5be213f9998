"""Tensor-level wrappers over the C ABI (one function per ``gnb_*`` entry point).

PyTorch is plumbing here: it owns the device memory and the stream; every byte of arithmetic
happens in ``libgnnome_b200.so``."""
import os
import sys

import torch

from . import _lib
from .graph import GraphIndex, current_stream_ptr


class LaunchLog:
    """Opt-in accounting of the library's kernel launches (used by bench.py): counts every C-ABI
    compute call and, when ``timing`` is on, brackets it with CUDA events on the launching stream
    so that per-kernel device time can be read back after a synchronize."""
    enabled = False
    timing = False
    counts = {}
    events = {}

    @classmethod
    def reset(cls, enabled=True, timing=False):
        cls.enabled, cls.timing, cls.counts, cls.events = enabled, timing, {}, {}

    @classmethod
    def total(cls):
        return sum(cls.counts.values())

    @classmethod
    def times_ms(cls):
        """name -> list of per-launch milliseconds (call after torch.cuda.synchronize())."""
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in cls.events.items()}


_TRACE = bool(os.environ.get('GNB_TRACE'))   # debugging: print every C-ABI call and synchronise after it


class _logged:
    def __init__(self, name, device, launches=1):
        self.name, self.device, self.launches = name, device, launches

    def __enter__(self):
        self.dev_ctx = torch.cuda.device(self.device)
        self.dev_ctx.__enter__()
        if _TRACE:
            print(f'[gnb] {self.name} ...', file=sys.stderr, flush=True)
        if LaunchLog.enabled:
            LaunchLog.counts[self.name] = LaunchLog.counts.get(self.name, 0) + self.launches
            if LaunchLog.timing:
                self.start = torch.cuda.Event(enable_timing=True)
                self.start.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if _TRACE and exc[0] is None:
            torch.cuda.synchronize(self.device)
            print(f'[gnb] {self.name} done', file=sys.stderr, flush=True)
        if LaunchLog.enabled and LaunchLog.timing and exc[0] is None:
            end = torch.cuda.Event(enable_timing=True)
            end.record(torch.cuda.current_stream(self.device))
            LaunchLog.events.setdefault(self.name, []).append((self.start, end))
        return self.dev_ctx.__exit__(*exc)


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
    with _logged('gnb_encode', x.device):
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
    with _logged('gnb_node_linear', a.device):
        _lib.check(lib.gnb_node_linear(_f32(a, 'a'), rows, K, _f32(Wt, 'Wt'), _f32(bias, 'bias'), M,
                                       _f32(out, 'out'), out.stride(0), current_stream_ptr(a.device)),
                   'gnb_node_linear')
    return out


def pack_linear_tc(W):
    """nn.Linear weight [M][K] -> packed fp16 (hi, lo) blocks for the tensor-core kernels."""
    lib = _lib.load()
    M, K = W.shape
    out = torch.empty(lib.gnb_packed_linear_bytes(M, K), dtype=torch.uint8, device=W.device)
    with _logged('gnb_pack_linear_tc', W.device):
        _lib.check(lib.gnb_pack_linear_tc(_f32(W, 'W'), M, K, out.data_ptr(), current_stream_ptr(W.device)),
                   'gnb_pack_linear_tc')
    return out


def edge_forward(gi: GraphIndex, H, P, We_t, scale_e, shift_e, e, F, carry, flags):
    lib = _lib.load()
    with _logged('gnb_edge_forward', e.device):
        _lib.check(lib.gnb_edge_forward(gi.ref(), H, _f32(P, 'P'), P.stride(0), _f32(We_t, 'We_t'),
                                        _f32(scale_e, 'scale_e'), _f32(shift_e, 'shift_e'), _f32(e, 'e'),
                                        _f32(F, 'F'), _f32(carry, 'carry'), flags,
                                        current_stream_ptr(e.device)), 'gnb_edge_forward')


def node_update(gi: GraphIndex, H, P, e, F, carry, h_in, scale_h, shift_h, h_out, flags, chunk,
                node_begin=0, node_end=None, xp_ptr=None, xp_row=None, xp_buf=None):
    lib = _lib.load()
    node_end = gi.N if node_end is None else node_end
    with _logged('gnb_node_update', h_in.device):
        _lib.check(lib.gnb_node_update(gi.ref(), H, _f32(P, 'P'), P.stride(0), _f32(e, 'e'), _f32(F, 'F'),
                                       _f32(carry, 'carry'), _f32(h_in, 'h_in'), _f32(scale_h, 'scale_h'),
                                       _f32(shift_h, 'shift_h'), _f32(h_out, 'h_out'), flags, chunk,
                                       node_begin, node_end, _opt(xp_ptr), _opt(xp_row),
                                       None if xp_buf is None else _f32(xp_buf, 'xp_buf'),
                                       current_stream_ptr(h_in.device)), 'gnb_node_update')


def reverse_partial(gi: GraphIndex, H, P, e, node_begin, node_end, out):
    """out[i - node_begin] = (sum sigma * A3h[dst] | sum sigma) over the local out-edges of node i."""
    lib = _lib.load()
    with _logged('gnb_reverse_partial', e.device):
        _lib.check(lib.gnb_reverse_partial(gi.ref(), H, _f32(P, 'P'), P.stride(0), _f32(e, 'e'), node_begin, node_end,
                                           _f32(out, 'out'), current_stream_ptr(e.device)), 'gnb_reverse_partial')


def score_forward(gi: GraphIndex, H, hs, S, W1e_t, W2, b2, W3, b3, e, scores):
    lib = _lib.load()
    with _logged('gnb_score_forward', e.device):
        _lib.check(lib.gnb_score_forward(gi.ref(), H, hs, _f32(S, 'S'), _f32(W1e_t, 'W1e_t'), _f32(W2, 'W2'),
                                         _f32(b2, 'b2'), _f32(W3, 'W3'), _f32(b3, 'b3'), _f32(e, 'e'),
                                         _f32(scores, 'scores'), current_stream_ptr(e.device)),
                   'gnb_score_forward')


def gather_rows(x, idx, out=None):
    """out[r] = x[idx[r]]; ``x`` may be a column block of a wider table (unit column stride)."""
    lib = _lib.load()
    rows, W = idx.numel(), x.shape[1]
    if out is None:
        out = torch.empty((rows, W), dtype=torch.float32, device=x.device)
    if x.dtype != torch.float32 or not x.is_cuda or x.stride(1) != 1:
        raise ValueError('x must be a float32 CUDA matrix with unit column stride')
    with _logged('gnb_gather_rows', x.device):
        _lib.check(lib.gnb_gather_rows_ld(x.data_ptr(), x.stride(0), idx.data_ptr(), rows, W, _f32(out, 'out'),
                                          out.stride(0), current_stream_ptr(x.device)), 'gnb_gather_rows_ld')
    return out


def scatter_rows(x, idx, out=None):
    lib = _lib.load()
    rows, W = idx.numel(), x.shape[1]
    if out is None:
        out = torch.empty((rows, W), dtype=torch.float32, device=x.device)
    with _logged('gnb_scatter_rows', x.device):
        _lib.check(lib.gnb_scatter_rows(_f32(x, 'x'), idx.data_ptr(), rows, W, _f32(out, 'out'),
                                        current_stream_ptr(x.device)), 'gnb_scatter_rows')
    return out


# ------------------------------------------------------------------------------------------------
# split16 path (TMA-fed tcgen05 kernels): state = fp16 [rows][2][K] (row = hi | lo), x = 16 * (hi + lo)
# ------------------------------------------------------------------------------------------------
def _img(t, name):
    if t.dtype != torch.float16 or not t.is_cuda or not t.is_contiguous() or t.ndim != 3 or t.shape[1] != 2:
        raise ValueError(f'{name} must be a contiguous float16 CUDA tensor of shape [rows][2][K] (got {t.dtype}, '
                         f'{tuple(t.shape)}, {t.device})')
    return t.data_ptr()


def empty_split16(rows, K, device):
    return torch.empty((rows, 2, K), dtype=torch.float16, device=device)


def split_rows(x, idx=None, out=None, scale=None):
    """fp32 rows -> split16 images; ``out[r] = split(x[idx[r]])`` when ``idx`` is given; ``scale``: optional 0-dim float32
    CUDA tensor multiplied in before the split (a power of two, see the header)."""
    lib = _lib.load()
    rows = x.shape[0] if idx is None else idx.numel()
    K = x.shape[1]
    if out is None:
        out = empty_split16(rows, K, x.device)
    with _logged('gnb_split_rows', x.device):
        _lib.check(lib.gnb_split_rows(_f32(x, 'x'), _opt(idx), rows, K, _img(out, 'out'), _opt(scale),
                                      current_stream_ptr(x.device)), 'gnb_split_rows')
    return out


def merge_rows(x16, idx=None, out=None):
    """split16 images -> fp32 rows; ``out[idx[r]] = merge(x16[r])`` when ``idx`` is given."""
    lib = _lib.load()
    rows, _, K = x16.shape
    if out is None:
        out = torch.empty((rows, K), dtype=torch.float32, device=x16.device)
    with _logged('gnb_merge_rows', x16.device):
        _lib.check(lib.gnb_merge_rows(_img(x16, 'x16'), _opt(idx), rows, K, _f32(out, 'out'),
                                      current_stream_ptr(x16.device)), 'gnb_merge_rows')
    return out


def encode2(x, idx, W1, b1, W2t, b2, rows, want16=True, want32=False):
    """Two-layer encoder writing split16 images and / or fp32 rows: returns ``(out16, out32)``."""
    lib = _lib.load()
    hid, in_f = W1.shape
    H = W2t.shape[1]
    out16 = empty_split16(rows, H, x.device) if want16 else None
    out32 = torch.empty((rows, H), dtype=torch.float32, device=x.device) if want32 else None
    with _logged('gnb_encode2', x.device):
        _lib.check(lib.gnb_encode2(_f32(x, 'x'), _opt(idx), rows, in_f, hid, H, _f32(W1, 'W1'), _f32(b1, 'b1'),
                                   _f32(W2t, 'W2t'), _f32(b2, 'b2'), _opt(out16), _opt(out32),
                                   current_stream_ptr(x.device)), 'gnb_encode2')
    return out16, out32


def node_linear_tc2(x16, Wp, bias, M, out=None, out_scale=None):
    """out = out_scale * (X @ W.T) + bias with X in split16 format; Wp = pack_linear_tc(W[M][K]); ``out_scale``: optional
    0-dim float32 CUDA tensor."""
    lib = _lib.load()
    rows, _, K = x16.shape
    if out is None:
        out = torch.empty((rows, M), dtype=torch.float32, device=x16.device)
    with _logged('gnb_node_linear_tc2', x16.device):
        _lib.check(lib.gnb_node_linear_tc2(_img(x16, 'x16'), rows, K, Wp.data_ptr(), _f32(bias, 'bias'), M,
                                           _f32(out, 'out'), out.stride(0), _opt(out_scale),
                                           current_stream_ptr(x16.device)), 'gnb_node_linear_tc2')
    return out


def edge_forward_tc2(gi: GraphIndex, H, P, Wp, e16, F, carry, flags):
    """The fused edge pass on the split16 state; the norm affine is folded into ``Wp`` and the B1h / B2h blocks of
    ``P`` by the caller (see the header)."""
    lib = _lib.load()
    with _logged('gnb_edge_forward_tc2', e16.device):
        _lib.check(lib.gnb_edge_forward_tc2(gi.ref(), H, _f32(P, 'P'), P.stride(0), Wp.data_ptr(), _img(e16, 'e16'),
                                            _f32(F, 'F'), _f32(carry, 'carry'), flags,
                                            current_stream_ptr(e16.device)), 'gnb_edge_forward_tc2')


def node_update2(gi: GraphIndex, H, P, e16, F, carry, h_in, scale_h, shift_h, h_out, h16_out, flags, chunk,
                 node_begin=0, node_end=None, xp_ptr=None, xp_row=None, xp_buf=None):
    lib = _lib.load()
    node_end = gi.N if node_end is None else node_end
    with _logged('gnb_node_update2', h_in.device):
        _lib.check(lib.gnb_node_update2(gi.ref(), H, _f32(P, 'P'), P.stride(0), _img(e16, 'e16'), _f32(F, 'F'),
                                        _f32(carry, 'carry'), _f32(h_in, 'h_in'), _f32(scale_h, 'scale_h'),
                                        _f32(shift_h, 'shift_h'), _f32(h_out, 'h_out'),
                                        None if h16_out is None else _img(h16_out, 'h16_out'), flags, chunk,
                                        node_begin, node_end, _opt(xp_ptr), _opt(xp_row),
                                        None if xp_buf is None else _f32(xp_buf, 'xp_buf'),
                                        current_stream_ptr(h_in.device)), 'gnb_node_update2')


def reverse_partial2(gi: GraphIndex, H, P, e16, node_begin, node_end, out):
    lib = _lib.load()
    with _logged('gnb_reverse_partial2', e16.device):
        _lib.check(lib.gnb_reverse_partial2(gi.ref(), H, _f32(P, 'P'), P.stride(0), _img(e16, 'e16'), node_begin,
                                            node_end, _f32(out, 'out'), current_stream_ptr(e16.device)),
                   'gnb_reverse_partial2')


def score_forward2(gi: GraphIndex, H, hs, S, W1e_t, W2, b2, W3, b3, e16, scores):
    lib = _lib.load()
    with _logged('gnb_score_forward2', e16.device):
        _lib.check(lib.gnb_score_forward2(gi.ref(), H, hs, _f32(S, 'S'), _f32(W1e_t, 'W1e_t'), _f32(W2, 'W2'),
                                          _f32(b2, 'b2'), _f32(W3, 'W3'), _f32(b3, 'b3'), _img(e16, 'e16'),
                                          _f32(scores, 'scores'), current_stream_ptr(e16.device)),
                   'gnb_score_forward2')


def score_forward_tc2(gi: GraphIndex, H, hs, S, Wp, W2, b2, W3, b3, e16, scores):
    lib = _lib.load()
    with _logged('gnb_score_forward_tc2', e16.device):
        _lib.check(lib.gnb_score_forward_tc2(gi.ref(), H, hs, _f32(S, 'S'), Wp.data_ptr(), _f32(W2, 'W2'),
                                             _f32(b2, 'b2'), _f32(W3, 'W3'), _f32(b3, 'b3'), _img(e16, 'e16'),
                                             _f32(scores, 'scores'), current_stream_ptr(e16.device)),
                   'gnb_score_forward_tc2')


def degree_rows(gi, swap=False):
    """(N, 2) float32: ``(in_degree, out_degree)`` of every node, ``(out, in)`` with ``swap`` (train.py:117-118)."""
    lib = _lib.load()
    x = torch.empty((gi.N, 2), dtype=torch.float32, device=gi.device)
    with _logged('gnb_degree_rows', gi.device):
        _lib.check(lib.gnb_degree_rows(gi.ref(), int(bool(swap)), x.data_ptr(), current_stream_ptr(gi.device)),
                   'gnb_degree_rows')
    return x


def zscore_cols(x, col_mask, want_stats=False):
    """In place ``x[:, c] = (x[:, c] - mean) / std`` (unbiased std) for the columns in ``col_mask`` (bit c)."""
    lib = _lib.load()
    if x.ndim != 2 or not 1 <= x.shape[1] <= 4:
        raise ValueError('x must be (rows, 1..4)')
    rows, cols = x.shape
    ws = torch.empty(lib.gnb_zscore_workspace() // 8, dtype=torch.float64, device=x.device)
    stats = torch.empty((cols, 2), dtype=torch.float64, device=x.device) if want_stats else None
    with _logged('gnb_zscore_cols', x.device, launches=3):
        _lib.check(lib.gnb_zscore_cols(_f32(x, 'x'), rows, cols, int(col_mask), _opt(stats), ws.data_ptr(),
                                       current_stream_ptr(x.device)), 'gnb_zscore_cols')
    return (x, stats) if want_stats else x


def node_subgraph(keep, src, dst, num_nodes):
    """``dgl.node_subgraph(g, keep, store_ids=True)`` on index tensors: ``keep`` (N,) bool / uint8, ``src`` / ``dst`` the
    parent's int32 edge list (CUDA).  Returns ``(node_id, edge_id, sub_src, sub_dst)``, all int32: the two '_ID' maps and
    the renumbered endpoints.  One host read of the two sizes sits between the two kernel passes."""
    import ctypes
    lib = _lib.load()
    device = src.device
    if not src.is_cuda or src.dtype != torch.int32 or dst.dtype != torch.int32 or not src.is_contiguous() or not dst.is_contiguous():
        raise ValueError('src / dst must be contiguous int32 CUDA tensors')
    keep = keep.to(device=device).ne(0).to(torch.uint8).contiguous()
    N, E = int(num_nodes), int(src.numel())
    if keep.numel() != N:
        raise ValueError(f'keep has {keep.numel()} entries for {N} nodes')
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.gnb_subgraph_workspace(N, E, ctypes.byref(nbytes)), 'gnb_subgraph_workspace')
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
    counts = torch.empty(2, dtype=torch.int64, device=device)
    with _logged('gnb_subgraph_count', device, launches=5):
        _lib.check(lib.gnb_subgraph_count(keep.data_ptr(), src.data_ptr(), dst.data_ptr(), N, E, ws.data_ptr(),
                                          nbytes.value, counts.data_ptr(), current_stream_ptr(device)), 'gnb_subgraph_count')
    n_sub, e_sub = (int(v) for v in counts.tolist())     # the one synchronisation: output sizes
    i32 = dict(dtype=torch.int32, device=device)
    node_id, edge_id = torch.empty(n_sub, **i32), torch.empty(e_sub, **i32)
    sub_src, sub_dst = torch.empty(e_sub, **i32), torch.empty(e_sub, **i32)
    with _logged('gnb_subgraph_fill', device, launches=2):
        _lib.check(lib.gnb_subgraph_fill(keep.data_ptr(), src.data_ptr(), dst.data_ptr(), N, E, ws.data_ptr(),
                                         node_id.data_ptr(), edge_id.data_ptr(), sub_src.data_ptr(), sub_dst.data_ptr(),
                                         current_stream_ptr(device)), 'gnb_subgraph_fill')
    return node_id, edge_id, sub_src, sub_dst
