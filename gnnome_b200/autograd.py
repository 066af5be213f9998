"""Training mode: the (Sym)GatedGCN layer, encoders and score predictor under torch autograd.

``train.py`` of the reference runs ``model(g, x, e)`` in ``model.train()`` mode, takes a BCE / symmetry loss of the
logits and calls ``loss.backward()`` (train.py:141-183, 329, 346).  The inference kernels are fused and keep no
activations, so training has its own path: the layer is written as layers/gated_gcn_full.py:82-142 writes it, on top
of a handful of graph primitives with hand-written CUDA forward and adjoint kernels (``csrc/gnb_train.cu``):

  ``GatherAdd3``   apply_edges(u_add_v) + B_3(e)            z_p = B1h[src_p] + B2h[dst_p] + B3e_p
  ``Agg``          update_all(u_mul_e, sum) / (copy_e, sum)  sum sigma*A[nbr] / (sum sigma + 1e-6), over in- or out-edges
  ``Gate``         relu, residual, sigmoid                    e' = relu(ehat) + e, sigma = sigmoid(e')
  ``BatchNormTrain``  BatchNorm1d with batch statistics over ALL E (or N) rows, running-stat update outside

The dense ``nn.Linear`` products are plain library GEMMs (``F.linear``), torch's autograd engine chains the pieces.
Edge rows are kept in dst-sorted position order; every sum runs over a CSR range in a fixed order (no atomics), so a
training step is bit-reproducible.  fp32 throughout.  ``normalization='layer'`` models (nn.LayerNorm, row kernels
of their own) run through this path in eval mode too: any hidden width that is a multiple of 4.
"""
import torch
import torch.nn.functional as F

from . import _lib
from .graph import GraphIndex, current_stream_ptr


def _c(t):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise ValueError(f'expected a float32 CUDA tensor, got {t.dtype} on {t.device}')
    return t if t.is_contiguous() else t.contiguous()


def _call(name, device, *args):
    from .ops import _logged          # launch accounting (bench.py), device guard
    lib = _lib.load()
    with _logged(name, device):
        _lib.check(getattr(lib, name)(*args, current_stream_ptr(device)), name)


_TC_WIDTHS = (64, 128, 256)      # contraction widths gnb_node_linear_tc2 is built for
_pack_cache = {}                 # id(weight) -> (version, packed W, packed W^T): re-packed when the parameter changes


def _packed(weight, transposed):
    from . import ops
    hit = _pack_cache.get(id(weight))
    if hit is None or hit[0] != weight._version or hit[3] is not weight:
        w = weight.detach().float().contiguous()
        hit = _pack_cache[id(weight)] = (weight._version, ops.pack_linear_tc(w), ops.pack_linear_tc(w.t().contiguous()), weight)
    return hit[2 if transposed else 1]


class TcLinear(torch.autograd.Function):
    """``x @ weight.T + bias`` for the training path on the tensor cores: forward and the input gradient
    ``gx = g @ weight`` run on ``gnb_node_linear_tc2`` (tcgen05, fp16 hi/lo split operands, fp32 accumulate: fp32-level
    accuracy) after one ``gnb_split_rows`` pass over the operand; the weight gradient
    ``gW = g.T @ x`` (a reduction over all rows into an out x in matrix) is a library GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from . import ops
        x = _c(x)
        M = weight.shape[0]
        b = bias.detach() if bias is not None else torch.zeros(M, dtype=torch.float32, device=x.device)
        out = ops.node_linear_tc2(ops.split_rows(x), _packed(weight, False), _c(b.float()), M)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        from . import ops
        x, weight = ctx.saved_tensors
        g = _c(g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # Gradients are ~1 / E small and the fp16 pair resolves |x| ~ 2^-4 .. 2^16: bring the tensor's largest
            # magnitude to 2^12 with a power of two (exact; a device scalar, no host round trip) and undo it in the
            # product's epilogue.
            K = weight.shape[1]
            amax = g.abs().amax().clamp_min(1e-30)
            s = torch.exp2(torch.floor(torch.log2(4096.0 / amax))).reshape(())
            gx = ops.node_linear_tc2(ops.split_rows(g, scale=s), _packed(weight, True),
                                     torch.zeros(K, dtype=torch.float32, device=g.device), K, out_scale=(1.0 / s))
        if ctx.needs_input_grad[1]:
            gw = g.t() @ x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.sum(0)
        return gx, gw, gb


def linear(lin_or_weight, x, bias=None):
    """``nn.Linear`` / ``F.linear`` of the training path: on the tensor cores where both contraction widths (in for the
    forward, out for the input gradient) are ones the kernel is built for and there are rows to process, else
    ``F.linear``."""
    if isinstance(lin_or_weight, torch.nn.Linear):
        weight, bias = lin_or_weight.weight, lin_or_weight.bias
    else:
        weight = lin_or_weight
    M, K = weight.shape
    if (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.shape[0] > 0
            and K in _TC_WIDTHS and M in _TC_WIDTHS):
        return TcLinear.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class GatherAdd3(torch.autograd.Function):
    """z[p] = A[src_p] + B[dst_p] + C[p]; A, B node tables [N][W], C edge rows [E][W] (position order)."""

    @staticmethod
    def forward(ctx, gi: GraphIndex, A, B, C):
        A, B, C = _c(A), _c(B), _c(C)
        W = A.shape[1]
        out = torch.empty((gi.E, W), dtype=torch.float32, device=A.device)
        _call('gnb_t_gather_add3', A.device, gi.ref(), W, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
              C.data_ptr(), out.data_ptr())
        ctx.gi, ctx.W = gi, W
        return out

    @staticmethod
    def backward(ctx, g):
        gi, W = ctx.gi, ctx.W
        g = _c(g)
        gA = torch.empty((gi.N, W), dtype=torch.float32, device=g.device)
        gB = torch.empty_like(gA)
        _call('gnb_t_seg_sum', g.device, gi.ref(), W, g.data_ptr(), 1, gA.data_ptr(), W)   # edges that leave the node
        _call('gnb_t_seg_sum', g.device, gi.ref(), W, g.data_ptr(), 0, gB.data_ptr(), W)   # edges that enter the node
        return None, gA, gB, g


class Agg(torch.autograd.Function):
    """out[i] = sum_p sigma[p] * A[nbr_p] / (sum_p sigma[p] + 1e-6).
    mode 0: over the in-edges of i, nbr = src (gated_gcn_full.py:112-114); mode 1: out-edges, nbr = dst (:125-127)."""

    @staticmethod
    def forward(ctx, gi: GraphIndex, A, sigma, mode):
        A, sigma = _c(A), _c(sigma)
        W = A.shape[1]
        den = torch.empty((gi.N, W), dtype=torch.float32, device=A.device)
        out = torch.empty_like(den)
        _call('gnb_t_agg_fwd', A.device, gi.ref(), W, A.data_ptr(), A.stride(0), sigma.data_ptr(), mode, den.data_ptr(),
              out.data_ptr())
        ctx.gi, ctx.W, ctx.mode = gi, W, mode
        ctx.save_for_backward(A, sigma, den, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        gi, W, mode = ctx.gi, ctx.W, ctx.mode
        A, sigma, den, out = ctx.saved_tensors
        gout = _c(gout)
        gsigma = torch.empty_like(sigma)
        _call('gnb_t_agg_bwd_edge', gout.device, gi.ref(), W, gout.data_ptr(), out.data_ptr(), den.data_ptr(), A.data_ptr(),
              A.stride(0), mode, gsigma.data_ptr(), 0)
        gA = torch.empty_like(A)
        _call('gnb_t_agg_bwd_node', gout.device, gi.ref(), W, gout.data_ptr(), den.data_ptr(), sigma.data_ptr(), mode,
              gA.data_ptr(), W)
        return None, gA, gsigma, None


class AggRaw(torch.autograd.Function):
    """(num, den)[i] = (sum_p sigma[p] * A[nbr_p], sum_p sigma[p]) -- ``Agg`` without the division, both outputs
    differentiable: the sharded training step adds the partial sums of several ranks before dividing."""

    @staticmethod
    def forward(ctx, gi: GraphIndex, A, sigma, mode):
        A, sigma = _c(A), _c(sigma)
        W = A.shape[1]
        den = torch.empty((gi.N, W), dtype=torch.float32, device=A.device)
        num = torch.empty_like(den)
        _call('gnb_t_agg_fwd', A.device, gi.ref(), W, A.data_ptr(), A.stride(0), sigma.data_ptr(), mode | 2, den.data_ptr(),
              num.data_ptr())
        ctx.gi, ctx.W, ctx.mode = gi, W, mode
        ctx.save_for_backward(A, sigma)
        return num, den

    @staticmethod
    def backward(ctx, gnum, gden):
        gi, W, mode = ctx.gi, ctx.W, ctx.mode
        A, sigma = ctx.saved_tensors
        gnum = torch.zeros((gi.N, W), dtype=torch.float32, device=A.device) if gnum is None else _c(gnum)
        gden = torch.zeros((gi.N, W), dtype=torch.float32, device=A.device) if gden is None else _c(gden)
        gsigma = torch.empty_like(sigma)
        _call('gnb_t_agg_bwd_edge', A.device, gi.ref(), W, gnum.data_ptr(), gden.data_ptr(), None, A.data_ptr(),
              A.stride(0), mode | 2, gsigma.data_ptr(), 0)
        gA = torch.empty_like(A)
        _call('gnb_t_agg_bwd_node', A.device, gi.ref(), W, gnum.data_ptr(), None, sigma.data_ptr(), mode | 2,
              gA.data_ptr(), W)
        return None, gA, gsigma, None


class Gate(torch.autograd.Function):
    """e' = relu(ehat) (+ e_in), sigma = sigmoid(e')  (gated_gcn_full.py:107-111)."""

    @staticmethod
    def forward(ctx, ehat, e_in):
        ehat = _c(ehat)
        e_in = None if e_in is None else _c(e_in)
        rows, W = ehat.shape
        e_out, sigma = torch.empty_like(ehat), torch.empty_like(ehat)
        _call('gnb_t_gate_fwd', ehat.device, ehat.data_ptr(), None if e_in is None else e_in.data_ptr(), rows, W,
              e_out.data_ptr(), sigma.data_ptr())
        ctx.has_res = e_in is not None
        ctx.save_for_backward(ehat, sigma)
        return e_out, sigma

    @staticmethod
    def backward(ctx, g_e, g_sigma):
        ehat, sigma = ctx.saved_tensors
        rows, W = ehat.shape
        g_e = torch.zeros_like(ehat) if g_e is None else _c(g_e)
        g_sigma = torch.zeros_like(ehat) if g_sigma is None else _c(g_sigma)
        g_ehat = torch.empty_like(ehat)
        g_ein = torch.empty_like(ehat) if ctx.has_res else None
        _call('gnb_t_gate_bwd', ehat.device, g_e.data_ptr(), g_sigma.data_ptr(), ehat.data_ptr(), sigma.data_ptr(), rows, W,
              g_ehat.data_ptr(), None if g_ein is None else g_ein.data_ptr())
        return g_ehat, g_ein


def _col_stats(a, b=None, shift_a=None, shift_b=None):
    """fp64 [2][W]: column sums of a' = a - shift_a and of a' * b' (b None: a' * a'); deterministic (per-block fp32
    partial sums, fp64 combine in block order)."""
    lib = _lib.load()
    rows, W = a.shape
    out = torch.empty((2, W), dtype=torch.float64, device=a.device)
    ws = torch.empty(max(lib.gnb_t_col_stats_workspace(rows, W), 8), dtype=torch.uint8, device=a.device)
    ptr = lambda t: None if t is None else t.data_ptr()
    _call('gnb_t_col_stats', a.device, a.data_ptr(), ptr(b), ptr(shift_a), ptr(shift_b), rows, W, out.data_ptr(),
          ws.data_ptr())
    return out


def _affine2(x, y, a, b, c, shift_x=None, shift_y=None):
    """a * (x - shift_x) (+ b * (y - shift_y)) + c per channel; the fp64 coefficient vectors are rounded to fp32 here."""
    rows, W = x.shape
    out = torch.empty_like(x)
    f32 = lambda t: None if t is None else t.float().contiguous()  # noqa: E731
    a, b, c, shift_x, shift_y = (f32(t) for t in (a, b, c, shift_x, shift_y))
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    _call('gnb_t_affine2', x.device, x.data_ptr(), ptr(y), a.data_ptr(), ptr(b), c.data_ptr(), ptr(shift_x), ptr(shift_y),
          rows, W, out.data_ptr())
    return out


class BatchNormTrain(torch.autograd.Function):
    """nn.BatchNorm1d in training mode: statistics over all rows (biased variance for normalising).
    Returns (y, batch_mean, batch_var_biased); the running-statistics update is the caller's (it is applied twice
    per layer for ``bn_e``, gated_gcn_full.py:106,119)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = _c(x)
        n = x.shape[0]
        mean32 = (_col_stats(x)[0] / n).float().contiguous()          # pass 1: mean
        st = _col_stats(x, shift_a=mean32)                             # pass 2: moments about it (no cancellation)
        mean = mean32.double() + st[0] / n
        var = (st[1] / n - (st[0] / n) ** 2).clamp_min_(0.0)
        rstd = torch.rsqrt(var + eps)
        # y = a * (x - mean) + bias, centred on the fp32 first-pass mean (its fp64 correction goes into the constant):
        # a * x + (bias - a * mean) loses the digits of a channel whose mean is large against its spread
        a = weight.detach().double() * rstd
        c = bias.detach().double() - (mean - mean32.double()) * a
        y = _affine2(x, None, a, None, c, shift_x=mean32)
        ctx.save_for_backward(x, weight, mean, rstd, mean32)
        ctx.n = n
        mean_f, var_f = mean.float(), var.float()
        ctx.mark_non_differentiable(mean_f, var_f)
        return y, mean_f, var_f

    @staticmethod
    def backward(ctx, g, _gm, _gv):
        x, weight, mean, rstd, mean32 = ctx.saved_tensors
        n = ctx.n
        g = _c(g)
        st = _col_stats(g, x, shift_b=mean32)                           # sum g, sum g * (x - mean32)
        c1 = st[0]
        c2 = rstd * (st[1] - (mean - mean32.double()) * c1)             # sum g * xhat
        w = weight.detach().double()
        a = w * rstd
        b = -w * rstd * rstd * c2 / n
        # gx = a * (g - mean(g)) + b * (x - mean): centred like the forward (torch's backward centres x too)
        c = -a * c1 / n + b * (mean32.double() - mean)
        gx = _affine2(g, x, a, b, c, shift_y=mean32)
        return gx, c2.to(weight.dtype), c1.to(weight.dtype), None


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the channels of each row (normalization='layer', gated_gcn_full.py:39-42)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = _c(x)
        rows, W = x.shape
        y, xhat = torch.empty_like(x), torch.empty_like(x)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w, b = _c(weight.detach()), _c(bias.detach())
        _call('gnb_t_layer_norm_fwd', x.device, x.data_ptr(), w.data_ptr(), b.data_ptr(), rows, W, float(eps), y.data_ptr(),
              xhat.data_ptr(), rstd.data_ptr())
        ctx.save_for_backward(xhat, rstd, w)
        return y

    @staticmethod
    def backward(ctx, g):
        xhat, rstd, w = ctx.saved_tensors
        g = _c(g)
        rows, W = g.shape
        gx = torch.empty_like(g)
        _call('gnb_t_layer_norm_bwd', g.device, g.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), w.data_ptr(), rows, W,
              gx.data_ptr())
        st = _col_stats(g, xhat)                       # sum g, sum g * xhat
        return gx, st[1].to(w.dtype), st[0].to(w.dtype), None


def normalize(norm, x, training, updates=1):
    """``norm(x)`` for the layer's ``bn_h`` / ``bn_e`` sub-module: BatchNorm1d or LayerNorm."""
    if isinstance(norm, torch.nn.LayerNorm):
        return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)
    return batch_norm(norm, x, training, updates)


def batch_norm(bn: torch.nn.BatchNorm1d, x, training, updates=1):
    """``bn(x)`` through the CUDA primitives; in training mode the running statistics are updated ``updates`` times
    with the same batch statistics (unbiased variance, momentum), like calling the module ``updates`` times."""
    if not training:
        var = bn.running_var.detach().double()
        a = bn.weight.double() / torch.sqrt(var + bn.eps)
        c = bn.bias.double() - bn.running_mean.detach().double() * a
        return x * a.float() + c.float()
    y, mean, var = BatchNormTrain.apply(x, bn.weight, bn.bias, bn.eps)
    if bn.track_running_stats:
        n = x.shape[0]
        with torch.no_grad():
            unbiased = var * (n / max(n - 1, 1))
            for _ in range(updates):
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
                bn.running_var.mul_(1 - m).add_(unbiased.to(bn.running_var.dtype), alpha=m)
    return y


# ------------------------------------------------------------------------------------------------
# layer / model forward in training mode
# ------------------------------------------------------------------------------------------------
def layer_forward(conv, gi: GraphIndex, h, e_pos):
    """One (Sym)GatedGCN layer, edge rows in position order (gated_gcn_full.py:82-142 / :182-230)."""
    if conv.normalization not in ('batch', 'layer'):
        raise NotImplementedError(f"normalization={conv.normalization!r} (the reference itself fails on 'none': bn_e is "
                                  f"used unconditionally, gated_gcn_full.py:106)")
    sym = conv._symmetric
    # library GEMMs here: this single-GPU path is pinned to the reference's training fixture, whose gradients sit on
    # ReLU kinks; the sharded trainer (train_dist.py) runs its Linears on the tensor cores
    A1h, A2h = conv.A_1(h), conv.A_2(h)                                  # :91-92
    B1h, B2h, B3e = conv.B_1(h), conv.B_2(h), conv.B_3(e_pos)            # :95-97
    z = GatherAdd3.apply(gi, B1h, B2h, B3e)                              # :104-105
    ehat = normalize(conv.bn_e, z, conv.training, updates=2 if sym else 1)    # :106 (+ :119 on the reversed graph)
    e_new, sigma = Gate.apply(ehat, e_pos if conv.residual else None)    # :107-111
    u = A1h + Agg.apply(gi, A2h, sigma, 0)                               # :112-114
    if sym:
        u = u + Agg.apply(gi, conv.A_3(h), sigma, 1)                     # :93, :125-127 (same sigma)
    u = normalize(conv.bn_h, u, conv.training)                           # :131-132
    h_new = torch.relu(u)                                                # :134
    if conv.residual:
        h_new = h_new + h                                                # :136-137
    h_new = F.dropout(h_new, conv.dropout, training=conv.training)       # :139
    return h_new, e_new


def predictor_forward(pred, gi: GraphIndex, x, e_pos):
    """ScorePredictor (score_predictor.py:12-24) with W1 split into its src / dst / edge column blocks."""
    H = pred.in_features
    W1, b1 = pred.W1.weight, pred.W1.bias
    S1 = F.linear(x, W1[:, :H])
    S2 = F.linear(x, W1[:, H:2 * H], b1)
    E1 = F.linear(e_pos, W1[:, 2 * H:])      # W1 column blocks are views (a packed copy per step would need its own cache)
    hid = torch.relu(GatherAdd3.apply(gi, S1, S2, E1))
    return pred.W3(torch.relu(pred.W2(hid)))                             # [E][1], position order


def _device_replica(model, dev):
    """A copy of an eval-mode model on ``dev``, rebuilt when a parameter or buffer changes."""
    key = tuple((id(t), t._version) for t in list(model.parameters()) + list(model.buffers())) + (str(dev),)
    cache = model.__dict__.get('_gnb_replica')
    if cache is None or cache[0] != key:
        import copy
        model.__dict__.pop('_gnb_replica', None)
        cache = model.__dict__['_gnb_replica'] = (key, copy.deepcopy(model).to(dev).eval())
    return cache[1]


def model_forward(model, graph, x, e):
    """``SymGatedGCNModel`` / ``GatedGCNModel(directed=True)`` forward under autograd (models/full_graph.py:22-30, 42-53)."""
    gi = GraphIndex.from_graph(graph)
    dev = gi.device
    sym_model = hasattr(model, 'linear1_node')
    if not sym_model and not getattr(model, 'directed', True):
        raise NotImplementedError('training GatedGCNModel(directed=False) is not built')
    if any(p.device != dev for p in model.parameters()):
        if model.training:
            raise RuntimeError(f'training needs the model on {dev}: call model.to(device) first')
        model = _device_replica(model, dev)      # eval with CPU-resident parameters (inference.py:388)
    out_dev = x.device
    x_d = x.to(device=dev, dtype=torch.float32)
    order = gi.in_eid[:gi.E].long()
    e_d = e.to(device=dev, dtype=torch.float32)[order]                   # edge rows in position order from here on
    if sym_model:
        h = model.linear2_node(torch.relu(model.linear1_node(x_d)))      # :26
        ee = model.linear2_edge(torch.relu(model.linear1_edge(e_d)))     # :27
    else:
        ne, en = model.node_encoder, model.edge_encoder              # node_encoder.py:28-33, edge_encoder.py:27-32
        h = ne.linear2(torch.relu(ne.linear1(x_d)))
        ee = en.linear2(torch.relu(en.linear1(e_d)))
    for conv in model.gnn.convs:                                         # processor.py:16-19
        h, ee = layer_forward(conv, gi, h, ee)
    s_pos = predictor_forward(model.predictor, gi, h, ee)
    scores = torch.empty_like(s_pos).index_copy(0, order, s_pos)         # back to edge-id order
    return scores.to(out_dev)
