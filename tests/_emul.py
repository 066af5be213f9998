"""TEST INFRASTRUCTURE: a torch (CPU) emulation of the per-op contracts of ``gnnome_b200.partition``'s
kernel object (``CudaKernels``), i.e. of the C-ABI operators as ``include/gnnome_b200.h`` documents
them.  It lets the CPU ``gloo`` tests exercise the partition / halo-exchange logic without a GPU; it is
never imported by the product package."""
import torch
import torch.nn.functional as F

EPS = 1e-6


class _Graph:
    def __init__(self, src, dst, n):
        self.N, self.E = int(n), int(src.numel())
        dst, src = dst.long(), src.long()
        self.in_eid = torch.argsort(dst, stable=True)
        self.src, self.dst = src[self.in_eid], dst[self.in_eid]       # position order


class EmulKernels:
    def __init__(self, dtype=torch.float64):
        self.dtype = dtype

    def stage(self, src_local, dst_local, n_local):
        return _Graph(src_local.cpu(), dst_local.cpu(), n_local)

    def position_eids(self, gi):
        return gi.in_eid

    def encode_nodes(self, x, lin1, lin2, rows):
        return self.encode(x, None, lin1, lin2, rows)

    def encode_edges(self, e, idx, lin1, lin2, rows):
        return self.encode(e, idx, lin1, lin2, rows)

    def width(self, h):
        return h.shape[1]

    def encode(self, x, idx, lin1, lin2, rows):
        x = x.to(self.dtype)
        if idx is not None:
            x = x[idx.long()]
        w = lambda t: t.detach().to(self.dtype)
        return F.linear(torch.relu(F.linear(x[:rows], w(lin1.weight), w(lin1.bias))), w(lin2.weight), w(lin2.bias))

    def layer_pack(self, conv):
        return conv

    def node_linear_layer(self, conv, h, M, out):
        H = h.shape[1]
        lin = lambda m: F.linear(h, m.weight.detach().to(self.dtype), m.bias.detach().to(self.dtype))
        out[:, 0:2 * H:2] = lin(conv.B_1)
        out[:, 1:2 * H:2] = lin(conv.A_2)
        out[:, 2 * H:3 * H] = lin(conv.B_2)
        if conv._symmetric:
            out[:, 3 * H:4 * H] = lin(conv.A_3)
            out[:, 4 * H:5 * H] = lin(conv.A_1)
        else:
            out[:, 3 * H:4 * H] = lin(conv.A_1)
        return out

    def project_rows(self, conv, h, idx, M, out):
        H = h.shape[1]
        assert M == 2 * H
        hs = h[idx.long()]
        lin = lambda m: F.linear(hs, m.weight.detach().to(self.dtype), m.bias.detach().to(self.dtype))
        out[:, 0:2 * H:2] = lin(conv.B_1)
        out[:, 1:2 * H:2] = lin(conv.A_2)
        return out

    def gather_rows(self, table, idx, out=None):
        res = table[idx.long()]
        if out is not None:
            out.copy_(res)
            return out
        return res

    def _affine(self, bn):
        d = lambda t: t.detach().to(self.dtype)
        scale = d(bn.weight) / torch.sqrt(d(bn.running_var) + bn.eps)
        return scale, d(bn.bias) - d(bn.running_mean) * scale

    def edge_forward(self, gi, H, P, conv, e_pos, Fb, carry, flags):
        sc, sh = self._affine(conv.bn_e)
        z = P[gi.src, 0:2 * H:2] + P[gi.dst, 2 * H:3 * H] + F.linear(
            e_pos, conv.B_3.weight.detach().to(self.dtype), conv.B_3.bias.detach().to(self.dtype))
        v = torch.relu(z * sc + sh)
        if flags & 2:
            v = v + e_pos
        e_pos.copy_(v)
        sg = torch.sigmoid(v)
        num = torch.zeros((gi.N, H), dtype=self.dtype).index_add_(0, gi.dst, sg * P[gi.src, 1:2 * H:2])
        den = torch.zeros((gi.N, H), dtype=self.dtype).index_add_(0, gi.dst, sg)
        Fb.copy_(num / (den + EPS))

    def carry_shape(self, gi, H):
        return (1, 4, H)

    def _reverse_sums(self, gi, H, P, e_pos):
        sg = torch.sigmoid(e_pos)
        num = torch.zeros((gi.N, H), dtype=self.dtype).index_add_(0, gi.src, sg * P[gi.dst, 3 * H:4 * H])
        den = torch.zeros((gi.N, H), dtype=self.dtype).index_add_(0, gi.src, sg)
        return num, den

    def reverse_partial(self, gi, H, P, e_pos, node_begin, node_end, out):
        num, den = self._reverse_sums(gi, H, P, e_pos)
        out[:, :H] = num[node_begin:node_end]
        out[:, H:] = den[node_begin:node_end]

    def node_update(self, gi, H, P, conv, e_pos, Fb, carry, h_in, flags, n_own, xp_ptr, xp_row, xp_buf):
        sym = bool(flags & 1)
        u = P[:n_own, (4 if sym else 3) * H:(5 if sym else 4) * H] + Fb[:n_own]
        if sym:
            num, den = self._reverse_sums(gi, H, P, e_pos)
            num, den = num[:n_own].clone(), den[:n_own].clone()
            if xp_buf is not None:
                node_of = torch.repeat_interleave(torch.arange(n_own), (xp_ptr[1:] - xp_ptr[:-1]).long())
                rows = xp_buf[xp_row.long()]
                num.index_add_(0, node_of, rows[:, :H])
                den.index_add_(0, node_of, rows[:, H:])
            u = u + num / (den + EPS)
        sc, sh = self._affine(conv.bn_h)
        v = torch.relu(u * sc + sh)
        if flags & 2:
            v = v + h_in[:n_own]
        return v

    def score_node_rows(self, pred, x):
        H, hs = pred.in_features, pred.hidden_edge_scores
        W1, b1 = pred.W1.weight.detach().to(self.dtype), pred.W1.bias.detach().to(self.dtype)
        return torch.cat((x @ W1[:, :H].t(), x @ W1[:, H:2 * H].t() + b1), dim=1)

    def score_forward(self, pred, gi, S, e_pos, scores):
        H, hs = pred.in_features, pred.hidden_edge_scores
        d = lambda t: t.detach().to(self.dtype)
        hid = torch.relu(S[gi.src, :hs] + S[gi.dst, hs:] + e_pos @ d(pred.W1.weight)[:, 2 * H:].t())
        out = F.linear(torch.relu(F.linear(hid, d(pred.W2.weight), d(pred.W2.bias))), d(pred.W3.weight), d(pred.W3.bias))
        scores[gi.in_eid] = out.to(scores.dtype)
        return scores


class TorchPrims:
    """TEST INFRASTRUCTURE: plain-torch (autograd-native) implementation of the contracts
    ``gnnome_b200.train_dist.CudaPrims`` offers, so that the CPU ``gloo`` tests can check the distributed logic of the
    sharded training step (statistics, exchanges and their adjoints, loss scaling, gradient reduction)."""

    def stage(self, src_local, dst_local, n_local):
        return _Graph(src_local.cpu(), dst_local.cpu(), n_local)

    def position_eids(self, gi):
        return gi.in_eid

    def gather_add3(self, gi, A_, B_, C_):
        return A_.index_select(0, gi.src) + B_.index_select(0, gi.dst) + C_

    def agg_in(self, gi, A_, sigma):
        z = lambda: torch.zeros((gi.N, sigma.shape[1]), dtype=sigma.dtype)  # noqa: E731
        num = z().index_add(0, gi.dst, sigma * A_.index_select(0, gi.src))
        den = z().index_add(0, gi.dst, sigma)
        return num / (den + EPS)

    def agg_out_raw(self, gi, A_, sigma):
        z = lambda: torch.zeros((gi.N, sigma.shape[1]), dtype=sigma.dtype)  # noqa: E731
        return z().index_add(0, gi.src, sigma * A_.index_select(0, gi.dst)), z().index_add(0, gi.src, sigma)

    def gate(self, ehat, e_in):
        e_new = torch.relu(ehat) if e_in is None else torch.relu(ehat) + e_in
        return e_new, torch.sigmoid(e_new)

    def col_stats(self, a, b=None, shift_a=None, shift_b=None):
        a = a.double() - (0 if shift_a is None else shift_a.double())
        b = a if b is None else b.double() - (0 if shift_b is None else shift_b.double())
        return torch.stack((a.sum(0), (a * b).sum(0)))

    def affine2(self, x, y, a, b, c, shift_x=None, shift_y=None):
        out = a * (x.double() - (0 if shift_x is None else shift_x.double())) + c
        if y is not None:
            out = out + b * (y.double() - (0 if shift_y is None else shift_y.double()))
        return out.to(x.dtype)

    def layer_norm(self, norm, x):
        return F.layer_norm(x, (x.shape[1],), norm.weight, norm.bias, norm.eps)

    def linear(self, lin, x):
        return lin(x)
