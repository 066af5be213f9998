"""(Sym)GatedGCN layers with the reference's constructor / parameter names, running on the
sm_100a kernels.  Mirrors ``layers/gated_gcn_full.py`` of the reference (SymGatedGCN :8-142,
GatedGCN :145-230): same ``nn.Linear`` / norm sub-modules (so ``state_dict`` keys and shapes are
identical and ``weights/weights.pt`` loads strictly), same ``forward(g, h, e) -> (h, e)``.

The arithmetic is NOT torch: one concatenated node projection, one fused edge pass over the
dst-CSR, one reverse aggregation + node update over the src-CSR (see ``csrc/gnb_layers.cu``)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, ops
from ..graph import GraphIndex


_BACKEND = 'tc2'
_SPLIT16_WIDTHS = (64, 128, 256)


def set_backend(name):
    """'tc2' (default): TMA-fed tcgen05 kernels with the state held as split fp16 (hi, lo) images
    (hidden_features 64 / 128 / 256; hidden_features = 32 is too narrow for the 128-byte swizzle and runs 'ffma');
    'ffma': the CUDA-core fp32 kernels -- an independent code path kept as the on-device cross-check."""
    global _BACKEND
    if name not in ('tc2', 'ffma'):
        raise ValueError(name)
    _BACKEND = name


def get_backend():
    return _BACKEND


def effective_backend(H):
    """The kernel family that actually runs a layer of width H under the current backend."""
    if _BACKEND == 'tc2':
        return 'tc2' if H in _SPLIT16_WIDTHS else 'ffma'
    return _BACKEND


def state_format(H):
    """'split16' (fp16 hi/lo images, gnb_tma.cuh) or 'fp32' rows: how h / e are held between kernels."""
    return 'split16' if effective_backend(H) == 'tc2' else 'fp32'


def _bn_affine(norm, device):
    """Eval-mode BatchNorm1d as a per-channel affine, computed in fp64 (gated_gcn_full.py:37-38)."""
    var = norm.running_var.detach().double()
    scale = norm.weight.detach().double() / torch.sqrt(var + norm.eps)
    shift = norm.bias.detach().double() - norm.running_mean.detach().double() * scale
    return scale.to(device), shift.to(device)


class _GatedGCNBase(nn.Module):
    _symmetric = True

    def __init__(self, in_channels, out_channels, normalization, dropout=None, residual=True):
        super().__init__()
        self.dropout = dropout if dropout else 0.0                 # gated_gcn_full.py:16-19
        self.normalization = normalization
        self.residual = residual
        if in_channels != out_channels:                            # :24-25
            self.residual = False
        self.in_channels, self.out_channels = in_channels, out_channels
        dtype = torch.float32
        self.A_1 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.A_2 = nn.Linear(in_channels, out_channels, dtype=dtype)
        if self._symmetric:
            self.A_3 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.B_1 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.B_2 = nn.Linear(in_channels, out_channels, dtype=dtype)
        self.B_3 = nn.Linear(in_channels, out_channels, dtype=dtype)
        if normalization == 'batch':
            self.bn_h = nn.BatchNorm1d(out_channels, track_running_stats=True)
            self.bn_e = nn.BatchNorm1d(out_channels, track_running_stats=True)
        elif normalization == 'layer':
            self.bn_h = nn.LayerNorm(out_channels)
            self.bn_e = nn.LayerNorm(out_channels)

    # -- parameter packing ---------------------------------------------------------------------
    def _flags(self):
        return (_lib.GNB_F_SYMMETRIC if self._symmetric else 0) | (_lib.GNB_F_RESIDUAL if self.residual else 0)

    def _pack(self, device, family=None):
        """Kernel-side views of the parameters for one kernel family ('tc2' / 'ffma', see _pack_now), cached until a
        parameter / buffer changes (torch bumps ``_version`` on every in-place update) or moves."""
        family = family or effective_backend(self.out_channels)
        key = tuple((id(t), t._version, t.device) for t in list(self.parameters()) + list(self.buffers()))
        key += (str(device),)
        cache = self.__dict__.setdefault('_pack_cache', {})
        hit = cache.get(family)
        if hit is None or hit[0] != key:
            hit = cache[family] = (key, self._pack_now(device, family))
        return hit[1]

    def _pack_now(self, device, family):
        """The concatenated node projection with the (B1, A2) rows interleaved per channel, the
        edge projection, and the eval-mode norm affines (b_B3 folded into the edge shift)."""
        H = self.out_channels
        dev = dict(device=device, dtype=torch.float32)
        w = lambda lin: lin.weight.detach().to(**dev)
        b = lambda lin: lin.bias.detach().to(**dev)
        blocks_w = [torch.stack((w(self.B_1), w(self.A_2)), dim=1).reshape(2 * H, -1), w(self.B_2)]
        blocks_b = [torch.stack((b(self.B_1), b(self.A_2)), dim=1).reshape(2 * H), b(self.B_2)]
        if self._symmetric:
            blocks_w.append(w(self.A_3))
            blocks_b.append(b(self.A_3))
        blocks_w.append(w(self.A_1))
        blocks_b.append(b(self.A_1))
        Wn = torch.cat(blocks_w, dim=0).contiguous()                 # [5H or 4H][H_in] (nn.Linear layout)
        bn = torch.cat(blocks_b, dim=0).contiguous()
        if self.normalization == 'batch':
            se, te = _bn_affine(self.bn_e, device)
            sh, th = _bn_affine(self.bn_h, device)
        else:
            raise NotImplementedError(f"normalization={self.normalization!r}: only eval-mode 'batch' runs on the "
                                      f"CUDA path so far")
        te = te + se * self.B_3.bias.detach().double().to(device)
        f32 = lambda t: t.to(torch.float32).contiguous()
        if family == 'tc2':
            # gnb_edge_forward_tc2 takes the edge norm folded into its operands (fp64 here, rounded once):
            #   bn_e(B1h[s] + B2h[d] + B_3 e) = (se * B1h)[s] + (se * B2h + te)[d] + (diag(se) B_3) e
            Wn64, bn64, W364 = Wn.double(), bn.double(), w(self.B_3).double()
            Wn64[0:2 * H:2] *= se[:, None]                           # B_1 rows of the interleaved (B1, A2) block
            bn64[0:2 * H:2] *= se
            Wn64[2 * H:3 * H] *= se[:, None]                         # B_2 block
            bn64[2 * H:3 * H] = bn64[2 * H:3 * H] * se + te
            Wn_t, We_t = ops.pack_linear_tc(f32(Wn64)), ops.pack_linear_tc(f32(W364 * se[:, None]))
            bn = f32(bn64)
        else:
            Wn_t = Wn.t().contiguous()                               # k-major [H_in][5H or 4H]
            We_t = w(self.B_3).t().contiguous()                      # k-major [H_in][H]
        return dict(Wn_t=Wn_t, bn=bn, We_t=We_t, scale_e=f32(se), shift_e=f32(te), scale_h=f32(sh), shift_h=f32(th))

    # -- position-order fast path (what the processor / model use) -----------------------------
    def forward_positions(self, gi: GraphIndex, h, e_pos, ws=None):
        """One layer with edge rows already in dst-sorted position order.  ``e_pos`` is updated IN
        PLACE (legal: e'_p depends only on e_p, h[src_p], h[dst_p]) and returned."""
        if self.training:
            raise NotImplementedError('training mode runs through gnnome_b200.autograd: call forward()')
        if self.in_channels != self.out_channels:
            raise NotImplementedError('in_channels != out_channels is not supported by the CUDA path')
        H = self.out_channels
        dev = h.device
        pk = self._pack(dev, 'ffma')
        ws = ws if ws is not None else {}
        n_blocks = 5 if self._symmetric else 4
        P = ws.get('P')
        if P is None or P.shape != (gi.N, n_blocks * H):
            P = ws['P'] = torch.empty((gi.N, n_blocks * H), dtype=torch.float32, device=dev)
        Fb = ws.get('F')
        if Fb is None or Fb.shape != (gi.N, H):
            Fb = ws['F'] = torch.empty((gi.N, H), dtype=torch.float32, device=dev)
        carry = ws.get('carry')
        n_chunks = gi.num_chunks(H, 'ffma')
        if carry is None or carry.shape != (n_chunks, 4, H):
            carry = ws['carry'] = torch.empty((n_chunks, 4, H), dtype=torch.float32, device=dev)
        h_out = ws.pop('h_spare', None)
        if h_out is None or h_out.shape != h.shape or h_out.data_ptr() == h.data_ptr():
            h_out = torch.empty_like(h)
        flags = self._flags()
        ops.node_linear(h, pk['Wn_t'], pk['bn'], out=P)
        ops.edge_forward(gi, H, P, pk['We_t'], pk['scale_e'], pk['shift_e'], e_pos, Fb, carry, flags)
        ops.node_update(gi, H, P, e_pos, Fb, carry, h, pk['scale_h'], pk['shift_h'], h_out, flags,
                        gi.chunk(H, 'ffma'))
        ws['h_spare'] = h  # ping-pong: the caller no longer needs the input h
        return h_out, e_pos

    def forward_positions16(self, gi: GraphIndex, h32, h16, e16, ws=None):
        """One layer on the split16 path: ``h32`` fp32 rows (residual), ``h16`` / ``e16`` split fp16 images
        (the tensor-core operands).  ``e16`` is updated IN PLACE; returns ``(h32', h16', e16)``."""
        if self.training:
            raise NotImplementedError('training mode runs through gnnome_b200.autograd: call forward()')
        if self.in_channels != self.out_channels:
            raise NotImplementedError('in_channels != out_channels is not supported by the CUDA path')
        H = self.out_channels
        dev = h32.device
        pk = self._pack(dev, 'tc2')
        ws = ws if ws is not None else {}
        n_blocks = 5 if self._symmetric else 4

        def buf(name, shape, dtype=torch.float32):
            t = ws.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = ws[name] = torch.empty(shape, dtype=dtype, device=dev)
            return t

        P = buf('P', (gi.N, n_blocks * H))
        Fb = buf('F', (gi.N, H))
        carry = buf('carry', (gi.num_chunks(H, 'tc2'), 4, H))
        h_out = ws.pop('h_spare', None)
        if h_out is None or h_out.shape != h32.shape or h_out.data_ptr() == h32.data_ptr():
            h_out = torch.empty_like(h32)
        h16_out = ws.pop('h16_spare', None)
        if h16_out is None or h16_out.shape != h16.shape or h16_out.data_ptr() == h16.data_ptr():
            h16_out = torch.empty_like(h16)
        flags = self._flags()
        ops.node_linear_tc2(h16, pk['Wn_t'], pk['bn'], n_blocks * H, out=P)
        ops.edge_forward_tc2(gi, H, P, pk['We_t'], e16, Fb, carry, flags)
        ops.node_update2(gi, H, P, e16, Fb, carry, h32, pk['scale_h'], pk['shift_h'], h_out, h16_out, flags,
                         gi.chunk(H, 'tc2'))
        ws['h_spare'], ws['h16_spare'] = h32, h16   # ping-pong: the caller no longer needs the inputs
        return h_out, h16_out, e16

    # -- reference-compatible layer API ------------------------------------------------------------
    def forward(self, g, h, e):
        """``h, e = conv(g, h, e)`` with ``e`` in the graph's edge-id order (gated_gcn_full.py:82)."""
        gi = GraphIndex.from_graph(g)
        out_dev = h.device
        if self.training or self.normalization == 'layer' or self.in_channels != self.out_channels:
            # gnnome_b200.autograd primitives (any width that is a multiple of 4), position order inside
            from ..autograd import layer_forward
            order = gi.in_eid[:gi.E].long()
            conv = self
            if any(p.device != gi.device for p in self.parameters()):
                if self.training:
                    raise RuntimeError(f'training needs the layer on {gi.device}: call .to(device) first')
                from ..autograd import _device_replica
                conv = _device_replica(self, gi.device)
            h_new, e_pos = layer_forward(conv, gi, h.to(device=gi.device, dtype=torch.float32),
                                         e.to(device=gi.device, dtype=torch.float32)[order])
            return h_new.to(out_dev), torch.empty_like(e_pos).index_copy(0, order, e_pos).to(out_dev)
        h_d = h.detach().to(device=gi.device, dtype=torch.float32).contiguous()
        e_d = e.detach().to(device=gi.device, dtype=torch.float32).contiguous()
        if state_format(self.out_channels) == 'split16' and gi.E > 0:
            e16 = ops.split_rows(e_d, gi.in_eid[:gi.E])
            h_new, _, e16 = self.forward_positions16(gi, h_d.clone(), ops.split_rows(h_d), e16)
            e_new = ops.merge_rows(e16, gi.in_eid[:gi.E])
        else:
            e_pos = ops.gather_rows(e_d, gi.in_eid[:gi.E])
            h_new, e_pos = self.forward_positions(gi, h_d.clone(), e_pos)
            e_new = ops.scatter_rows(e_pos, gi.in_eid[:gi.E])
        h_new = F.dropout(h_new, self.dropout, training=self.training)  # :139
        return h_new.to(out_dev), e_new.to(out_dev)


class SymGatedGCN(_GatedGCNBase):
    """reference layers/gated_gcn_full.py:8-142"""
    _symmetric = True


class GatedGCN(_GatedGCNBase):
    """reference layers/gated_gcn_full.py:145-230 (no A_3, no reverse aggregation)"""
    _symmetric = False
