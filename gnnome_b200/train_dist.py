"""Sharded training step -- BASELINE.json config 5 (``train.py`` full forward + backward, BCE on the edge scores,
synthetic 4M nodes / 24M edges, hidden 256, L = 8, up to 8 B200s of one box).

The reference trains on one device (``train.py:141-145, 329, 346``; SURVEY.md section 2.1: no distributed code).  Here
the graph is partitioned exactly like the inference path (``gnnome_b200.partition``: rank r owns a destination-node
range, every edge that enters it and the state of both), one process per GPU, and one optimisation step is

  forward   per layer: halo exchange of the ``h`` rows of remote SOURCE nodes (owner -> consumer), the layer on the
            local graph (``gnnome_b200.autograd`` primitives: hand-written CUDA forward / adjoint kernels), train-mode
            BatchNorm with the batch statistics ALL-REDUCED over the ranks (``bn_e`` over all E rows, ``bn_h`` over all
            N rows -- ``layers/gated_gcn_full.py:37-38,106,119,132``; two updates of ``bn_e``'s running statistics per
            Sym layer, as the reference's two calls make), the reverse aggregation's partial sums
            ``(sum sigma * A3h[dst], sum sigma)`` of remote sources sent to their owners and added BEFORE the division
            (``:125-127``);
  loss      ``binary_cross_entropy_with_logits(pos_weight)`` summed over the owned edges / global E (``train.py:143-144``);
  backward  torch autograd through the same exchanges run backwards (the gradient of a halo row goes back to its
            owner and is added there; the gradient of an owner's sum goes out to the ranks that contributed), BatchNorm
            backward with all-reduced ``(sum g, sum g * xhat)``;
  gradient all-reduce (one flat buffer) and the optimiser step.

Every layer is check-pointed: only its inputs ``(h, e)`` are kept and the layer is recomputed during backward (all ranks
recompute the same layer at the same point, so the collectives inside stay matched); without it the ~8 E x H
activations per layer (24 M x 256 x 4 B = 24.6 GB each at config 5) would not fit 8 x 180 GB.

The arithmetic goes through a ``prims`` object (``CudaPrims``: the autograd functions over the C ABI); the CPU ``gloo``
tests plug a plain-torch implementation of the same contracts in (tests/_emul.py) to check the distributed logic --
statistics, exchanges and their adjoints, loss scaling, gradient reduction -- against the oracle's autograd.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .partition import HaloPlan, Shard

EPS = 1e-6   # gated_gcn_full.py:114,127


# ------------------------------------------------------------------------------------------------
# differentiable halo exchanges (adjoint of each other)
# ------------------------------------------------------------------------------------------------
class _Exchange:
    def __init__(self, shard: Shard, plan: HaloPlan):
        self.plan, self.n_own, self.n_halo = plan, shard.n_own, shard.n_halo
        self.send_idx = plan.send_idx.long()

    def to_consumers(self, rows_own):
        """rows of owned nodes [n_own][W] -> the rows of this rank's halo nodes [n_halo][W]."""
        send = rows_own.index_select(0, self.send_idx).contiguous()
        return self.plan.to_consumers(send, rows_own.new_empty((self.n_halo, rows_own.shape[1])))

    def to_owners(self, rows_halo):
        """per-halo-node rows [n_halo][W] -> summed into the owners' rows [n_own][W]."""
        recv = self.plan.to_owners(rows_halo.contiguous(), rows_halo.new_empty((self.plan.n_send, rows_halo.shape[1])))
        return rows_halo.new_zeros((self.n_own, rows_halo.shape[1])).index_add_(0, self.send_idx, recv)


class HaloGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex: _Exchange, rows_own):
        ctx.ex = ex
        return ex.to_consumers(rows_own)

    @staticmethod
    def backward(ctx, g):
        return None, ctx.ex.to_owners(g)


class HaloScatterAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex: _Exchange, rows_halo):
        ctx.ex = ex
        return ex.to_owners(rows_halo)

    @staticmethod
    def backward(ctx, g):
        return None, ctx.ex.to_consumers(g)


# ------------------------------------------------------------------------------------------------
# train-mode BatchNorm with statistics over the rows of ALL ranks
# ------------------------------------------------------------------------------------------------
class DistBatchNormTrain(torch.autograd.Function):
    """``gnnome_b200.autograd.BatchNormTrain`` with the column sums all-reduced: y, global mean, global biased variance.
    The gradients it returns for weight / bias are this rank's PARTIAL sums (the gradient all-reduce completes them)."""

    @staticmethod
    def forward(ctx, prims, group, n_global, x, weight, bias, eps):
        n = float(n_global)
        st0 = prims.col_stats(x)
        _all_reduce(st0, group)
        mean_lo = (st0[0] / n).to(x.dtype).contiguous()                 # pass 1: mean (in the state's precision)
        st = prims.col_stats(x, shift_a=mean_lo)                        # pass 2: moments about it
        _all_reduce(st, group)
        mean = mean_lo.double() + st[0] / n
        var = (st[1] / n - (st[0] / n) ** 2).clamp_min_(0.0)
        rstd = torch.rsqrt(var + eps)
        a = weight.detach().double() * rstd
        c = bias.detach().double() - (mean - mean_lo.double()) * a
        y = prims.affine2(x, None, a, None, c, shift_x=mean_lo)
        ctx.save_for_backward(x, weight, mean, rstd, mean_lo)
        ctx.prims, ctx.group, ctx.n = prims, group, n
        mean_f, var_f = mean.to(x.dtype), var.to(x.dtype)
        ctx.mark_non_differentiable(mean_f, var_f)
        return y, mean_f, var_f

    @staticmethod
    def backward(ctx, g, _gm, _gv):
        x, weight, mean, rstd, mean_lo = ctx.saved_tensors
        prims, n = ctx.prims, ctx.n
        g = g.contiguous()
        st_local = prims.col_stats(g, x, shift_b=mean_lo)               # sum g, sum g * (x - mean_lo) over OWN rows
        st = st_local.clone()
        _all_reduce(st, ctx.group)
        dm = mean - mean_lo.double()
        c1, c2 = st[0], rstd * (st[1] - dm * st[0])                     # global sum g, sum g * xhat
        w = weight.detach().double()
        a = w * rstd
        b = -w * rstd * rstd * c2 / n
        c = -a * c1 / n + b * (mean_lo.double() - mean)
        gx = prims.affine2(g, x, a, b, c, shift_y=mean_lo)
        c2_local = rstd * (st_local[1] - dm * st_local[0])
        return None, None, None, gx, c2_local.to(weight.dtype), st_local[0].to(weight.dtype), None


def _all_reduce(t, group):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


# ------------------------------------------------------------------------------------------------
# layer checkpoint: keep (h, e), recompute the layer in backward
# ------------------------------------------------------------------------------------------------
class _LayerCheckpoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fn, h, e):
        ctx.fn = fn
        ctx.rng_cpu = torch.get_rng_state()
        ctx.rng_cuda = torch.cuda.get_rng_state(h.device) if h.is_cuda else None
        ctx.save_for_backward(h, e)
        with torch.no_grad():
            return fn(h, e, True)

    @staticmethod
    def backward(ctx, gh, ge):
        h, e = ctx.saved_tensors
        h_, e_ = h.detach().requires_grad_(True), e.detach().requires_grad_(True)
        with torch.random.fork_rng(devices=[h.device] if h.is_cuda else []):
            torch.set_rng_state(ctx.rng_cpu)
            if ctx.rng_cuda is not None:
                torch.cuda.set_rng_state(ctx.rng_cuda, h.device)
            with torch.enable_grad():
                h2, e2 = ctx.fn(h_, e_, False)          # False: the running statistics were updated the first time
        outs, grads = [], []
        for o, g in ((h2, gh), (e2, ge)):
            if g is not None:
                outs.append(o)
                grads.append(g)
        torch.autograd.backward(outs, grads)
        return None, h_.grad, e_.grad


# ------------------------------------------------------------------------------------------------
# the CUDA primitives
# ------------------------------------------------------------------------------------------------
class CudaPrims:
    """The product path: ``gnnome_b200.autograd``'s functions (hand-written forward / adjoint kernels, C ABI)."""

    def __init__(self, device):
        from . import autograd as A
        from .graph import GraphIndex
        self.A, self.GraphIndex, self.device = A, GraphIndex, device

    def stage(self, src_local, dst_local, n_local):
        return self.GraphIndex(src_local, dst_local, n_local, self.device)

    def position_eids(self, gi):
        return gi.in_eid[:gi.E].long()

    def gather_add3(self, gi, A_, B_, C_):
        return self.A.GatherAdd3.apply(gi, A_, B_, C_)

    def agg_in(self, gi, A_, sigma):
        return self.A.Agg.apply(gi, A_, sigma, 0)

    def agg_out_raw(self, gi, A_, sigma):
        return self.A.AggRaw.apply(gi, A_, sigma, 1)

    def gate(self, ehat, e_in):
        return self.A.Gate.apply(ehat, e_in)

    def col_stats(self, a, b=None, shift_a=None, shift_b=None):
        if a.shape[0] == 0:
            return torch.zeros((2, a.shape[1]), dtype=torch.float64, device=a.device)
        return self.A._col_stats(self.A._c(a), None if b is None else self.A._c(b), shift_a, shift_b)

    def affine2(self, x, y, a, b, c, shift_x=None, shift_y=None):
        if x.shape[0] == 0:
            return torch.empty_like(x)
        return self.A._affine2(self.A._c(x), None if y is None else self.A._c(y), a, b, c, shift_x, shift_y)

    def layer_norm(self, norm, x):
        return self.A.LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)

    def linear(self, lin, x):
        return self.A.linear(lin, x)


# ------------------------------------------------------------------------------------------------
# the sharded training step
# ------------------------------------------------------------------------------------------------
class ShardedTrainer:
    """``SymGatedGCNModel`` / ``GatedGCNModel`` (directed or not) in ``model.train()`` mode on this rank's shard.

    ``src, dst, x, e, y`` are the GLOBAL graph, features and edge labels (every rank passes the same; only the shard is
    kept on the device).  ``step(optimizer)`` runs forward, loss, backward, the gradient all-reduce and the optimiser
    step and returns the global mean BCE loss (a 0-dim tensor, identical on every rank)."""

    def __init__(self, model, src, dst, num_nodes, x, e, y, rank, world, device, pos_weight=None, prims=None, group=None,
                 dtype=torch.float32, checkpoint=True):
        device = torch.device(device)
        if device.type == 'cuda' and device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.model, self.rank, self.world, self.device, self.group = model, rank, world, device, group
        self.dtype, self.checkpoint = dtype, checkpoint
        self.p = prims if prims is not None else CudaPrims(self.device)
        self.sym = hasattr(model, 'linear1_node')
        if any(prm.device != self.device for prm in model.parameters()):
            raise RuntimeError(f'sharded training needs the model on {self.device}: call model.to(device) first')
        src, dst = torch.as_tensor(src), torch.as_tensor(dst)
        if self.device.type == 'cuda':
            src, dst = src.to(self.device), dst.to(self.device)
        # GatedGCNModel(directed=False), models/full_graph.py:47-52: the layers (and their edge BatchNorm statistics) run
        # on the graph with every edge doubled by its reverse (ids [E, 2E), the same encoded features, :49), the
        # predictor and the loss see the original edges (:51-52).  The DOUBLED graph is what gets partitioned; the
        # reversed copies an owner holds are scored too and carry weight 0 in the loss.
        self.undirected = not self.sym and not getattr(model, 'directed', True)
        e_orig = int(src.numel())
        if self.undirected:
            src, dst = torch.cat((src, dst)), torch.cat((dst, src))
        self.shard = sh = Shard(src, dst, num_nodes, rank, world)
        self.plan = HaloPlan(sh, self.device, group)
        self.ex = _Exchange(sh, self.plan)
        self.n_global, self.e_global, self.e_loss = int(num_nodes), int(src.numel()), e_orig
        self.gi = self.p.stage(sh.src_local.to(self.device), sh.dst_local.to(self.device), sh.n_local)
        order = self.p.position_eids(self.gi)                          # local edge id at every dst-sorted position
        self.order = order
        ids = sh.edge_ids
        feat = ids % max(e_orig, 1) if self.undirected else ids
        pick = lambda t: torch.as_tensor(t)[feat.to(torch.as_tensor(t).device)].to(device=self.device, dtype=dtype)  # noqa: E731
        self.x_own = torch.as_tensor(x)[sh.lo:sh.hi].to(device=self.device, dtype=dtype).contiguous()
        self.e_pos = pick(e)[order].contiguous()                       # edge rows in position order from here on
        self.y_pos = pick(y)[order].contiguous()
        self.w_pos = (ids < e_orig).to(device=self.device, dtype=dtype)[order].contiguous() if self.undirected else None
        self.pos_weight = None if pos_weight is None else torch.as_tensor(pos_weight, dtype=dtype, device=self.device)
        self.owned_edge_ids = ids

    # -- pieces ------------------------------------------------------------------------------------
    def _with_halo(self, rows_own):
        if not self.plan.active:
            return rows_own
        halo = HaloGather.apply(self.ex, rows_own)
        return torch.cat((rows_own, halo), 0) if self.shard.n_halo else rows_own

    def _norm(self, norm, x, n_global, updates, first):
        if isinstance(norm, torch.nn.LayerNorm):
            return self.p.layer_norm(norm, x)
        y, mean, var = DistBatchNormTrain.apply(self.p, self.group, n_global, x, norm.weight, norm.bias, norm.eps)
        if first and norm.track_running_stats:
            with torch.no_grad():
                unbiased = var * (n_global / max(n_global - 1, 1))
                for _ in range(updates):                               # gated_gcn_full.py:106 and :119 both call bn_e
                    norm.num_batches_tracked += 1
                    mom = norm.momentum if norm.momentum is not None else 1.0 / float(norm.num_batches_tracked)
                    norm.running_mean.mul_(1 - mom).add_(mean.to(norm.running_mean.dtype), alpha=mom)
                    norm.running_var.mul_(1 - mom).add_(unbiased.to(norm.running_var.dtype), alpha=mom)
        return y

    def _layer(self, conv, h_own, e, first):
        """layers/gated_gcn_full.py:82-142 (:182-230 without A_3) on the shard; edge rows in position order."""
        p, gi, sh = self.p, self.gi, self.shard
        n_own, H = sh.n_own, conv.out_channels
        h_loc = self._with_halo(h_own)
        A1h, A2h = p.linear(conv.A_1, h_own), p.linear(conv.A_2, h_loc)                        # :91-92
        B1h, B2h, B3e = p.linear(conv.B_1, h_loc), p.linear(conv.B_2, h_loc), p.linear(conv.B_3, e)   # :95-97
        z = p.gather_add3(gi, B1h, B2h, B3e)                                                   # :104-105
        ehat = self._norm(conv.bn_e, z, self.e_global, 2 if conv._symmetric else 1, first)    # :106 (+ :119)
        e_new, sigma = p.gate(ehat, e if conv.residual else None)                              # :107-111
        u = A1h + p.agg_in(gi, A2h, sigma)[:n_own]                                             # :112-114
        if conv._symmetric:                                                                    # :93, :125-127
            nd = torch.cat(p.agg_out_raw(gi, p.linear(conv.A_3, h_loc), sigma), 1)             # [n_local][2H], local edges
            tot = nd[:n_own]
            if self.plan.active:
                tot = tot + HaloScatterAdd.apply(self.ex, nd[n_own:])                          # partial sums of remote sources
            u = u + tot[:, :H] / (tot[:, H:] + EPS)
        u = self._norm(conv.bn_h, u, self.n_global, 1, first)                                  # :131-132
        h_new = torch.relu(u)                                                                  # :134
        if conv.residual:
            h_new = h_new + h_own                                                              # :136-137
        h_new = F.dropout(h_new, conv.dropout, training=True)                                  # :139
        return h_new, e_new

    def _predictor(self, pred, h_own, e):
        """layers/score_predictor.py:12-24 with W1 split into its src / dst / edge column blocks."""
        H = pred.in_features
        h_loc = self._with_halo(h_own)
        W1, b1 = pred.W1.weight, pred.W1.bias
        S1 = F.linear(h_loc, W1[:, :H])
        S2 = F.linear(h_loc, W1[:, H:2 * H], b1)
        E1 = F.linear(e, W1[:, 2 * H:])
        hid = torch.relu(self.p.gather_add3(self.gi, S1, S2, E1))
        return pred.W3(torch.relu(pred.W2(hid)))

    # -- the step ----------------------------------------------------------------------------------
    def forward(self):
        """Logits of the owned edges in dst-sorted POSITION order, (E_own, 1)."""
        m = self.model
        if not m.training:
            raise RuntimeError('ShardedTrainer runs model.train() mode (eval: gnnome_b200.partition.ShardedForward)')
        if self.sym:
            h = m.linear2_node(torch.relu(m.linear1_node(self.x_own)))                         # full_graph.py:26
            e = m.linear2_edge(torch.relu(m.linear1_edge(self.e_pos)))                         # :27
        else:
            h = m.node_encoder.linear2(torch.relu(m.node_encoder.linear1(self.x_own)))
            e = m.edge_encoder.linear2(torch.relu(m.edge_encoder.linear1(self.e_pos)))
        for conv in m.gnn.convs:                                                               # processor.py:16-19
            fn = lambda hh, ee, first, conv=conv: self._layer(conv, hh, ee, first)             # noqa: E731
            if self.checkpoint:
                h, e = _LayerCheckpoint.apply(fn, h, e)
            else:
                h, e = fn(h, e, True)
        return self._predictor(m.predictor, h, e)

    def loss(self, scores_pos):
        """train.py:143-144: mean BCE-with-logits over ALL edges; this rank's share (the shares add up to the loss)."""
        return F.binary_cross_entropy_with_logits(scores_pos.squeeze(-1), self.y_pos, weight=self.w_pos,
                                                  pos_weight=self.pos_weight, reduction='sum') / self.e_loss

    def reduce_gradients(self):
        """Sum the ranks' gradient contributions: one flat all-reduce over every parameter (missing gradients count
        as zeros so that all ranks reduce the same buffer)."""
        params = [prm for prm in self.model.parameters() if prm.requires_grad]
        flat = torch.cat([(prm.grad if prm.grad is not None else torch.zeros_like(prm)).reshape(-1) for prm in params])
        _all_reduce(flat, self.group)
        off = 0
        for prm in params:
            n = prm.numel()
            g = flat[off:off + n].view_as(prm)
            if prm.grad is None:
                prm.grad = g.clone()
            else:
                prm.grad.copy_(g)
            off += n

    def step(self, optimizer=None):
        self.model.zero_grad(set_to_none=True)
        share = self.loss(self.forward())
        share.backward()
        self.reduce_gradients()
        if optimizer is not None:
            optimizer.step()
        total = share.detach().clone()
        _all_reduce(total, self.group)
        return total

    def scores_in_edge_order(self, scores_pos):
        """(E_own, 1) logits ordered like ``owned_edge_ids`` (ascending global edge ids; for ``directed=False`` models the
        ids >= E are the reversed copies, whose scores the reference never looks at)."""
        return torch.empty_like(scores_pos).index_copy(0, self.order, scores_pos)
