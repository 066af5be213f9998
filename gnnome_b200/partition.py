"""Destination-node-range partitioning of the GatedGCN + edge-score pass over the GPUs of one box.

The reference has no distributed code; this is the one strategy the path
needs (section 8e).  One process per GPU (torchrun).  Rank r owns a contiguous range of nodes
``[lo_r, hi_r)`` (balanced by in-edge count), every edge whose DESTINATION lies in the range (its
``e`` row never leaves the rank: e'_k depends only on e_k, h[src_k], h[dst_k],
layers/gated_gcn_full.py:104-109) and computes ``h'`` for its nodes.  Per layer two exchanges:

  (1) src halo   owner -> consumer: the ``(B1h, A2h)`` rows of remote SOURCE nodes of owned edges
                 (gated_gcn_full.py:104,112 read ``B1h[src]`` / ``A2h[src]``);
  (2) reverse    consumer -> owner: the partial sums ``(sum sigma*A3h[dst], sum sigma)`` that owned
                 edges contribute to the out-edge aggregation of remote source nodes (:125-126); the
                 owner adds them -- in rank order, so the result is deterministic -- BEFORE the
                 division ``num / (den + 1e-6)`` (:127).

plus one src-halo exchange of the predictor's projected node rows (score_predictor.py:13).  All
exchanges are ``all_to_all_single`` over packed fp32 rows; the index tables are built once per graph.

The arithmetic is delegated to a ``kernels`` object (``CudaKernels``: the C-ABI calls of
``gnnome_b200.ops``); the CPU ``gloo`` tests plug a torch emulation of the same per-op contracts in
to check the partition / halo logic without a GPU.
"""
import os

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------
# index tables
# ------------------------------------------------------------------------------------------------
def node_bounds(dst, num_nodes, world):
    """``world + 1`` node boundaries that balance the number of in-edges per range."""
    if world == 1:
        return [0, int(num_nodes)]
    deg = torch.bincount(dst.long(), minlength=num_nodes)
    cum = torch.cumsum(deg, 0)
    total = int(cum[-1]) if num_nodes else 0
    targets = torch.tensor([total * k // world for k in range(1, world)], dtype=cum.dtype, device=cum.device)
    cuts = (torch.searchsorted(cum, targets, right=False) + 1).clamp_(max=num_nodes).tolist()
    bounds = [0] + cuts + [int(num_nodes)]
    for i in range(1, len(bounds)):                      # keep them monotone on degenerate inputs
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


class Shard:
    """The local view of rank ``rank``: owned nodes ``[0, n_own)``, then the halo source nodes."""

    def __init__(self, src, dst, num_nodes, rank, world, bounds=None):
        src, dst = src.long(), dst.long()
        self.rank, self.world, self.num_nodes_global = rank, world, int(num_nodes)
        self.bounds = bounds if bounds is not None else node_bounds(dst, num_nodes, world)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_own = self.hi - self.lo
        own = (dst >= self.lo) & (dst < self.hi)
        self.edge_ids = torch.nonzero(own).squeeze(1)                      # ascending global edge ids
        s, d = src[self.edge_ids], dst[self.edge_ids]
        remote = (s < self.lo) | (s >= self.hi)
        self.halo_nodes = torch.unique(s[remote])                          # sorted global ids
        self.n_halo = int(self.halo_nodes.numel())
        self.n_local = self.n_own + self.n_halo
        s_loc = s - self.lo
        if self.n_halo:
            s_loc = torch.where(remote, self.n_own + torch.searchsorted(self.halo_nodes, s), s_loc)
        self.src_local = s_loc.to(torch.int32)
        self.dst_local = (d - self.lo).to(torch.int32)
        self.num_edges = int(self.edge_ids.numel())
        # halo nodes grouped by owner (they are sorted, so the groups are contiguous in rank order)
        b = torch.tensor(self.bounds, dtype=torch.long, device=self.halo_nodes.device)
        owner = torch.searchsorted(b, self.halo_nodes, right=True) - 1
        self.recv_counts = torch.bincount(owner, minlength=world).tolist() if self.n_halo else [0] * world


class DistComm:
    """The collectives the sharded path needs, over ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests).
    ``tests/test_gpu_partition.py`` plugs an in-process implementation in to run k shards on ONE GPU."""

    def __init__(self, group=None):
        self.group = group

    def all_to_all(self, out, inp, out_splits=None, in_splits=None):
        dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=self.group)

    def all_to_all_start(self, out, inp, out_splits=None, in_splits=None):
        """The same exchange started asynchronously: the collective runs on the backend's own stream behind the work
        already queued on the current one; ``.wait()`` on the returned handle orders the current stream behind it."""
        return dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits,
                                      group=self.group, async_op=True)

    def all_reduce_max(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)


class _Done:
    def wait(self):
        return True


class HaloPlan:
    """Who sends which owned rows to whom.  ``send_idx`` (local owned node ids, grouped by peer rank)
    is obtained by exchanging the halo lists once."""

    def __init__(self, shard: Shard, device, group=None, comm=None):
        self.world, self.group = shard.world, group
        self.comm = comm if comm is not None else DistComm(group)
        self.recv_counts = list(shard.recv_counts)
        world = shard.world
        self.active = False          # does ANY rank exchange rows?  (a collective has to be entered by every rank)
        if world == 1:
            self.send_counts = [0]
            self.send_idx = torch.zeros(0, dtype=torch.int32, device=device)
            return
        cnt_in = torch.tensor(self.recv_counts, dtype=torch.int64, device=device)
        cnt_out = torch.empty(world, dtype=torch.int64, device=device)
        self.comm.all_to_all(cnt_out, cnt_in)
        self.send_counts = cnt_out.tolist()
        want = shard.halo_nodes.to(device=device, dtype=torch.int64).contiguous()
        asked = torch.empty(sum(self.send_counts), dtype=torch.int64, device=device)
        self.comm.all_to_all(asked, want, self.send_counts, self.recv_counts)
        if asked.numel() and (int(asked.min()) < shard.lo or int(asked.max()) >= shard.hi):
            raise RuntimeError('halo plan: a peer asked for a node this rank does not own')
        self.send_idx = (asked - shard.lo).to(torch.int32).contiguous()
        busy = torch.tensor([int(shard.n_halo > 0 or asked.numel() > 0)], dtype=torch.int64, device=device)
        self.comm.all_reduce_max(busy)
        self.active = bool(busy.item())
        # reverse exchange: rows arrive in send_idx order; group them per owned node (stable => rank order)
        order = torch.argsort(self.send_idx.long(), stable=True)
        self.xp_row = order.to(torch.int32).contiguous()
        cnt = torch.bincount(self.send_idx.long(), minlength=shard.n_own)
        self.xp_ptr = torch.zeros(shard.n_own + 1, dtype=torch.int32, device=device)
        self.xp_ptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)

    @property
    def n_send(self):
        return int(self.send_idx.numel())

    def to_consumers(self, rows_out, recv):
        """owner -> consumer: ``rows_out`` [n_send][W] (send_idx order) -> ``recv`` [n_halo][W]."""
        if self.world > 1:
            self.comm.all_to_all(recv, rows_out, self.recv_counts, self.send_counts)
        return recv

    def to_consumers_start(self, rows_out, recv):
        """``to_consumers`` without waiting: returns a handle whose ``wait()`` must be called before ``recv`` is read."""
        if self.world > 1:
            start = getattr(self.comm, 'all_to_all_start', None)
            if start is not None:
                return start(recv, rows_out, self.recv_counts, self.send_counts)
            self.comm.all_to_all(recv, rows_out, self.recv_counts, self.send_counts)
        return _Done()

    def to_owners(self, rows_out, recv):
        """consumer -> owner: ``rows_out`` [n_halo][W] -> ``recv`` [n_send][W] (send_idx order)."""
        if self.world > 1:
            self.comm.all_to_all(recv, rows_out, self.send_counts, self.recv_counts)
        return recv


# ------------------------------------------------------------------------------------------------
# the per-op contracts the sharded forward is written against
# ------------------------------------------------------------------------------------------------
class CudaKernels:
    """The product path: every method is one C-ABI call (gnnome_b200.ops).  Node / edge state objects are
    opaque to ShardedForward: on the split16 path h = (fp32 rows, fp16 images), e = fp16 images; on the
    fp32 path both are plain fp32 matrices."""

    def __init__(self, device):
        from . import ops
        from .graph import GraphIndex
        from .layers.gated_gcn import state_format
        self.ops, self.GraphIndex, self.device, self.state_format = ops, GraphIndex, device, state_format
        self.spare = {}

    def stage(self, src_local, dst_local, n_local):
        return self.GraphIndex(src_local, dst_local, n_local, self.device)

    def position_eids(self, gi):
        return gi.in_eid[:gi.E]

    def _split(self, H):
        return self.state_format(H) == 'split16'

    def encode_nodes(self, x, lin1, lin2, rows):
        from .layers.encoders import encode_rows, encode_rows2
        if self._split(lin2.out_features):
            h16, h32 = encode_rows2(x, None, lin1, lin2, rows, want32=True)
            return (h32, h16)
        return encode_rows(x, None, lin1, lin2, rows)

    def encode_edges(self, e, idx, lin1, lin2, rows):
        from .layers.encoders import encode_rows, encode_rows2
        if self._split(lin2.out_features):
            return encode_rows2(e, idx, lin1, lin2, rows)[0]
        return encode_rows(e, idx, lin1, lin2, rows)

    def width(self, h):
        return (h[0] if isinstance(h, tuple) else h).shape[1]

    def layer_pack(self, conv):
        return conv._pack(self.device, self._family(conv.out_channels))

    def node_linear_layer(self, pk, h, M, out):
        if isinstance(h, tuple):
            return self.ops.node_linear_tc2(h[1], pk['Wn_t'], pk['bn'], M, out=out)
        return self.ops.node_linear(h, pk['Wn_t'], pk['bn'], out=out)

    def gather_rows(self, table, idx, out=None):
        return self.ops.gather_rows(table, idx, out=out)

    def project_rows(self, pk, h, idx, M, out):
        """The first ``M`` columns of the layer's node projection for the rows ``idx`` only (the halo rows a peer is
        waiting for), bit-identical to the same rows of ``node_linear_layer``: a row's products do not depend on the
        tile it is computed in.  Split16 state only (None otherwise: the caller falls back to gathering from P)."""
        if not isinstance(h, tuple):
            return None
        h16 = h[1]                                                   # [rows][2][K] fp16 == [rows][K] 4-byte words
        rows16 = self.ops.gather_rows(h16.view(h16.shape[0], -1).view(torch.float32), idx,
                                      out=self._scratch('h16_send', (idx.numel(), h16.shape[2])))
        return self.ops.node_linear_tc2(rows16.view(torch.float16).view(idx.numel(), 2, h16.shape[2]), pk['Wn_t'], pk['bn'],
                                        M, out=out)

    def _scratch(self, name, shape):
        t = self.spare.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.spare[name] = torch.empty(shape, dtype=torch.float32, device=self.device)
        return t

    def edge_forward(self, gi, H, P, pk, e_pos, F, carry, flags):
        if e_pos.dtype == torch.float16:
            self.ops.edge_forward_tc2(gi, H, P, pk['We_t'], e_pos, F, carry, flags)
        else:
            self.ops.edge_forward(gi, H, P, pk['We_t'], pk['scale_e'], pk['shift_e'], e_pos, F, carry, flags)

    def _family(self, H):
        return 'tc2' if self._split(H) else 'ffma'

    def carry_shape(self, gi, H):
        return (gi.num_chunks(H, self._family(H)), 4, H)

    def reverse_partial(self, gi, H, P, e_pos, node_begin, node_end, out):
        fn = self.ops.reverse_partial2 if e_pos.dtype == torch.float16 else self.ops.reverse_partial
        fn(gi, H, P, e_pos, node_begin, node_end, out)

    def _like(self, name, t):
        b = self.spare.pop(name, None)
        if b is None or b.shape != t.shape or b.dtype != t.dtype or b.data_ptr() == t.data_ptr():
            b = torch.empty_like(t)
        return b

    def node_update(self, gi, H, P, pk, e_pos, F, carry, h_in, flags, n_own, xp_ptr, xp_row, xp_buf):
        """Returns the new node state; the input state's buffers are recycled for the next layer."""
        kw = dict(node_end=n_own, xp_ptr=xp_ptr, xp_row=xp_row, xp_buf=xp_buf)
        if isinstance(h_in, tuple):
            h32, h16 = h_in
            o32, o16 = self._like('h32', h32), self._like('h16', h16)
            self.ops.node_update2(gi, H, P, e_pos, F, carry, h32, pk['scale_h'], pk['shift_h'], o32, o16, flags,
                                  gi.chunk(H, 'tc2'), **kw)
            self.spare['h32'], self.spare['h16'] = h32, h16
            return (o32, o16)
        out = self._like('h32', h_in)
        self.ops.node_update(gi, H, P, e_pos, F, carry, h_in, pk['scale_h'], pk['shift_h'], out, flags,
                             gi.chunk(H, 'ffma'), **kw)
        self.spare['h32'] = h_in
        return out

    def score_node_rows(self, predictor, h):
        return predictor.node_rows16(h[1]) if isinstance(h, tuple) else predictor.node_rows(h)

    def score_forward(self, predictor, gi, S, e_pos, scores):
        if e_pos.dtype == torch.float16:
            return predictor.score_positions16(gi, S, e_pos, scores)
        return predictor.score_positions(gi, S, e_pos, scores)


# ------------------------------------------------------------------------------------------------
# the sharded forward
# ------------------------------------------------------------------------------------------------
class ShardedForward:
    """``SymGatedGCNModel`` / ``GatedGCNModel(directed=True)`` forward (eval) on this rank's shard.

    ``step()`` returns the (E_own, 1) logits of the owned edges, ordered like ``owned_edge_ids``
    (ascending global DGL edge ids).  ``GatedGCNModel(directed=False)`` is sharded over its doubled graph.  Inputs are the GLOBAL graph / features (every rank passes the
    same); only the shard is kept on the device."""

    def __init__(self, model, src, dst, num_nodes, x, e, rank, world, device, kernels=None, group=None,
                 dtype=torch.float32, comm=None):
        self.model, self.rank, self.world, self.device, self.group = model, rank, world, device, group
        self.dtype = dtype   # fp32 on the CUDA path; the CPU emulation in the tests runs fp64
        self.k = kernels if kernels is not None else CudaKernels(device)
        if model.training:
            raise NotImplementedError('ShardedForward runs eval mode')
        self.sym = hasattr(model, 'linear1_node')
        src, dst = torch.as_tensor(src), torch.as_tensor(dst)
        if torch.device(device).type == 'cuda':          # build the index tables on the GPU (sort / unique of E ids)
            src, dst = src.to(device), dst.to(device)
        # GatedGCNModel(directed=False), models/full_graph.py:47-52: the layers run on the graph with every edge doubled
        # by its reverse (ids [E, 2E), same input features) and the predictor scores the original edges.  Here the
        # DOUBLED graph is what gets partitioned; a rank scores every edge it owns and returns the original ones.
        self.undirected = not self.sym and not getattr(model, 'directed', True)
        self.num_edges_orig = int(src.numel())
        if self.undirected:
            src, dst = torch.cat((src, dst)), torch.cat((dst, src))
        self.shard = sh = Shard(src, dst, num_nodes, rank, world)
        self.plan = HaloPlan(sh, device, group, comm)
        self.owned_edge_ids = sh.edge_ids
        self._keep = None
        if self.undirected:
            self._keep = torch.nonzero(sh.edge_ids < self.num_edges_orig).squeeze(1).to(device)
            self.owned_edge_ids = sh.edge_ids[self._keep.to(sh.edge_ids.device)]
        self.gi = self.k.stage(sh.src_local.to(device), sh.dst_local.to(device), sh.n_local)
        self.x_own = torch.as_tensor(x)[sh.lo:sh.hi].to(device=device, dtype=dtype).contiguous()
        e = torch.as_tensor(e)
        feat_ids = sh.edge_ids % max(self.num_edges_orig, 1) if self.undirected else sh.edge_ids   # :49 e = cat(e, e)
        self.e_own = e[feat_ids.to(e.device)].to(device=device, dtype=dtype).contiguous()
        del src, dst
        self._host = None
        self.ws = {}
        # project + send the halo rows ahead of the full projection (False / GNB_OVERLAP=0: the serial schedule)
        self.overlap = os.environ.get('GNB_OVERLAP', '1') != '0'

    def local_sizes(self):
        return self.shard.n_local, self.shard.num_edges

    def _buf(self, name, shape):
        t = self.ws.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.ws[name] = torch.empty(shape, dtype=self.dtype, device=self.device)
        return t

    def forward(self, gi, x_own, e_own):
        m, k, sh, plan = self.model, self.k, self.shard, self.plan
        n_own, n_halo, n_local = sh.n_own, sh.n_halo, sh.n_local
        if self.sym:
            lin_n, lin_e = (m.linear1_node, m.linear2_node), (m.linear1_edge, m.linear2_edge)
        else:
            lin_n = (m.node_encoder.linear1, m.node_encoder.linear2)
            lin_e = (m.edge_encoder.linear1, m.edge_encoder.linear2)
        h = k.encode_nodes(x_own, lin_n[0], lin_n[1], n_own)
        e_pos = k.encode_edges(e_own, k.position_eids(gi), lin_e[0], lin_e[1], sh.num_edges)
        H = k.width(h)
        nb = 5 if self.sym else 4
        P = self._buf('P', (n_local, nb * H))
        Fb = self._buf('F', (n_local, H))
        for conv in m.gnn.convs:
            pk = k.layer_pack(conv)
            flags = conv._flags()
            carry = self._buf('carry', k.carry_shape(gi, H))
            # exchange (1).  The (B1h', A2h) rows the peers are waiting for are projected FIRST, from the few h rows
            # they come from, and travel while the whole table is projected (the exchange used to sit between the
            # projection and the edge pass: ~0.9 ms per layer of an 8-GPU step with nothing to hide it behind).
            early = None
            if plan.active and self.overlap and hasattr(k, 'project_rows'):
                early = k.project_rows(pk, h, plan.send_idx, 2 * H, self._buf('s1', (plan.n_send, 2 * H)))
            if early is not None:
                pending = plan.to_consumers_start(early, self._buf('r1', (n_halo, 2 * H)))
                k.node_linear_layer(pk, h, nb * H, P[:n_own])
                pending.wait()
                P[n_own:, :2 * H].copy_(self.ws['r1'])
            else:
                k.node_linear_layer(pk, h, nb * H, P[:n_own])
                if plan.active:
                    out_rows = k.gather_rows(P[:n_own, :2 * H], plan.send_idx, self._buf('s1', (plan.n_send, 2 * H)))
                    halo = plan.to_consumers(out_rows, self._buf('r1', (n_halo, 2 * H)))
                    P[n_own:, :2 * H].copy_(halo)
            k.edge_forward(gi, H, P, pk, e_pos, Fb, carry, flags)
            # exchange (2) stays between the two kernels it connects.  Hiding it was tried (all owned nodes updated
            # without the remote sums while they travel, the nodes that receive some computed again from a node list):
            # bit-identical, but the second pass over those scattered nodes costs more than the exchange it hides
            # (5.7 ms against ~1 ms per layer at 2 GPUs on the bench graph, profiles/r02h).
            xp_buf = None
            if self.sym and plan.active:
                part = self._buf('s2', (n_halo, 2 * H))
                k.reverse_partial(gi, H, P, e_pos, n_own, n_local, part)
                xp_buf = plan.to_owners(part, self._buf('r2', (plan.n_send, 2 * H)))
            if plan.n_send == 0:                                             # took part in the exchange, received nothing
                xp_buf = None
            h = k.node_update(gi, H, P, pk, e_pos, Fb, carry, h, flags, n_own,
                              plan.xp_ptr if xp_buf is not None else None,
                              plan.xp_row if xp_buf is not None else None, xp_buf)
        # ---- predictor: S = [x W1s^T | x W1d^T + b1]; the src half of remote sources is a halo ----
        S_own = k.score_node_rows(m.predictor, h)
        hs = S_own.shape[1] // 2
        S = self._buf('S', (n_local, 2 * hs))
        S[:n_own].copy_(S_own)
        if plan.active:
            out_rows = k.gather_rows(S_own[:, :hs], plan.send_idx, self._buf('s3', (plan.n_send, hs)))
            S[n_own:, :hs].copy_(plan.to_consumers(out_rows, self._buf('r3', (n_halo, hs))))
        scores = torch.empty((sh.num_edges, 1), dtype=self.dtype, device=self.device)
        k.score_forward(m.predictor, gi, S, e_pos, scores)
        return scores if self._keep is None else scores[self._keep]

    def step(self):
        return self.forward(self.gi, self.x_own, self.e_own)

    # ---- end to end from pinned host buffers (bench.py `e2e`) ------------------------------------
    def e2e_prepare(self):
        sh = self.shard
        pin = lambda t: t.cpu().contiguous().pin_memory()
        self._host = dict(src=pin(sh.src_local), dst=pin(sh.dst_local), x=pin(self.x_own), e=pin(self.e_own),
                          out=torch.empty((int(self.owned_edge_ids.numel()), 1), dtype=torch.float32).pin_memory())
        return sum(self._host[n].numel() * self._host[n].element_size() for n in ('src', 'dst', 'x', 'e')), \
            self._host['out'].numel() * 4

    def e2e_step(self):
        """H2D of this rank's shard (local edge list, x, e), graph staging, forward, D2H of the scores."""
        hb, dev = self._host, self.device
        gi = self.k.stage(hb['src'].to(dev, non_blocking=True), hb['dst'].to(dev, non_blocking=True),
                          self.shard.n_local)
        out = self.forward(gi, hb['x'].to(dev, non_blocking=True), hb['e'].to(dev, non_blocking=True))
        hb['out'].copy_(out, non_blocking=True)
        return hb['out']


def gather_scores(runner: ShardedForward, scores, num_edges_global, dst_rank=0):
    """Collect the per-rank logits into global DGL edge-id order on ``dst_rank`` (None elsewhere)."""
    world = runner.world
    if world == 1:
        return scores
    counts = [None] * world
    dist.all_gather_object(counts, int(scores.shape[0]), group=runner.group)
    cap = max(counts)                                     # equal-sized (padded) buffers: gloo has no all_gather_v
    my_ids = torch.full((cap,), -1, dtype=torch.int64, device=scores.device)
    my_ids[:scores.shape[0]] = runner.owned_edge_ids.to(device=scores.device, dtype=torch.int64)
    my_vals = torch.zeros((cap, 1), dtype=scores.dtype, device=scores.device)
    my_vals[:scores.shape[0]] = scores
    ids = [torch.empty_like(my_ids) for _ in range(world)]
    vals = [torch.empty_like(my_vals) for _ in range(world)]
    dist.all_gather(ids, my_ids, group=runner.group)
    dist.all_gather(vals, my_vals, group=runner.group)
    ids = [t[:c] for t, c in zip(ids, counts)]
    vals = [t[:c] for t, c in zip(vals, counts)]
    if runner.rank != dst_rank:
        return None
    out = torch.empty((num_edges_global, 1), dtype=scores.dtype, device=scores.device)
    out[torch.cat(ids)] = torch.cat(vals)
    return out
