"""Straight-line CPU restatement of GNNome's GatedGCN + edge-score path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the checker for the CUDA path and the timed
``cpu_baseline`` of ``bench.py``.  It is *functional*: parameters come in as a plain
``state_dict``-style mapping (the reference's key names), the graph as ``(src, dst, N)`` index
tensors, so it shares no code with the product package.

Every function cites the reference lines it restates (paths relative to the reference root).
``dtype=torch.float64`` gives the "true value" used to bound the fp32 error of both sides.
"""
import torch
import torch.nn.functional as F

EPS_GATE = 1e-6  # layers/gated_gcn_full.py:114,127


def _lin(p, name, x):
    return F.linear(x, p[name + '.weight'], p[name + '.bias'])


def _norm(p, name, x, normalization, training, bn_momentum=0.1):
    """layers/gated_gcn_full.py:37-42 -- BatchNorm1d(track_running_stats=True) or LayerNorm."""
    if normalization == 'batch':
        rm, rv = p.get(name + '.running_mean'), p.get(name + '.running_var')
        out = F.batch_norm(x, rm, rv, p[name + '.weight'], p[name + '.bias'],
                           training=training, momentum=bn_momentum, eps=1e-5)
        nbt = p.get(name + '.num_batches_tracked')
        if training and nbt is not None:
            nbt += 1
        return out
    if normalization == 'layer':
        return F.layer_norm(x, (x.shape[-1],), p[name + '.weight'], p[name + '.bias'], 1e-5)
    raise ValueError(normalization)


def _segment_sum(values, index, n):
    """dgl update_all(..., fn.sum): zero-filled sum over edges grouped by ``index``."""
    out = torch.zeros((n,) + tuple(values.shape[1:]), dtype=values.dtype, device=values.device)
    out.index_add_(0, index, values)
    return out


def sym_gated_gcn_layer(p, prefix, src, dst, n, h, e, normalization='batch', training=False,
                        dropout=0.0, residual=True, faithful=True):
    """layers/gated_gcn_full.py:82-142 (SymGatedGCN.forward).

    ``faithful=True`` repeats the reference's op sequence (the gate is evaluated a second time
    on the reversed graph, ``:117-124``, including the second ``bn_e`` call that updates the
    running statistics twice per layer); ``faithful=False`` uses the single-sigma form that the
    CUDA path implements (bitwise the same values in eval mode).
    """
    q = lambda s: f'{prefix}{s}'
    h_in, e_in = h, e                                              # :85-86
    A1h, A2h, A3h = _lin(p, q('A_1'), h), _lin(p, q('A_2'), h), _lin(p, q('A_3'), h)  # :91-93
    B1h, B2h, B3e = _lin(p, q('B_1'), h), _lin(p, q('B_2'), h), _lin(p, q('B_3'), e)  # :95-97

    e_ji = B1h[src] + B2h[dst] + B3e                               # :104-105
    e_ji = F.relu(_norm(p, q('bn_e'), e_ji, normalization, training))  # :106-107
    if residual:
        e_ji = e_ji + e_in                                         # :108-109
    sigma_f = torch.sigmoid(e_ji)                                  # :111
    num_f = _segment_sum(A2h[src] * sigma_f, dst, n)               # :112
    den_f = _segment_sum(sigma_f, dst, n)                          # :113
    h_forward = num_f / (den_f + EPS_GATE)                         # :114

    if faithful:
        e_ik = B2h[dst] + B1h[src] + B3e                           # :117-118 (reverse graph: u=dst, v=src)
        e_ik = F.relu(_norm(p, q('bn_e'), e_ik, normalization, training))  # :119-120
        if residual:
            e_ik = e_ik + e_in                                     # :121-122
        sigma_b = torch.sigmoid(e_ik)                              # :124
    else:
        sigma_b = sigma_f
    num_b = _segment_sum(A3h[dst] * sigma_b, src, n)               # :125
    den_b = _segment_sum(sigma_b, src, n)                          # :126
    h_backward = num_b / (den_b + EPS_GATE)                        # :127

    h = A1h + h_forward + h_backward                               # :129
    if normalization != 'none':
        h = _norm(p, q('bn_h'), h, normalization, training)        # :131-132
    h = F.relu(h)                                                  # :134
    if residual:
        h = h + h_in                                               # :136-137
    h = F.dropout(h, dropout, training=training)                   # :139
    return h, e_ji                                                 # :140-142


def gated_gcn_layer(p, prefix, src, dst, n, h, e, normalization='batch', training=False,
                    dropout=0.0, residual=True):
    """layers/gated_gcn_full.py:182-230 (GatedGCN.forward): no A_3, no reverse aggregation."""
    q = lambda s: f'{prefix}{s}'
    h_in, e_in = h, e
    A1h, A2h = _lin(p, q('A_1'), h), _lin(p, q('A_2'), h)          # :194-195
    B1h, B2h, B3e = _lin(p, q('B_1'), h), _lin(p, q('B_2'), h), _lin(p, q('B_3'), e)  # :198-200
    e_ji = B1h[src] + B2h[dst] + B3e                               # :205-206
    e_ji = F.relu(_norm(p, q('bn_e'), e_ji, normalization, training))  # :207-208
    if residual:
        e_ji = e_ji + e_in                                         # :209-210
    sigma_f = torch.sigmoid(e_ji)                                  # :212
    h_forward = _segment_sum(A2h[src] * sigma_f, dst, n) / (_segment_sum(sigma_f, dst, n) + EPS_GATE)  # :213-215
    h = A1h + h_forward                                            # :217
    if normalization != 'none':
        h = _norm(p, q('bn_h'), h, normalization, training)        # :219-220
    h = F.relu(h)
    if residual:
        h = h + h_in
    h = F.dropout(h, dropout, training=training)
    return h, e_ji


def score_predictor(p, prefix, src, dst, x, e):
    """layers/score_predictor.py:12-24."""
    data = torch.cat((x[src], x[dst], e), dim=1)                   # :13
    hdn = torch.relu(_lin(p, prefix + 'W1', data))                 # :14-15
    return _lin(p, prefix + 'W3', torch.relu(_lin(p, prefix + 'W2', hdn)))  # :16


def _cast_params(state_dict, dtype):
    out = {}
    for k, v in state_dict.items():
        out[k] = v.to(dtype) if torch.is_floating_point(v) else v.clone()
    return out


def num_layers_of(state_dict):
    idx = [int(k.split('.')[2]) for k in state_dict if k.startswith('gnn.convs.')]
    return max(idx) + 1 if idx else 0


def model_forward(state_dict, src, dst, n, x, e, model='sym', normalization='batch', directed=True,
                  training=False, dropout=0.0, dtype=torch.float32, faithful=True,
                  return_layers=False, cast=True):
    """models/full_graph.py:22-30 (SymGatedGCNModel) and :42-53 (GatedGCNModel).

    ``state_dict`` uses the reference's key names.  With ``cast=False`` the mapping is used as
    is (leaf tensors with ``requires_grad`` for gradient oracles; BN buffers updated in place).
    """
    p = _cast_params(state_dict, dtype) if cast else state_dict
    src, dst = src.long(), dst.long()
    x, e = x.to(dtype), e.to(dtype)
    n_layers = num_layers_of(p)
    per_layer = []
    if model == 'sym':
        h = _lin(p, 'linear2_node', torch.relu(_lin(p, 'linear1_node', x)))   # full_graph.py:26
        ee = _lin(p, 'linear2_edge', torch.relu(_lin(p, 'linear1_edge', e)))  # :27
        for i in range(n_layers):                                              # processor.py:16-19
            h, ee = sym_gated_gcn_layer(p, f'gnn.convs.{i}.', src, dst, n, h, ee, normalization,
                                        training, dropout, faithful=faithful)
            if return_layers:
                per_layer.append((h, ee))
        scores = score_predictor(p, 'predictor.', src, dst, h, ee)             # :29
    elif model == 'gated':
        h = _lin(p, 'node_encoder.linear2', torch.relu(_lin(p, 'node_encoder.linear1', x)))  # node_encoder.py:28-33
        ee = _lin(p, 'edge_encoder.linear2', torch.relu(_lin(p, 'edge_encoder.linear1', e)))  # edge_encoder.py:27-32
        if directed:
            gs, gd = src, dst                                                  # full_graph.py:45-46
        else:
            gs, gd = torch.cat((src, dst)), torch.cat((dst, src))              # :48 add_reverse_edges
            ee = torch.cat((ee, ee), dim=0)                                    # :49
        for i in range(n_layers):                                              # processor.py:29-32
            h, ee = gated_gcn_layer(p, f'gnn.convs.{i}.', gs, gd, n, h, ee, normalization,
                                    training, dropout)
            if return_layers:
                per_layer.append((h, ee))
        if not directed:
            ee = ee[:src.numel()]                                              # :51
        scores = score_predictor(p, 'predictor.', src, dst, h, ee)             # :52
    else:
        raise ValueError(model)
    if return_layers:
        return scores, per_layer
    return scores


def init_state_dict(model='sym', node_features=2, edge_features=2, hidden=64, hidden_ne=16,
                    num_layers=8, hidden_edge_scores=64, normalization='batch', seed=0):
    """Random-init parameters with the reference's key names and shapes (models/full_graph.py:10-20,
    :34-40; layers/gated_gcn_full.py:28-42) drawn from torch's default nn.Linear distributions
    (weight and bias ~ U(+-1/sqrt(fan_in))); BN/LN affine = 1/0, running stats = 0/1."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def linear(name, fin, fout):
        bound = 1.0 / fin ** 0.5
        w = torch.empty(fout, fin).uniform_(-bound, bound, generator=g)
        b = torch.empty(fout).uniform_(-bound, bound, generator=g)
        sd[name + '.weight'], sd[name + '.bias'] = w, b

    def norm(name, c):
        sd[name + '.weight'], sd[name + '.bias'] = torch.ones(c), torch.zeros(c)
        if normalization == 'batch':
            sd[name + '.running_mean'] = torch.zeros(c)
            sd[name + '.running_var'] = torch.ones(c)
            sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    if model == 'sym':
        linear('linear1_node', node_features, hidden_ne)
        linear('linear2_node', hidden_ne, hidden)
        linear('linear1_edge', edge_features, hidden_ne)
        linear('linear2_edge', hidden_ne, hidden)
        lins = ('A_1', 'A_2', 'A_3', 'B_1', 'B_2', 'B_3')
    else:
        linear('node_encoder.linear1', node_features, hidden_ne)
        linear('node_encoder.linear2', hidden_ne, hidden)
        linear('edge_encoder.linear1', edge_features, hidden_ne)
        linear('edge_encoder.linear2', hidden_ne, hidden)
        lins = ('A_1', 'A_2', 'B_1', 'B_2', 'B_3')
    for i in range(num_layers):
        for l in lins:
            linear(f'gnn.convs.{i}.{l}', hidden, hidden)
        norm(f'gnn.convs.{i}.bn_h', hidden)
        norm(f'gnn.convs.{i}.bn_e', hidden)
    linear('predictor.W1', 3 * hidden, hidden_edge_scores)
    linear('predictor.W2', hidden_edge_scores, 32)
    linear('predictor.W3', 32, 1)
    return sd


# ---- the callers' code either side of model(g, x, e) ---------------------------------

def zscore(v):
    """``(v - v.mean()) / v.std()`` with torch's unbiased std (utils/data_utils.py:36, train.py:113-116)."""
    return (v - v.mean()) / v.std()


def edge_input_features(overlap_length, overlap_similarity, use_similarities=True):
    """utils/data_utils.py:31-41 (``preprocess_graph``): e = [z(overlap_length.float()), overlap_similarity]."""
    ol_len = zscore(overlap_length.float())
    if not use_similarities:
        return ol_len.unsqueeze(-1)
    return torch.cat((ol_len.unsqueeze(-1), overlap_similarity.unsqueeze(-1)), dim=1)


def node_input_features(src, dst, n, reverse=False):
    """utils/data_utils.py:50-51 (float in/out degrees) + train.py:112-120 / inference.py:413-420 (z-score, concatenate;
    ``reverse`` swaps the columns: train.py:117-118)."""
    in_deg = torch.bincount(dst.long(), minlength=n).float().unsqueeze(1)
    out_deg = torch.bincount(src.long(), minlength=n).float().unsqueeze(1)
    pe_in, pe_out = zscore(in_deg), zscore(out_deg)
    return torch.cat((pe_out, pe_in) if reverse else (pe_in, pe_out), dim=1)


def symmetry_loss(org_scores, rev_scores, labels, pos_weight, alpha):
    """train.py:103-109."""
    bce = lambda s: F.binary_cross_entropy_with_logits(s, labels, pos_weight=pos_weight, reduction='none')  # noqa: E731
    return (bce(org_scores) + bce(rev_scores) + alpha * torch.abs(org_scores - rev_scores)).mean()


def node_subgraph_ids(src, dst, n, keep):
    """``dgl.node_subgraph(g, keep, store_ids=True)`` as train.py:95 uses it (DGL 0.8 documented behaviour): kept nodes
    renumbered in increasing id order, induced edges in the order of their ids.
    Returns (node_id, edge_id, sub_src, sub_dst) as int64 tensors."""
    keep = keep.bool()
    node_id = torch.nonzero(keep).flatten()
    new_id = torch.cumsum(keep.long(), 0) - 1
    s, d = src.long(), dst.long()
    edge_id = torch.nonzero(keep[s] & keep[d]).flatten()
    return node_id, edge_id, new_id[s[edge_id]], new_id[d[edge_id]]


def partition_node_features(in_deg, out_deg, node_id, reverse=False):
    """train.py:125-133: the parent's degrees at the batch's ``_ID``, z-scored within the batch."""
    pe_in, pe_out = zscore(in_deg[node_id].unsqueeze(1)), zscore(out_deg[node_id].unsqueeze(1))
    return torch.cat((pe_out, pe_in) if reverse else (pe_in, pe_out), dim=1)


# ---- greedy decoder walks -- pure-Python loops, small cases only ---------------------------

def greedy_walk(start, log_probs, succs, edges, visited_old):
    """inference.py:70-114 (``greedy_forwards``; ``greedy_backwards_rc`` :117-161 is the same loop from ``start ^ 1``)
    with ``RANDOM`` and ``early_stopping`` off (:25-28).  Returns (walk, visited, float32 sum of log-probabilities)."""
    current, walk, visited = start, [], set()
    total = torch.zeros(1, dtype=torch.float32)
    while True:
        walk.append(current)
        visited.update((current, current ^ 1))
        nbrs = succs.get(current, [])
        if len(nbrs) == 1:
            if nbrs[0] in visited_old or nbrs[0] in visited:
                break
            total += log_probs[edges[current, nbrs[0]]]
            current = nbrs[0]
            continue
        free = [v for v in nbrs if not (v in visited_old or v in visited)]
        if not free:
            break
        p = log_probs[[edges[current, v] for v in free]]
        best = int(torch.argmax(p))            # torch.topk(k=1): first of equal maxima for short lists
        total += p[best]
        current = free[best]
    return walk, visited, total


def run_greedy_both_ways(src, dst, log_probs, succs, edges, visited):
    """inference.py:164-168 -> (walk_f, walk_b, sumLogProb_f, sumLogProb_b)."""
    tmp = set(visited) | {src, src ^ 1, dst, dst ^ 1}
    walk_f, vis_f, sum_f = greedy_walk(dst, log_probs, succs, edges, tmp)
    walk_b, _, sum_b = greedy_walk(src ^ 1, log_probs, succs, edges, tmp | vis_f)
    return walk_f, [w ^ 1 for w in reversed(walk_b)], sum_f, sum_b


def contig_length(walk, edges, prefix_length, read_length):
    """inference.py:30-37 (``graph.edges[u, v].data['prefix_length']`` looked up through the ``edges`` dict)."""
    return int(sum(int(prefix_length[edges[u, v]]) for u, v in zip(walk[:-1], walk[1:])) + int(read_length[walk[-1]]))


def jumped_nodes(walk, succs, preds):
    """inference.py:316-322: nodes between consecutive walk nodes (successor of one, predecessor of the next), both strands."""
    trans = set()
    for ss, dd in zip(walk[:-1], walk[1:]):
        t1 = set(succs.get(ss, [])) & set(preds.get(dd, []))
        trans |= t1 | {t ^ 1 for t in t1}
    return trans
