"""Greedy decoder walks -- the step after the scoring pass (reference inference.py:70-164,
231-322).  Host-side: the walks are sequential pointer chasing over a few successors per node, so they run in C++ on the
CPU (``csrc/gnb_walks.cu``), the ``nb_paths`` candidates of one iteration on a pool of threads.

``WalkGraph`` holds the reference's ``succs`` / ``preds`` / ``edges`` structures ({idx}_succ.pkl, {idx}_pred.pkl,
{idx}_edges.pkl, inference.py:445-455) as CSR arrays and offers the reference's functions on them:

* ``run_greedy_both_ways`` for a whole batch of start edges (inference.py:164-168 inside the loop of :231-247),
* ``get_contig_length`` (:30-37), and the jumped-over nodes of an accepted walk (:316-322).

Results are identical to the reference's: the same walks node for node, and the same float32 ``sumLogProb`` bit for bit
(``RANDOM`` and ``early_stopping`` are off, as shipped).  Exact ties between successors go to the first in list order,
which is torch's choice for up to 16 candidates."""
import ctypes

import numpy as np
import torch

from . import _lib


def _np(t, dtype):
    if torch.is_tensor(t):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t), dtype=dtype)


def _csr_from_dicts(num_nodes, lists, edge_of=None):
    """CSR of ``lists`` (node -> list of neighbours, list order kept) and, with ``edge_of`` ((u, v) -> id), the edge id
    of every entry.  The pair look-up is a sort + binary search over the dict's keys instead of one dict access per
    entry (60 M of them on a 10 M-node graph)."""
    from itertools import chain
    empty = ()
    counts = np.fromiter((len(lists.get(i, empty)) for i in range(num_nodes)), dtype=np.int64, count=num_nodes)
    ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    total = int(ptr[-1])
    node = np.fromiter(chain.from_iterable(lists.get(i, empty) for i in range(num_nodes)), dtype=np.int64, count=total)
    edge = np.zeros(total, dtype=np.int32)
    if edge_of is not None and total:
        m = len(edge_of)
        keys = np.fromiter(chain.from_iterable(edge_of.keys()), dtype=np.int64, count=2 * m).reshape(m, 2)
        vals = np.fromiter(edge_of.values(), dtype=np.int64, count=m)
        flat = keys[:, 0] * num_nodes + keys[:, 1]
        order = np.argsort(flat, kind='stable')
        want = np.repeat(np.arange(num_nodes, dtype=np.int64), counts) * num_nodes + node
        pos = np.searchsorted(flat[order], want)
        if (pos >= m).any() or (flat[order][np.minimum(pos, m - 1)] != want).any():
            bad = int(np.nonzero((pos >= m) | (flat[order][np.minimum(pos, m - 1)] != want))[0][0])
            raise KeyError((int(want[bad] // num_nodes), int(want[bad] % num_nodes)))
        edge = vals[order][pos].astype(np.int32)
    return ptr, node.astype(np.int32), edge


def _csr_from_edges(key_nodes, other_nodes, eids, num_nodes):
    order = np.argsort(key_nodes, kind='stable')               # lists in edge-id order
    ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(key_nodes, minlength=num_nodes), out=ptr[1:])
    return ptr, other_nodes[order].astype(np.int32), eids[order].astype(np.int32)


class WalkGraph:
    """CSR view of ``succs`` (node -> list of successors), ``preds`` and ``edges`` ((u, v) -> edge id)."""

    def __init__(self, num_nodes, succ_csr, pred_csr=None):
        self.N = int(num_nodes)
        if self.N % 2:
            raise ValueError('nodes come in strand pairs (2k, 2k + 1): num_nodes must be even')
        self._succ, self._pred = succ_csr, pred_csr
        self.max_edge_id = -1
        for csr in (succ_csr, pred_csr):
            if csr is None:
                continue
            ptr, node, edge = csr
            if ptr.shape != (self.N + 1,) or int(ptr[-1]) != node.size or edge.size != node.size:
                raise ValueError('malformed CSR: ptr must have N + 1 entries and end at len(node) == len(edge)')
            if node.size and (int(node.min()) < 0 or int(node.max()) >= self.N):
                raise IndexError(f'neighbour id outside [0, {self.N}) in the successor / predecessor lists')
            if edge.size:
                if int(edge.min()) < 0:
                    raise IndexError('negative edge id')
                self.max_edge_id = max(self.max_edge_id, int(edge.max()))
        self._lib = _lib.load()

    @classmethod
    def from_dicts(cls, num_nodes, succs, edges, preds=None):
        """From the reference's pickled dicts ({idx}_succ.pkl, {idx}_edges.pkl, {idx}_pred.pkl)."""
        return cls(num_nodes, _csr_from_dicts(num_nodes, succs, edges),
                   None if preds is None else _csr_from_dicts(num_nodes, preds))

    @classmethod
    def from_edge_list(cls, src, dst, num_nodes):
        """succs / preds in edge-id order and ``edges[(u, v)]`` = the LAST edge id of that pair (a dict filled in edge
        order keeps the last write), without Python loops."""
        src, dst = _np(src, np.int64), _np(dst, np.int64)
        eid = np.arange(src.size, dtype=np.int64)
        _, inverse = np.unique(src * int(num_nodes) + dst, return_inverse=True)
        last = np.zeros(int(inverse.max()) + 1 if src.size else 0, dtype=np.int64)
        np.maximum.at(last, inverse, eid)
        pair_eid = last[inverse] if src.size else eid
        return cls(num_nodes, _csr_from_edges(src, dst, pair_eid, num_nodes), _csr_from_edges(dst, src, pair_eid, num_nodes))

    def _struct(self, csr):
        ptr, node, edge = csr
        return _lib.GnbWalkGraph(self.N, ptr.ctypes.data, node.ctypes.data, edge.ctypes.data)

    def _visited_bytes(self, visited):
        if visited is None:
            return None
        if isinstance(visited, (set, frozenset, list, tuple)):
            v = np.zeros(self.N, dtype=np.uint8)
            if len(visited):
                v[np.fromiter(visited, dtype=np.int64, count=len(visited))] = 1
            return v
        v = _np(visited, np.uint8)
        if v.shape != (self.N,):
            raise ValueError('visited must be a set of nodes or an (N,) mask')
        return v

    def run_greedy_both_ways(self, candidates, log_probs, visited=None, threads=0):
        """For every start edge ``(src, dst)`` in ``candidates``: ``(walk_f, walk_b, sumLogProb_f, sumLogProb_b)`` of
        ``run_greedy_both_ways(src, dst, logProbs, succs, preds, edges, visited)`` (inference.py:164-168); the sums are
        float32 scalars (numpy).  ``visited``: a set of nodes or an (N,) mask."""
        cand = _np(candidates, np.int32).reshape(-1, 2)
        n = cand.shape[0]
        src, dst = np.ascontiguousarray(cand[:, 0]), np.ascontiguousarray(cand[:, 1])
        logp = _np(log_probs, np.float32).reshape(-1)
        if logp.size <= self.max_edge_id:   # the reference raises IndexError on logProbs[edges[...]] (inference.py:95)
            raise IndexError(f'log_probs has {logp.size} entries but the graph refers to edge id {self.max_edge_id}')
        if n and (int(cand.min()) < 0 or int(cand.max()) >= self.N):
            raise IndexError(f'candidate endpoint outside [0, {self.N})')
        vis = self._visited_bytes(visited)
        off = np.zeros(n + 1, dtype=np.int64)
        back = np.zeros(max(n, 1), dtype=np.int64)
        sums = np.zeros((max(n, 1), 2), dtype=np.float32)
        g = self._struct(self._succ)
        args = lambda buf: (ctypes.byref(g), logp.ctypes.data, None if vis is None else vis.ctypes.data, n,  # noqa: E731
                            src.ctypes.data, dst.ctypes.data, int(threads), buf.ctypes.data, buf.size, off.ctypes.data,
                            back.ctypes.data, sums.ctypes.data)
        buf = np.empty(max(4096, 64 * n), dtype=np.int32)
        rc = self._lib.gnb_greedy_walks(*args(buf))
        if rc == -2:                                    # GNB_E_WORKSPACE: off[n] holds the size needed
            buf = np.empty(int(off[n]), dtype=np.int32)
            rc = self._lib.gnb_greedy_walks(*args(buf))
        _lib.check(rc, 'gnb_greedy_walks')
        out = []
        for k in range(n):
            w = buf[off[k]:off[k + 1]]
            out.append((w[back[k]:].tolist(), w[:back[k]].tolist(), sums[k, 0], sums[k, 1]))
        return out

    def get_contig_length(self, walk, prefix_length, read_length):
        """inference.py:30-37: sum of ``prefix_length`` over the walk's edges + ``read_length`` of its last node."""
        w = _np(walk, np.int32)
        pl, rl = _np(prefix_length, np.int64), _np(read_length, np.int64)
        if pl.size <= self.max_edge_id:
            raise IndexError(f'prefix_length has {pl.size} entries but the graph refers to edge id {self.max_edge_id}')
        if w.size and (int(w.min()) < 0 or int(w.max()) >= self.N or rl.size <= int(w.max())):
            raise IndexError('walk node outside the graph / read_length')
        out = ctypes.c_int64(0)
        g = self._struct(self._succ)
        _lib.check(self._lib.gnb_walk_contig_length(ctypes.byref(g), pl.ctypes.data, rl.ctypes.data, w.ctypes.data, w.size,
                                                    ctypes.byref(out)), 'gnb_walk_contig_length')
        return out.value

    def jumped_nodes(self, walk):
        """inference.py:316-322: the set ``trans`` of nodes an accepted walk jumps over (and their complements)."""
        if self._pred is None:
            raise ValueError('jumped_nodes needs the predecessor lists')
        w = _np(walk, np.int32)
        mark = np.zeros(self.N, dtype=np.uint8)
        gs, gp = self._struct(self._succ), self._struct(self._pred)
        _lib.check(self._lib.gnb_walk_jumped_nodes(ctypes.byref(gs), ctypes.byref(gp), w.ctypes.data, w.size,
                                                   mark.ctypes.data), 'gnb_walk_jumped_nodes')
        return set(np.nonzero(mark)[0].tolist())


def sample_edges(prob_edges, nb_paths, fast=False):
    """inference.py:54-67: ``nb_paths`` start edges drawn from the categorical distribution over the remaining edges'
    probabilities.  By default the same torch calls on the same tensors as the reference, so that a seeded run draws the
    same edges -- at the reference's cost: ``Categorical(probs.repeat(nb_paths, 1)).sample()`` takes one sample per row
    through torch's exponential-race path, i.e. nb_paths x E exponential variates per decoder iteration (5.5 s at 3 M
    edges and 50 paths).  ``fast=True`` draws the nb_paths samples from the one distribution (inverse-CDF, 25 ms there):
    the same distribution, different random numbers."""
    if prob_edges.shape[0] > 2 ** 24:           # torch.distributions.Categorical stops at 2**24 categories (:56-57)
        prob_edges = prob_edges[:2 ** 24]
    prob_edges = prob_edges.masked_fill(prob_edges < 1e-9, 1e-9)
    prob_edges = prob_edges / prob_edges.sum()
    if fast:
        return torch.multinomial(prob_edges, nb_paths, replacement=True)
    return torch.distributions.categorical.Categorical(prob_edges.repeat(nb_paths, 1)).sample()


def get_contigs_greedy(g, succs, preds, edges, len_threshold, nb_paths=50, use_labels=False, checkpoint_dir=None,
                       load_checkpoint=False, threads=0, fast_sampling=False):
    """``get_contigs_greedy`` (inference.py:167-361): repeat { drop the visited nodes, sample ``nb_paths`` start edges from
    what is left, walk greedily both ways from each, keep the candidate with the longest contig, mark it (and the nodes
    it jumps over) visited } until no edge is left or the best contig is shorter than ``len_threshold``.  Returns the
    list of walks.  ``g``: anything with ``edges()``, ``num_nodes()``, ``edata['score' | 'y', 'prefix_length']`` and
    ``ndata['read_length']`` (a DGLGraph, an ``AssemblyGraph``).  With the same torch seed the result is the reference's,
    walk for walk; its progress prints are not reproduced.  Checkpoints (every 10 contigs) use the reference's pickle.
    ``fast_sampling``: see ``sample_edges`` (gives up the identical random draws for a much cheaper iteration)."""
    import os
    import pickle
    n = g.num_nodes()
    src, dst = (_np(t, np.int64) for t in g.edges())
    wg = WalkGraph.from_dicts(n, succs, edges, preds)
    if use_labels:                                                      # :179-182
        prob_all = g.edata['y'].detach().cpu().float()
        log_probs = torch.log(prob_all.masked_fill(prob_all < 1e-9, 1e-9))
    else:                                                               # :183-184
        score = g.edata['score'].detach().cpu()
        prob_all = torch.sigmoid(score)
        log_probs = torch.log(torch.sigmoid(score))
    prob_all = prob_all.reshape(-1)
    prefix_length, read_length = g.edata['prefix_length'], g.ndata['read_length']
    all_contigs, all_walks_len, all_contigs_len = [], [], []
    visited = np.zeros(n, dtype=bool)
    ckpt_file = None if checkpoint_dir is None else os.path.join(checkpoint_dir, 'checkpoint.pkl')
    if load_checkpoint and ckpt_file is not None and os.path.isfile(ckpt_file):        # :189-196
        with open(ckpt_file, 'rb') as f:
            ck = pickle.load(f)
        all_contigs, all_walks_len, all_contigs_len = ck['walks'], ck['all_walks_len'], ck['all_contigs_len']
        if ck['visited']:
            visited[np.fromiter(ck['visited'], dtype=np.int64, count=len(ck['visited']))] = True
    while True:
        # get_subgraph (:40-51): the edges whose endpoints are both unvisited, in the order of their ids
        remaining = np.nonzero(~visited[src] & ~visited[dst])[0]
        if remaining.size == 0:                                         # :201-203
            break
        idx_edges = sample_edges(prob_all[torch.from_numpy(remaining)], nb_paths, fast_sampling)   # :205-210
        cands = list(dict.fromkeys((int(src[remaining[i]]), int(dst[remaining[i]])) for i in idx_edges.tolist()))
        walks, contig_lens = [], []
        for (s, d), (walk_f, walk_b, _, _) in zip(cands, wg.run_greedy_both_ways(cands, log_probs, visited, threads)):
            walk = walk_b + walk_f                                      # :257
            length = wg.get_contig_length(walk, prefix_length, read_length)             # :261
            if s == d or len(walk) < 2:                                 # :263-264, :282-287: a self-loop counts for nothing
                length = 0
            walks.append(walk)
            contig_lens.append(length)
        best = contig_lens.index(max(contig_lens))                      # :306-307: first of the longest
        best_walk = walks[best]
        if contig_lens[best] < len_threshold:                           # :335-336
            break
        visited[best_walk] = True                                       # visited_f | visited_b: the walk and its complements
        visited[np.asarray(best_walk) ^ 1] = True
        jumped = wg.jumped_nodes(best_walk)                             # :316-322
        if jumped:
            visited[np.fromiter(jumped, dtype=np.int64, count=len(jumped))] = True
        all_contigs.append(best_walk)
        all_walks_len.append(len(best_walk))
        all_contigs_len.append(contig_lens[best])
        if len(all_contigs) % 10 == 0 and checkpoint_dir is not None:   # :344-359
            ck = dict(walks=all_contigs, visited=set(np.nonzero(visited)[0].tolist()), all_walks_len=all_walks_len,
                      all_contigs_len=all_contigs_len)
            tmp = os.path.join(checkpoint_dir, 'checkpoint_tmp.pkl')
            with open(tmp, 'wb') as f:
                pickle.dump(ck, f)
            os.rename(tmp, ckpt_file)
    return all_contigs
