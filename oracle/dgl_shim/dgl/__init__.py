"""Minimal CPU stand-in for the part of ``dgl==0.8.1`` used by GNNome's model code.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Semantics restated from the DGL 0.8 docs:

* ``apply_edges(fn.u_add_v(a, b, out))``      out[k] = ndata[a][src_k] + ndata[b][dst_k]
* ``update_all(fn.u_mul_e(a, w, m), fn.sum(m, o))``  o[i] = sum_{k: dst_k = i} ndata[a][src_k] * edata[w][k]
* ``update_all(fn.copy_e(w, m), fn.sum(m, o))``      o[i] = sum_{k: dst_k = i} edata[w][k]
  (nodes without in-edges receive zeros)
* ``reverse(g, copy_ndata, copy_edata)``      edge k becomes (dst_k -> src_k); edge ids preserved
* ``add_reverse_edges(g, copy_edata=True)``   edges = [E originals..., E reversed...]
* ``apply_edges(udf)``                        udf(EdgeBatch) with .src/.dst/.data views over all edges
"""
import contextlib

import torch

from . import function  # noqa: F401
from . import nn  # noqa: F401


class _EdgeBatch:
    def __init__(self, g):
        s, d = g._src.long(), g._dst.long()
        self.src = {k: v[s] for k, v in g.ndata.items()}
        self.dst = {k: v[d] for k, v in g.ndata.items()}
        self.data = g.edata


class DGLGraph:
    def __init__(self, src, dst, num_nodes):
        self._src = torch.as_tensor(src)
        self._dst = torch.as_tensor(dst)
        self._n = int(num_nodes)
        self.ndata = {}
        self.edata = {}

    # -- structure ---------------------------------------------------------
    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def edges(self, form='uv'):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst.long(), minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src.long(), minlength=self._n)

    @property
    def device(self):
        return self._src.device

    def to(self, device):
        g = DGLGraph(self._src.to(device), self._dst.to(device), self._n)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        return g

    def int(self):
        g = DGLGraph(self._src.int(), self._dst.int(), self._n)
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        return g

    def long(self):
        g = DGLGraph(self._src.long(), self._dst.long(), self._n)
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        return g

    # -- frames ------------------------------------------------------------
    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = self.ndata, self.edata
        self.ndata, self.edata = dict(nd), dict(ed)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    # -- message passing -----------------------------------------------------
    def apply_edges(self, func):
        if isinstance(func, function._Binary):
            s, d = self._src.long(), self._dst.long()
            self.edata[func.out] = func.op(self.ndata[func.lhs][s], self.ndata[func.rhs][d])
        else:
            self.edata.update(func(_EdgeBatch(self)))

    def update_all(self, message_func, reduce_func):
        s, d = self._src.long(), self._dst.long()
        if isinstance(message_func, function._UMulE):
            m = self.ndata[message_func.u][s] * self.edata[message_func.e]
        elif isinstance(message_func, function._CopyE):
            m = self.edata[message_func.e]
        else:
            raise NotImplementedError(type(message_func))
        assert isinstance(reduce_func, function._Sum) and reduce_func.msg == message_func.out
        out = torch.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        out.index_add_(0, d, m)
        self.ndata[reduce_func.out] = out


def graph(data, num_nodes=None, idtype=None, device=None):
    src, dst = data
    src, dst = torch.as_tensor(src), torch.as_tensor(dst)
    if num_nodes is None:
        num_nodes = int(max(src.max(), dst.max())) + 1 if src.numel() else 0
    if idtype is not None:
        src, dst = src.to(idtype), dst.to(idtype)
    g = DGLGraph(src, dst, num_nodes)
    return g.to(device) if device is not None else g


def reverse(g, copy_ndata=True, copy_edata=False):
    r = DGLGraph(g._dst, g._src, g._n)
    if copy_ndata:
        r.ndata = dict(g.ndata)
    if copy_edata:
        r.edata = dict(g.edata)
    return r


def add_reverse_edges(g, copy_ndata=True, copy_edata=False):
    r = DGLGraph(torch.cat((g._src, g._dst)), torch.cat((g._dst, g._src)), g._n)
    if copy_ndata:
        r.ndata = dict(g.ndata)
    if copy_edata:
        r.edata = {k: torch.cat((v, v), dim=0) for k, v in g.edata.items()}
    return r


def add_self_loop(g):
    loop = torch.arange(g._n, dtype=g._src.dtype, device=g._src.device)
    r = DGLGraph(torch.cat((g._src, loop)), torch.cat((g._dst, loop)), g._n)
    r.ndata = dict(g.ndata)
    return r


def seed(_):  # dgl.seed, used by the reference's utils.set_seed
    return None


NID, EID = '_ID', '_ID'   # dgl.NID / dgl.EID: the feature names store_ids writes


def node_subgraph(g, nodes, relabel_nodes=True, store_ids=True):
    """DGL 0.8 ``dgl.node_subgraph``: the subgraph induced by ``nodes`` (a bool mask over the nodes or a tensor of ids).
    Nodes are relabelled in the order given (increasing ids for a mask), an edge is kept when both endpoints are, kept
    edges stay in the order of their ids (DGL doc example: edges 0..4 of a 5-cycle, nodes [0, 1, 4] -> EID [0, 4]);
    every node / edge feature is sliced, and ``store_ids`` adds the original ids as ``ndata['_ID']`` / ``edata['_ID']``."""
    assert relabel_nodes, 'only the relabelling form is used by the reference'
    nodes = torch.as_tensor(nodes)
    ids = torch.nonzero(nodes, as_tuple=False).flatten() if nodes.dtype == torch.bool else nodes.long()
    new_id = torch.full((g._n,), -1, dtype=torch.long)
    new_id[ids] = torch.arange(ids.numel())
    s, d = g._src.long(), g._dst.long()
    eids = torch.nonzero((new_id[s] >= 0) & (new_id[d] >= 0), as_tuple=False).flatten()
    sub = DGLGraph(new_id[s[eids]].to(g._src.dtype), new_id[d[eids]].to(g._dst.dtype), ids.numel())
    sub.ndata = {k: v[ids] for k, v in g.ndata.items()}
    sub.edata = {k: v[eids] for k, v in g.edata.items()}
    if store_ids:
        sub.ndata[NID], sub.edata[EID] = ids.to(g._src.dtype), eids.to(g._src.dtype)
    return sub


class _EdgeSpace:
    """``g.edges[u, v]``: the edges between the node pairs; ``.data[name]`` reads their features (DGL picks ONE edge id
    per pair through ``edge_ids``; which one, when a pair has several, is unspecified -- callers here use simple graphs)."""

    def __init__(self, g, eids):
        self.data = {k: v[eids] for k, v in g.edata.items()}


class _EdgesView:
    """``g.edges()`` -> (src, dst) and ``g.edges[u, v]`` -> _EdgeSpace, like DGL's EdgeView."""

    def __init__(self, g):
        self._g = g

    def __call__(self, form='uv'):
        return self._g._src, self._g._dst

    def __getitem__(self, key):
        u, v = key
        g = self._g
        if getattr(g, '_pair_eid', None) is None:
            g._pair_eid = {}
            for k, (a, b) in enumerate(zip(g._src.tolist(), g._dst.tolist())):
                g._pair_eid.setdefault((a, b), k)
        u = u.tolist() if torch.is_tensor(u) else list(u)
        v = v.tolist() if torch.is_tensor(v) else list(v)
        return _EdgeSpace(g, torch.tensor([g._pair_eid[(a, b)] for a, b in zip(u, v)], dtype=torch.long))


DGLGraph.edges = property(lambda self: _EdgesView(self))
