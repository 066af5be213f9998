"""Graph hand-off and the callers' code either side of ``model(g, x, e)``.

The reference keeps an assembly graph as a DGLGraph written with ``dgl.save_graphs``
(create_inference_graphs.py:27, graph_dataset.py:46-56); its callers then build the model inputs with torch on the host
(utils/data_utils.py:31-51, inference.py:413-420, train.py:112-120), call the model, and either store the scores
(inference.py:438-442) or form the loss (train.py:103-109, 138-145, 158-170).  This module is that slice without DGL:

* ``AssemblyGraph`` -- edge list + ``edata`` / ``ndata`` dicts, saved as a plain ``torch.save`` dict; ``from_dgl`` converts
  a DGLGraph where DGL exists (INTEGRATION.md has the one-off exporter), ``reversed()`` is ``dgl.reverse(g, True, True)``;
* ``preprocess_graph`` / ``add_positional_encoding`` / ``get_full_ne_features`` -- same names and results as the
  reference's, but the degrees come from the staged ``GraphIndex`` and the z-scores from ``gnb_zscore_cols`` on the GPU;
* ``compute_scores`` -- the ``get scores`` block of ``inference()``; writes ``{idx}_predicts.pt`` like the reference;
* ``symmetry_loss`` / ``get_bce_loss_full`` / ``get_symmetry_loss_full`` -- the full-graph losses of train.py;
* ``node_subgraph`` / ``mask_graph_strandwise`` and the ``*_partition`` feature / loss functions -- the reference's
  strand-wise masking and mini-batch code (train.py:91-100, 125-135, 148-155, 173-185) on device-built induced subgraphs
  (section 8(f) row 3; METIS itself is not built: any node set works).

Everything that touches N- or E-sized data runs on the CUDA device (there is no CPU fallback); the losses are the caller's
own torch code in the reference and stay torch here (4 bytes per edge)."""
import os

import torch
import torch.nn.functional as F

from . import ops
from .graph import GraphIndex, _cuda_device

FORMAT = 'gnnome_b200.assembly_graph'
FORMAT_VERSION = 1


class AssemblyGraph:
    """What the model's ``graph`` argument needs (``edges()``, ``num_nodes()``) plus the DGL-style feature dicts."""

    def __init__(self, src, dst, num_nodes, edata=None, ndata=None):
        self._src = torch.as_tensor(src).to(torch.int32).contiguous()
        self._dst = torch.as_tensor(dst).to(torch.int32).contiguous()
        if self._src.ndim != 1 or self._src.shape != self._dst.shape:
            raise ValueError('src/dst must be 1-D and of equal length')
        self._n = int(num_nodes)
        self.edata = dict(edata or {})
        self.ndata = dict(ndata or {})
        for k, v in self.edata.items():
            if v.shape[0] != self.num_edges():
                raise ValueError(f'edata[{k!r}] has {v.shape[0]} rows for {self.num_edges()} edges')
        for k, v in self.ndata.items():
            if v.shape[0] != self._n:
                raise ValueError(f'ndata[{k!r}] has {v.shape[0]} rows for {self._n} nodes')

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())

    number_of_nodes, number_of_edges = num_nodes, num_edges

    def edges(self):
        return self._src, self._dst

    def reversed(self):
        """``dgl.reverse(g, copy_ndata=True, copy_edata=True)`` (train.py:165): edge k becomes dst_k -> src_k, same edge
        ids, both feature dicts carried over -- so ``in_deg`` / ``out_deg`` are still those of the original graph."""
        return AssemblyGraph(self._dst, self._src, self._n, self.edata, self.ndata)

    @classmethod
    def from_dgl(cls, g):
        """From anything DGLGraph-shaped (``edges()``, ``num_nodes()``, ``edata``, ``ndata``)."""
        src, dst = g.edges()
        return cls(src.cpu(), dst.cpu(), g.num_nodes(), {k: v for k, v in g.edata.items()},
                   {k: v for k, v in g.ndata.items()})

    def save(self, path):
        cpu = lambda d: {k: v.detach().cpu() for k, v in d.items()}  # noqa: E731
        torch.save(dict(format=FORMAT, version=FORMAT_VERSION, src=self._src.cpu(), dst=self._dst.cpu(),
                        num_nodes=self._n, edata=cpu(self.edata), ndata=cpu(self.ndata)), path)

    @classmethod
    def load(cls, path):
        rec = torch.load(path, map_location='cpu', weights_only=True)
        if not isinstance(rec, dict) or rec.get('format') != FORMAT:
            raise ValueError(f'{path} is not a {FORMAT} file')
        if rec['version'] > FORMAT_VERSION:
            raise ValueError(f'{path}: format version {rec["version"]} is newer than this reader ({FORMAT_VERSION})')
        return cls(rec['src'], rec['dst'], rec['num_nodes'], rec['edata'], rec['ndata'])


def preprocess_graph(g, use_similarities=True, device=None):
    """utils/data_utils.py:31-41: ``g.edata['e'] = [z(overlap_length), overlap_similarity]`` ((E, 1) without the
    similarities), fp32, on the device.  The placeholder ``ndata['x'] = ones`` of :33 is never read and not made."""
    device = _cuda_device(device)
    E = g.num_edges()
    e = torch.empty((E, 2 if use_similarities else 1), dtype=torch.float32, device=device)
    e[:, 0] = g.edata['overlap_length'].to(device=device, dtype=torch.float32)
    if use_similarities:
        e[:, 1] = g.edata['overlap_similarity'].to(device=device, dtype=torch.float32)
    g.edata['e'] = ops.zscore_cols(e, 0b01)
    return g


def add_positional_encoding(g, device=None):
    """utils/data_utils.py:50-51 (``nb_pos_enc`` is 0 in configs/hyperparameters.py:26, so nothing else runs):
    ``ndata['in_deg']`` / ``['out_deg']`` as floats, read off the two CSR pointer arrays of the staged graph."""
    gi = GraphIndex.from_graph(g, device)
    deg = ops.degree_rows(gi)
    g.ndata['in_deg'], g.ndata['out_deg'] = deg[:, 0], deg[:, 1]
    return g


def get_full_ne_features(g, reverse=False, device=None):
    """train.py:112-122 and inference.py:413-420: ``x = [z(in_deg), z(out_deg)]`` (columns swapped for the reversed
    graph, whose ``ndata`` is the original's) and ``e = g.edata['e']``; both on the device."""
    device = _cuda_device(device)
    if 'in_deg' not in g.ndata or 'out_deg' not in g.ndata:
        add_positional_encoding(g, device)
    if 'e' not in g.edata:
        preprocess_graph(g, device=device)
    cols = ('out_deg', 'in_deg') if reverse else ('in_deg', 'out_deg')
    x = torch.stack([g.ndata[c].to(device=device, dtype=torch.float32) for c in cols], dim=1).contiguous()
    x = ops.zscore_cols(x, 0b11)
    return x, g.edata['e'].to(device)


def compute_scores(model, g, idx=0, inference_dir=None, device=None):
    """The ``get scores`` block of ``inference()`` (inference.py:408-442): an existing ``{idx}_predicts.pt`` wins
    (:429-431); otherwise ``model(g, x, e).squeeze()`` is stored in ``g.edata['score']`` and saved there (:438-442).
    ``model`` is a loaded ``gnnome_b200.models`` model in ``eval()`` mode.  Returns the (E,) fp32 CPU tensor the
    reference would have produced with its hard-coded ``device = 'cpu'`` (:388)."""
    path = None if inference_dir is None else os.path.join(inference_dir, f'{idx}_predicts.pt')
    if path is not None and os.path.isfile(path):
        g.edata['score'] = torch.load(path, map_location='cpu', weights_only=True)
        return g.edata['score']
    if model.training:
        raise RuntimeError('compute_scores needs model.eval() (inference.py:436)')
    with torch.no_grad():
        x, e = get_full_ne_features(g, reverse=False, device=device)
        scores = model(g, x, e).squeeze().cpu()
    g.edata['score'] = scores
    if path is not None:
        os.makedirs(inference_dir, exist_ok=True)
        torch.save(scores, path)
    return scores


def node_subgraph(g, keep, device=None):
    """``dgl.node_subgraph(g, keep, store_ids=True)`` (train.py:95): the subgraph induced by the nodes with ``keep`` set,
    renumbered in increasing id order, edges in the order of their ids; every ``ndata`` / ``edata`` entry is sliced and
    ``ndata['_ID']`` / ``edata['_ID']`` map back to ``g``.  All of it on the device (``gnb_subgraph_*``)."""
    gi = GraphIndex.from_graph(g, device)
    node_id, edge_id, sub_src, sub_dst = ops.node_subgraph(keep, gi.src, gi.dst, gi.N)
    nid, eid = node_id.long(), edge_id.long()
    sub = AssemblyGraph(sub_src, sub_dst, int(node_id.numel()),
                        {k: v.to(gi.device)[eid] for k, v in g.edata.items() if k != '_ID'},
                        {k: v.to(gi.device)[nid] for k, v in g.ndata.items() if k != '_ID'})
    sub.ndata['_ID'], sub.edata['_ID'] = node_id, edge_id
    return sub


def mask_graph_strandwise(g, fraction, device=None, generator=None):
    """train.py:91-100: keep each strand pair (nodes 2k, 2k+1) with probability ``fraction`` and take the induced
    subgraph.  The random draw is made on the device, as in the reference when it trains on a GPU."""
    device = _cuda_device(device)
    keep_half = torch.rand(g.num_nodes() // 2, device=device, generator=generator) < fraction
    keep = torch.zeros(g.num_nodes(), dtype=torch.bool, device=device)   # an unpaired last node is dropped
    keep[0:2 * keep_half.numel():2] = keep_half
    keep[1:2 * keep_half.numel():2] = keep_half
    return node_subgraph(g, keep, device)


def get_partition_ne_features(sub_g, g, reverse=False, device=None):
    """train.py:125-135: features of a mini-batch ``sub_g`` of ``g``: the parent's degrees at ``sub_g.ndata['_ID']``,
    z-scored within the batch, and the parent's ``e`` rows at ``sub_g.edata['_ID']``."""
    device = _cuda_device(device)
    if 'in_deg' not in g.ndata or 'out_deg' not in g.ndata:
        add_positional_encoding(g, device)
    if 'e' not in g.edata:
        preprocess_graph(g, device=device)
    nid, eid = sub_g.ndata['_ID'].to(device).long(), sub_g.edata['_ID'].to(device).long()
    cols = ('out_deg', 'in_deg') if reverse else ('in_deg', 'out_deg')
    x = torch.stack([g.ndata[c].to(device=device, dtype=torch.float32)[nid] for c in cols], dim=1).contiguous()
    return ops.zscore_cols(x, 0b11), g.edata['e'].to(device)[eid]


def get_bce_loss_partition(sub_g, g, model, pos_weight, device=None):
    """train.py:148-155 -> (loss, logits) of one mini-batch."""
    x, e = get_partition_ne_features(sub_g, g, reverse=False, device=device)
    logits = model(sub_g, x, e).squeeze(-1)
    labels = g.edata['y'].to(logits.device)[sub_g.edata['_ID'].to(logits.device).long()].to(logits.dtype)
    return F.binary_cross_entropy_with_logits(logits, labels, pos_weight=_pos_weight(pos_weight, logits)), logits


def get_symmetry_loss_partition(sub_g, g, model, pos_weight, alpha, device=None):
    """train.py:173-185: the symmetry loss of one mini-batch (second forward over the reversed batch)."""
    x, e = get_partition_ne_features(sub_g, g, reverse=False, device=device)
    logits_org = model(sub_g, x, e).squeeze(-1)
    labels = g.edata['y'].to(logits_org.device)[sub_g.edata['_ID'].to(logits_org.device).long()].to(logits_org.dtype)
    sub_rev = sub_g.reversed()
    x, e = get_partition_ne_features(sub_rev, g, reverse=True, device=device)
    logits_rev = model(sub_rev, x, e).squeeze(-1)
    return symmetry_loss(logits_org, logits_rev, labels, pos_weight, alpha=alpha), logits_org


def _pos_weight(pos_weight, like):
    return torch.as_tensor(pos_weight, dtype=like.dtype, device=like.device)


def symmetry_loss(org_scores, rev_scores, labels, pos_weight=1.0, alpha=1.0):
    """train.py:103-109: mean over edges of BCE(org) + BCE(rev) + alpha * |org - rev|."""
    w = _pos_weight(pos_weight, org_scores)
    per_edge = sum(F.binary_cross_entropy_with_logits(s, labels, pos_weight=w, reduction='none')
                   for s in (org_scores, rev_scores))
    return (per_edge + alpha * (org_scores - rev_scores).abs()).mean()


def get_bce_loss_full(g, model, pos_weight, device=None):
    """train.py:138-145 -> (loss, logits (E,))."""
    x, e = get_full_ne_features(g, reverse=False, device=device)
    logits = model(g, x, e).squeeze(-1)
    labels = g.edata['y'].to(device=logits.device, dtype=logits.dtype)
    return F.binary_cross_entropy_with_logits(logits, labels, pos_weight=_pos_weight(pos_weight, logits)), logits


def get_symmetry_loss_full(g, model, pos_weight, alpha, device=None):
    """train.py:158-170: a second forward over the reversed graph with the degree columns swapped -> (loss, logits_org)."""
    x, e = get_full_ne_features(g, reverse=False, device=device)
    logits_org = model(g, x, e).squeeze(-1)
    labels = g.edata['y'].to(device=logits_org.device, dtype=logits_org.dtype)
    g_rev = getattr(g, '_gnb_reversed', None)
    if g_rev is None:  # staged once: the reversed index is reused by every later step on this graph
        g_rev = g.reversed()
        try:
            g._gnb_reversed = g_rev
        except AttributeError:
            pass
    else:
        g_rev.edata, g_rev.ndata = dict(g.edata), dict(g.ndata)
    x, e = get_full_ne_features(g_rev, reverse=True, device=device)
    logits_rev = model(g_rev, x, e).squeeze(-1)
    return symmetry_loss(logits_org, logits_rev, labels, pos_weight, alpha=alpha), logits_org


class AssemblyGraphDataset:
    """graph_dataset.py:46-56 without DGL: every ``{root}/{assembler}/processed/{idx}.pt`` (``AssemblyGraph.save``
    files) is loaded, preprocessed and given its degrees; iteration yields ``(idx, graph)`` in index order."""

    def __init__(self, root, assembler, device=None, preprocess=True):
        """``preprocess=False`` only loads the graphs (no device work): enough when ``{idx}_predicts.pt`` already exist."""
        self.root = os.path.abspath(root)
        self.assembler = assembler
        self.assembly_dir = os.path.join(self.root, assembler)
        self.save_dir = os.path.join(self.assembly_dir, 'processed')
        self.info_dir = os.path.join(self.assembly_dir, 'info')
        self.graph_list = []
        for name in os.listdir(self.save_dir):
            stem, ext = os.path.splitext(name)
            if ext != '.pt' or not stem.isdigit():
                continue
            g = AssemblyGraph.load(os.path.join(self.save_dir, name))
            if preprocess:
                g = add_positional_encoding(preprocess_graph(g, device=device), device=device)
            self.graph_list.append((int(stem), g))
        self.graph_list.sort(key=lambda t: t[0])

    def __len__(self):
        return len(self.graph_list)

    def __getitem__(self, i):
        return self.graph_list[i]
