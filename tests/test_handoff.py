"""CPU side of the graph hand-off: the oracle's restatement of the callers' feature
and loss code against fixtures produced by the reference's OWN functions (oracle/make_golden_handoff.py), and the
``AssemblyGraph`` container (no device work)."""
import os
import sys

import pytest
import torch

from oracle import restatement as R


def _raw(rec):
    r = rec['raw']
    return r['src'], r['dst'], r['num_nodes'], r['overlap_length'], r['overlap_similarity'], r['y']


def test_oracle_features_match_reference_functions(golden):
    g = golden('handoff_scores')
    src, dst, n, ol_len, ol_sim, _ = _raw(g)
    assert ol_len.dtype == torch.int64 and g['e'].shape == (src.numel(), 2)
    assert torch.equal(R.edge_input_features(ol_len, ol_sim), g['e'])
    assert torch.equal(R.node_input_features(src, dst, n), g['x'])
    assert torch.equal(R.node_input_features(src, dst, n, reverse=True), g['x_rev'])
    assert torch.equal(g['x_rev'], g['x'].flip(1))          # reversed graph: same columns, swapped (train.py:117-118)
    assert R.edge_input_features(ol_len, ol_sim, use_similarities=False).shape == (src.numel(), 1)


def test_oracle_scores_from_raw_graph(golden, shipped_weights):
    g = golden('handoff_scores')
    src, dst, n, ol_len, ol_sim, _ = _raw(g)
    with torch.no_grad():
        logits = R.model_forward(shipped_weights, src, dst, n, R.node_input_features(src, dst, n),
                                 R.edge_input_features(ol_len, ol_sim)).squeeze()
    assert logits.shape == g['predicts'].shape == (src.numel(),)
    assert (torch.sigmoid(logits.double()) - torch.sigmoid(g['predicts'].double())).abs().max().item() <= 1e-5


@pytest.mark.parametrize('kind', ['bce', 'sym'])
def test_oracle_losses_match_reference_functions(golden, shipped_weights, kind):
    g = golden('handoff_losses')
    src, dst, n, ol_len, ol_sim, y = _raw(g)
    p = {k: v.clone() for k, v in shipped_weights.items()}
    e = R.edge_input_features(ol_len, ol_sim)
    pw = torch.tensor([g['pos_weight']])
    fwd = lambda s, d, rev: R.model_forward(p, s, d, n, R.node_input_features(src, dst, n, reverse=rev), e,  # noqa: E731
                                            training=True, cast=False).squeeze(-1)
    with torch.no_grad():
        org = fwd(src, dst, False)
        if kind == 'bce':
            loss = torch.nn.functional.binary_cross_entropy_with_logits(org, y, pos_weight=pw)
        else:
            loss = R.symmetry_loss(org, fwd(dst, src, True), y, pw, g['alpha'])   # dgl.reverse: endpoints swapped
    assert abs(loss.item() - g[kind]['loss'].item()) <= 2e-6 * max(1.0, abs(loss.item()))
    # bn_e sees two batches per layer and forward (gated_gcn_full.py:106,119); the symmetric loss runs two forwards
    nbt = int(g[kind]['buffers']['gnn.convs.0.bn_e.num_batches_tracked'] - shipped_weights['gnn.convs.0.bn_e.num_batches_tracked'])
    assert nbt == (2 if kind == 'bce' else 4)


def test_assembly_graph_container_round_trip(tmp_path, golden):
    from gnnome_b200.assembly import AssemblyGraph, FORMAT
    g = golden('handoff_scores')
    src, dst, n, ol_len, ol_sim, y = _raw(g)
    ag = AssemblyGraph(src, dst, n, dict(overlap_length=ol_len, overlap_similarity=ol_sim, y=y))
    assert ag.num_nodes() == n and ag.num_edges() == src.numel() and ag.edges()[0].dtype == torch.int32
    path = tmp_path / '0.pt'
    ag.save(path)
    back = AssemblyGraph.load(path)
    assert torch.equal(back.edges()[0], ag.edges()[0]) and torch.equal(back.edges()[1], ag.edges()[1])
    assert back.num_nodes() == n and set(back.edata) == {'overlap_length', 'overlap_similarity', 'y'}
    assert torch.equal(back.edata['overlap_length'], ol_len) and back.edata['overlap_length'].dtype == torch.int64
    rev = ag.reversed()                                           # dgl.reverse(g, True, True)
    assert torch.equal(rev.edges()[0], ag.edges()[1]) and torch.equal(rev.edges()[1], ag.edges()[0])
    assert rev.edata['y'] is ag.edata['y']
    torch.save(dict(format='something else'), path)
    with pytest.raises(ValueError, match=FORMAT):
        AssemblyGraph.load(path)
    with pytest.raises(ValueError, match='rows'):
        AssemblyGraph(src, dst, n, dict(y=y[:-1]))


def test_assembly_graph_from_dgl_shaped_graph(golden):
    """``from_dgl`` on the oracle's DGL stand-in (the exporter of INTEGRATION.md runs it on a real DGLGraph)."""
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'dgl_shim')
    sys.path.insert(0, shim)
    try:
        import dgl
    finally:
        sys.path.remove(shim)
    from gnnome_b200.assembly import AssemblyGraph
    g = golden('handoff_scores')
    src, dst, n, ol_len, ol_sim, _ = _raw(g)
    dg = dgl.graph((src, dst), num_nodes=n)
    dg.edata['overlap_length'], dg.edata['overlap_similarity'] = ol_len, ol_sim
    dg.ndata['in_deg'] = dg.in_degrees().float()
    ag = AssemblyGraph.from_dgl(dg)
    assert ag.num_edges() == src.numel() and torch.equal(ag.edges()[1], dst.to(torch.int32))
    assert torch.equal(ag.ndata['in_deg'], dg.ndata['in_deg'])


def test_feature_code_needs_the_gpu():
    from gnnome_b200.assembly import AssemblyGraph, preprocess_graph
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    ag = AssemblyGraph([0], [0], 1, dict(overlap_length=torch.tensor([5]), overlap_similarity=torch.tensor([1.0])))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        preprocess_graph(ag)


def test_oracle_subgraphs_match_reference_masking(golden):
    """Strand-wise masking (train.py:91-100) and a mini-batch (train.py:125-135) from the reference's own functions."""
    g = golden('handoff_subgraphs')
    src, dst, n, ol_len, ol_sim, y = _raw(g)
    keep = g['mask_keep']
    assert torch.equal(keep[0::2], keep[1::2])                     # both strands of a read go together
    node_id, edge_id, s, d = R.node_subgraph_ids(src, dst, n, keep)
    assert torch.equal(node_id, g['mask_node_id'].long()) and torch.equal(edge_id, g['mask_edge_id'].long())
    assert torch.equal(s, g['mask_src'].long()) and torch.equal(d, g['mask_dst'].long())
    assert torch.equal(R.edge_input_features(ol_len, ol_sim)[edge_id], g['mask_e'])      # features are sliced,
    assert torch.equal(torch.bincount(dst.long(), minlength=n).float()[node_id], g['mask_in_deg'])  # not recomputed
    node_id, edge_id, _, _ = R.node_subgraph_ids(src, dst, n, g['batch_keep'])
    assert torch.equal(node_id, g['batch_node_id'].long()) and torch.equal(edge_id, g['batch_edge_id'].long())
    ind, outd = torch.bincount(dst.long(), minlength=n).float(), torch.bincount(src.long(), minlength=n).float()
    assert torch.equal(R.partition_node_features(ind, outd, node_id), g['batch_x'])
    assert torch.equal(R.edge_input_features(ol_len, ol_sim)[edge_id], g['batch_e'])
