"""The oracle restatement against the fixtures produced by the reference's own code
(oracle/make_golden.py) and, where /root/reference exists, against the reference run live."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import reference_runner as rr
from oracle import restatement as R
from gnnome_b200 import synth


def _fwd(sd, g, **kw):
    with torch.no_grad():
        return R.model_forward(sd, g['src'], g['dst'], g['num_nodes'], g['x'], g['e'], **kw)


@pytest.mark.parametrize('name', ['sym_shipped_tiny', 'sym_shipped_2k'])
def test_restatement_matches_reference_golden_bitwise(golden, shipped_weights, name):
    g = golden(name)
    assert torch.equal(_fwd(shipped_weights, g, faithful=True), g['logits'])
    # the single-sigma form (what the CUDA path implements) is the same function in eval mode
    assert torch.equal(_fwd(shipped_weights, g, faithful=False), g['logits'])


def test_restatement_per_layer_states(golden, shipped_weights):
    g = golden('sym_shipped_tiny')
    _, layers = _fwd(shipped_weights, g, return_layers=True)
    assert len(layers) == 8
    for (h, e), (hr, er) in zip(layers, g['layers']):
        assert torch.equal(h, hr) and torch.equal(e, er)


@pytest.mark.parametrize('name,kw', [
    ('gated_directed', dict(model='gated', directed=True)),
    ('gated_undirected', dict(model='gated', directed=False)),
    ('sym_layernorm', dict(model='sym', normalization='layer')),
    ('sym_h128', dict(model='sym')),
])
def test_restatement_other_models(golden, name, kw):
    g = golden(name)
    out = _fwd(g['state_dict'], g, **kw)
    assert out.shape == g['logits'].shape
    torch.testing.assert_close(out, g['logits'], rtol=0, atol=2e-6)


def test_restatement_train_step(golden, shipped_weights):
    g = golden('sym_shipped_trainstep')
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
         for k, v in shipped_weights.items()}
    logits = R.model_forward(p, g['src'], g['dst'], g['num_nodes'], g['x'], g['e'],
                             training=True, cast=False).squeeze(-1)
    loss = F.binary_cross_entropy_with_logits(logits, g['y'], pos_weight=torch.tensor(g['pos_weight']))
    loss.backward()
    assert torch.equal(logits.detach(), g['logits'])
    assert torch.equal(loss.detach(), g['loss'])
    for k, gr in g['grads'].items():
        torch.testing.assert_close(p[k].grad, gr, rtol=1e-5, atol=1e-7, msg=k)
    for k, b in g['buffers'].items():  # incl. the double bn_e update: num_batches_tracked += 2
        torch.testing.assert_close(p[k].detach(), b, rtol=0, atol=0, msg=k)
    assert int(g['buffers']['gnn.convs.0.bn_e.num_batches_tracked']
               - shipped_weights['gnn.convs.0.bn_e.num_batches_tracked']) == 2
    assert int(g['buffers']['gnn.convs.0.bn_h.num_batches_tracked']
               - shipped_weights['gnn.convs.0.bn_h.num_batches_tracked']) == 1


def test_fp64_bounds_fp32_error(golden, shipped_weights):
    g = golden('sym_shipped_2k')
    truth = _fwd(shipped_weights, g, dtype=torch.float64)
    err = (torch.sigmoid(truth) - torch.sigmoid(g['logits'].double())).abs().max()
    assert err < 5e-5


@pytest.mark.skipif(not rr.available(), reason='reference tree not mounted')
def test_restatement_matches_live_reference(shipped_weights):
    dgl, layers, models = rr.load()
    src, dst = synth.make_assembly_graph(3000, 18000, seed=21)
    x, e = synth.make_features(src, dst, 3000, seed=21)
    src, dst, x, e = map(torch.from_numpy, (src, dst, x, e))
    m = models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch')
    assert len(m.state_dict()) == 190
    m.load_state_dict(shipped_weights, strict=True)
    m.eval()
    with rr.quiet(), torch.no_grad():
        ref = m(dgl.graph((src, dst), num_nodes=3000), x, e)
    out = R.model_forward(shipped_weights, src, dst, 3000, x, e)
    assert torch.equal(out, ref)


def test_init_state_dict_keys_match_shipped(shipped_weights):
    sd = R.init_state_dict(model='sym', hidden=64)
    assert set(sd) == set(shipped_weights)
    for k in sd:
        assert sd[k].shape == shipped_weights[k].shape, k


def test_synth_graph_shape():
    n, m = 20000, 120000
    src, dst = synth.make_assembly_graph(n, m, seed=0)
    assert src.dtype == np.int32 and src.shape == (m,) and dst.shape == (m,)
    assert src.min() >= 0 and src.max() < n and dst.min() >= 0 and dst.max() < n
    # reverse-complement pairing: edge 2k = (u, v), edge 2k+1 = (v^1, u^1)
    assert np.array_equal(src[1::2], dst[0::2] ^ 1) and np.array_equal(dst[1::2], src[0::2] ^ 1)
    src2, dst2 = synth.make_assembly_graph(n, m, seed=0)
    assert np.array_equal(src, src2) and np.array_equal(dst, dst2)
    outdeg = np.bincount(src, minlength=n)
    assert outdeg.max() > 20 * outdeg.mean()          # heavy tail
    x, e = synth.make_features(src, dst, n)
    assert x.shape == (n, 2) and e.shape == (m, 2)
    assert abs(float(x[:, 0].mean())) < 1e-3 and 0.9 <= float(e[:, 1].min()) and float(e[:, 1].max()) <= 1.0
