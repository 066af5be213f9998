"""GPU parity: the CUDA path (through the C ABI) against the oracle restatement and the golden
fixtures produced by the reference's own code.  Tolerance: edge probabilities within 1e-4
(BASELINE.json north_star); index work (graph staging) bit-exact."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from gnnome_b200 import synth

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-4


@pytest.fixture(scope='module')
def gnb():
    import gnnome_b200
    from gnnome_b200 import _lib
    _lib.load()  # fail loudly if the extension is missing
    return gnnome_b200


@pytest.fixture(params=['tc2', 'ffma'])
def backend(request, gnb):
    """'tc2' = TMA-fed tcgen05 kernels on split fp16 state (the product path), 'ffma' = CUDA-core fp32 kernels
    (an independent code path: the on-device cross-check)."""
    gnb.set_backend(request.param)
    yield request.param
    gnb.set_backend('tc2')


def _graph(n, m, seed):
    src, dst = synth.make_assembly_graph(n, m, seed=seed)
    x, e = synth.make_features(src, dst, n, seed=seed)
    return tuple(map(torch.from_numpy, (src, dst))) + (n,) + tuple(map(torch.from_numpy, (x, e)))


def _prob_err(a, b):
    return (torch.sigmoid(a.double().cpu()) - torch.sigmoid(b.double().cpu())).abs().max().item()


# ---------------------------------------------------------------- graph staging (bit-exact)
@pytest.mark.parametrize('n,m,seed', [(8, 0, 0), (1, 4, 0), (50, 400, 1), (3000, 18000, 2), (100000, 600000, 3)])
def test_graph_stage_exact(gnb, n, m, seed):
    rng = np.random.default_rng(seed)
    src = torch.from_numpy(rng.integers(0, n, size=m).astype(np.int32))
    dst = torch.from_numpy(rng.integers(0, n, size=m).astype(np.int32))
    gi = gnb.GraphIndex(src, dst, n)
    order = torch.argsort(dst.long(), stable=True)
    assert torch.equal(gi.in_eid[:m].cpu().long(), order)
    assert torch.equal(gi.in_src[:m].cpu(), src[order]) and torch.equal(gi.in_dst[:m].cpu(), dst[order])
    ptr = torch.zeros(n + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(torch.bincount(dst.long(), minlength=n), 0)
    assert torch.equal(gi.in_ptr.cpu().long(), ptr)
    order2 = torch.argsort(src[order].long(), stable=True)
    assert torch.equal(gi.out_pos[:m].cpu().long(), order2)
    assert torch.equal(gi.out_dst[:m].cpu(), dst[order][order2])
    ptr2 = torch.zeros(n + 1, dtype=torch.long)
    ptr2[1:] = torch.cumsum(torch.bincount(src.long(), minlength=n), 0)
    assert torch.equal(gi.out_ptr.cpu().long(), ptr2)


def test_gather_scatter_roundtrip(gnb):
    from gnnome_b200 import ops
    x = torch.randn(1000, 64, device='cuda')
    perm = torch.randperm(1000, device='cuda').to(torch.int32)
    y = ops.gather_rows(x, perm)
    assert torch.equal(y, x[perm.long()])
    assert torch.equal(ops.scatter_rows(y, perm), x)


# ---------------------------------------------------------------- individual operators
@pytest.mark.parametrize('rows,K,M', [(1, 32, 160), (130, 64, 320), (1000, 128, 640), (777, 256, 1280), (513, 256, 128)])
def test_node_linear(gnb, rows, K, M):
    from gnnome_b200 import ops
    g = torch.Generator().manual_seed(rows)
    a, w, b = torch.randn(rows, K, generator=g), torch.randn(M, K, generator=g) / K ** 0.5, torch.randn(M, generator=g)
    out = ops.node_linear(a.cuda(), w.t().contiguous().cuda(), b.cuda())
    ref = (a.double() @ w.double().t() + b.double())
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize('rows,K,M', [(1, 64, 128), (130, 64, 320), (1000, 128, 640), (777, 256, 1280), (64, 256, 128)])
def test_node_linear_tc2(gnb, rows, K, M):
    """TMA-fed edition: X given as split fp16 (hi, lo) images."""
    from gnnome_b200 import ops
    g = torch.Generator().manual_seed(rows)
    a, w, b = torch.randn(rows, K, generator=g), torch.randn(M, K, generator=g) / K ** 0.5, torch.randn(M, generator=g)
    out = ops.node_linear_tc2(ops.split_rows(a.cuda()), ops.pack_linear_tc(w.cuda()), b.cuda(), M)
    ref = (a.double() @ w.double().t() + b.double())
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize('rows,K', [(1, 64), (77, 128), (1000, 256)])
def test_split16_roundtrip(gnb, rows, K):
    """x -> (hi, lo) fp16 images -> x keeps 22 significant bits (absolute floor 2^-20 from fp16 subnormals)."""
    from gnnome_b200 import ops
    x = torch.randn(rows, K, device='cuda') * 5
    idx = torch.randperm(rows, device='cuda').to(torch.int32)
    y = ops.merge_rows(ops.split_rows(x))
    assert ((y - x).abs() <= x.abs() * 2.0 ** -22 + 2.0 ** -20).all()
    y2 = ops.merge_rows(ops.split_rows(x, idx), idx)       # gather on the way in, scatter on the way out
    assert torch.equal(y, y2)


@pytest.mark.parametrize('H,hs', [(64, 64), (128, 32), (256, 64), (256, 128)])
def test_score_tc2_vs_cuda_core(gnb, H, hs):
    from gnnome_b200 import ops
    src, dst, n, _, _ = _graph(3000, 18002, seed=hs)
    gi = gnb.GraphIndex(src, dst, n)
    torch.manual_seed(H + hs)
    pred = gnb.layers.ScorePredictor(H, hs)
    x, e = torch.randn(n, H, device='cuda'), torch.randn(gi.E, H, device='cuda')
    with torch.no_grad():
        S = pred.node_rows16(ops.split_rows(x))
        e16 = ops.split_rows(e)
        a = pred.score_positions16(gi, S, e16, tensor_cores=True)
        b = pred.score_positions16(gi, S, e16, tensor_cores=False)
        p = {'predictor.' + k: v.double() for k, v in pred.state_dict().items()}
        ref = R.score_predictor(p, 'predictor.', src.long(), dst.long(), x.cpu().double(), e.cpu().double()[torch.argsort(gi.in_eid[:gi.E].cpu().long())])
    assert (a - b).abs().max().item() < 2e-5
    assert (a.cpu().double() - ref).abs().max().item() < 5e-5


@pytest.mark.parametrize('scale', [1e-3, 1.0, 300.0, 3e4])
def test_node_linear_tc_dynamic_range(gnb, scale):
    """The fp16 split keeps fp32-level accuracy up to |x| ~ 1e5; below |x| ~ 1 the fp16 subnormal
    range puts an ABSOLUTE floor of ~2^-21 per input element on the error."""
    from gnnome_b200 import ops
    g = torch.Generator().manual_seed(7)
    a = torch.randn(300, 256, generator=g) * scale
    w, b = torch.randn(640, 256, generator=g) / 16, torch.zeros(640)
    out = ops.node_linear_tc2(ops.split_rows(a.cuda()), ops.pack_linear_tc(w.cuda()), b.cuda(), 640)
    ref = a.double() @ w.double().t()
    assert (out.cpu().double() - ref).abs().max().item() < 1e-5 * max(scale, 1.0)


@pytest.mark.parametrize('rows,H', [(5, 64), (1000, 256), (64, 32)])
def test_encode(gnb, rows, H):
    from gnnome_b200 import ops
    g = torch.Generator().manual_seed(H)
    x = torch.randn(rows + 7, 2, generator=g)
    idx = torch.randint(0, rows + 7, (rows,), generator=g).to(torch.int32)
    W1, b1 = torch.randn(16, 2, generator=g), torch.randn(16, generator=g)
    W2, b2 = torch.randn(H, 16, generator=g), torch.randn(H, generator=g)
    ref = torch.relu(x[idx.long()].double() @ W1.double().t() + b1.double()) @ W2.double().t() + b2.double()
    out = ops.encode(x.cuda(), idx.cuda(), W1.cuda(), b1.cuda(), W2.t().contiguous().cuda(), b2.cuda(), rows)
    assert (out.cpu().double() - ref).abs().max().item() < 1e-5
    if H >= 64:
        o16, o32 = ops.encode2(x.cuda(), idx.cuda(), W1.cuda(), b1.cuda(), W2.t().contiguous().cuda(), b2.cuda(), rows,
                               want16=True, want32=True)
        assert (o32.cpu().double() - ref).abs().max().item() < 1e-5
        assert (ops.merge_rows(o16).cpu().double() - ref).abs().max().item() < 2e-5
    out2 = ops.encode(x[:rows].contiguous().cuda(), None, W1.cuda(), b1.cuda(), W2.t().contiguous().cuda(), b2.cuda(), rows)
    ref2 = torch.relu(x[:rows].double() @ W1.double().t() + b1.double()) @ W2.double().t() + b2.double()
    assert (out2.cpu().double() - ref2).abs().max().item() < 1e-5


# ---------------------------------------------------------------- one layer, all widths, nasty graphs
def _hub_graph(n, m, hub_deg, seed):
    """Random graph plus one node with ``hub_deg`` in-edges and ``hub_deg`` out-edges (segments that
    span many aggregation chunks), isolated nodes, self loops and multi-edges."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n // 2, size=m)
    dst = rng.integers(0, n // 2, size=m)          # nodes >= n/2 stay isolated except the hub edges
    hub = n // 3
    src = np.concatenate([src, rng.integers(0, n, size=hub_deg), np.full(hub_deg, hub), [5, 5, 5]])
    dst = np.concatenate([dst, np.full(hub_deg, hub), rng.integers(0, n, size=hub_deg), [5, 6, 6]])
    return torch.from_numpy(src.astype(np.int32)), torch.from_numpy(dst.astype(np.int32))


@pytest.mark.parametrize('H', [32, 64, 128, 256])
@pytest.mark.parametrize('sym', [True, False])
def test_single_layer_vs_oracle(gnb, backend, H, sym):
    n = 700
    src, dst = _hub_graph(n, 3000, 1500, seed=H)
    m = src.numel()
    torch.manual_seed(H + sym)
    layer = (gnb.layers.SymGatedGCN if sym else gnb.layers.GatedGCN)(H, H, 'batch')
    with torch.no_grad():
        for bn in (layer.bn_h, layer.bn_e):
            bn.running_mean.normal_(0, 0.3)
            bn.running_var.uniform_(0.05, 2.0)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
    layer.eval()
    h, e = torch.randn(n, H), torch.randn(m, H)
    p = {'L.' + k: v for k, v in layer.state_dict().items()}
    fn = R.sym_gated_gcn_layer if sym else R.gated_gcn_layer
    with torch.no_grad():
        p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
        h_ref, e_ref = fn(p64, 'L.', src.long(), dst.long(), n, h.double(), e.double())
        h_out, e_out = layer((src, dst, n), h.cuda(), e.cuda())
    eh = (h_out.cpu().double() - h_ref).abs().max().item()
    ee = (e_out.cpu().double() - e_ref).abs().max().item()
    assert ee < 5e-5 * max(1.0, e_ref.abs().max().item()), f'e err {ee}'
    assert eh < 5e-5 * max(1.0, h_ref.abs().max().item()), f'h err {eh}'


# ---------------------------------------------------------------- whole model vs reference goldens
@pytest.mark.parametrize('name', ['sym_shipped_tiny', 'sym_shipped_2k'])
def test_shipped_model_vs_reference_golden(gnb, backend, golden, shipped_weights, name):
    g = golden(name)
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    model.load_state_dict(shipped_weights, strict=True)
    model.eval()
    with torch.no_grad():
        out = model((g['src'], g['dst'], g['num_nodes']), g['x'], g['e'])   # CPU inputs, like inference.py:388
    assert out.shape == g['logits'].shape and out.dtype == torch.float32 and out.device.type == 'cpu'
    err = _prob_err(out, g['logits'])
    rel = ((out - g['logits']).abs() / g['logits'].abs().clamp_min(1.0)).max().item()
    assert err <= PROB_TOL, f'prob err {err}'
    assert rel <= 1e-3, f'logit rel err {rel}'


@pytest.mark.parametrize('name,ctor', [
    ('gated_directed', lambda m: m.GatedGCNModel(2, 2, 32, 16, 3, 64, 'batch', directed=True)),
    ('gated_undirected', lambda m: m.GatedGCNModel(2, 2, 32, 16, 3, 64, 'batch', directed=False)),
    ('sym_h128', lambda m: m.SymGatedGCNModel(2, 2, 128, 16, 2, 64, 'batch')),
])
def test_other_models_vs_reference_golden(gnb, golden, name, ctor):
    g = golden(name)
    model = ctor(gnb.models)
    model.load_state_dict(g['state_dict'], strict=True)
    model.eval()
    with torch.no_grad():
        out = model((g['src'], g['dst'], g['num_nodes']), g['x'].cuda(), g['e'].cuda())
    assert out.is_cuda and out.shape == g['logits'].shape
    assert _prob_err(out, g['logits']) <= PROB_TOL
    assert (out.cpu() - g['logits']).abs().max().item() <= 1e-4


@pytest.mark.parametrize('H,L', [(64, 3), (128, 2), (256, 2)])
def test_undirected_model_on_the_split16_path(gnb, H, L):
    """GatedGCNModel(directed=False) (models/full_graph.py:47-52: add_reverse_edges, e = cat(e, e), e[:E]) at the widths
    of the tcgen05 kernels: the doubled graph runs on the split16 path and the original edges' rows are picked out of
    the fp16 (hi, lo) images.  Against the oracle and against the CUDA-core fp32 kernels (the path the reference's
    hidden-32 fixture exercises, test_other_models_vs_reference_golden)."""
    src, dst, n, x, e = _graph(5000, 30000, seed=H + 1)
    sd = R.init_state_dict(model='gated', hidden=H, num_layers=L, seed=H)
    model = gnb.models.GatedGCNModel(2, 2, H, 16, L, 64, 'batch', directed=False)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        out = model((src, dst, n), x.cuda(), e.cuda())
        ref = R.model_forward(sd, src, dst, n, x, e, model='gated', directed=False)
        gnb.set_backend('ffma')
        try:
            out32 = model((src, dst, n), x.cuda(), e.cuda())
        finally:
            gnb.set_backend('tc2')
    assert out.shape == (src.numel(), 1) and out.is_cuda
    assert _prob_err(out, ref) <= PROB_TOL
    assert _prob_err(out, out32) <= 2e-5


@pytest.mark.parametrize('H,L,n,m', [(64, 8, 20000, 120000), (128, 4, 10000, 60000), (256, 3, 6000, 36000)])
def test_model_vs_oracle_fp64(gnb, backend, shipped_weights, H, L, n, m):
    """Against the fp64 evaluation of the oracle ("true value"): our error must stay within the
    tolerance and within a small multiple of the fp32 reference's own error."""
    src, dst, n, x, e = _graph(n, m, seed=H)
    sd = shipped_weights if H == 64 else R.init_state_dict(hidden=H, num_layers=L, seed=H)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch')
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        out = model((src, dst, n), x, e)
        truth = R.model_forward(sd, src, dst, n, x, e, dtype=torch.float64, faithful=False)
        ref32 = R.model_forward(sd, src, dst, n, x, e, faithful=False)
    ours, theirs = _prob_err(out, truth), _prob_err(ref32, truth)
    assert ours <= PROB_TOL, f'ours {ours} (fp32 reference: {theirs})'
    assert ours <= max(4 * theirs, 2e-5), f'ours {ours} vs fp32 reference {theirs}'


# ---------------------------------------------------------------- size-independent properties
def test_deterministic_and_graph_cache(gnb, shipped_weights):
    src, dst, n, x, e = _graph(50000, 300000, seed=5)
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch')
    model.load_state_dict(shipped_weights)
    model.eval()
    gi = gnb.GraphIndex(src, dst, n)
    with torch.no_grad():
        a = model(gi, x.cuda(), e.cuda())
        b = model(gi, x.cuda(), e.cuda())
    assert torch.equal(a, b)          # fixed summation order: bit-reproducible
    assert torch.isfinite(a).all()


def test_edge_relabelling_equivariance(gnb, shipped_weights):
    """Permuting the edge ids permutes the scores (the function does not depend on edge order
    beyond fp32 summation order)."""
    src, dst, n, x, e = _graph(4000, 24000, seed=9)
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch')
    model.load_state_dict(shipped_weights)
    model.eval()
    perm = torch.randperm(src.numel(), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = model((src, dst, n), x, e)
        b = model((src[perm], dst[perm], n), x, e[perm])
    assert _prob_err(a[perm], b) <= PROB_TOL


def test_empty_and_degenerate_graphs(gnb, backend, shipped_weights):
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch')
    model.load_state_dict(shipped_weights)
    model.eval()
    with torch.no_grad():
        out = model((torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), 5),
                    torch.randn(5, 2), torch.zeros(0, 2))
        assert out.shape == (0, 1)
        # a single self loop
        s = torch.tensor([0], dtype=torch.int32)
        x, e = torch.randn(1, 2), torch.randn(1, 2)
        out = model((s, s, 1), x, e)
        ref = R.model_forward(shipped_weights, s, s, 1, x, e)
    assert _prob_err(out, ref) <= PROB_TOL


def test_full_size_properties_cfg2(gnb):
    """BASELINE config 2 (1M nodes / 6M edges, H=128, L=8) is too big for the oracle: size-independent properties
    instead -- bit-reproducible, finite, and equivariant under a relabelling of the edges."""
    n, m, H, L = 1_000_000, 6_000_000, 128, 8
    src, dst, n, x, e = _graph(n, m, seed=0)
    torch.manual_seed(0)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').cuda().eval()
    xd, ed = x.cuda(), e.cuda()
    with torch.no_grad():
        gi = gnb.GraphIndex(src, dst, n)
        a = model(gi, xd, ed)
        b = model(gi, xd, ed)
        assert torch.equal(a, b) and torch.isfinite(a).all() and a.shape == (m, 1)
        perm = torch.randperm(m, generator=torch.Generator().manual_seed(1))
        c = model((src[perm], dst[perm], n), xd, ed[perm.cuda()])
    assert _prob_err(a[perm.cuda()], c) <= PROB_TOL


def test_cross_backend_at_cfg2_full_size(gnb):
    """Parity at the measured size (the oracle cannot hold it): BASELINE config 2 in full -- 1M nodes / 6M edges, H=128,
    L=8 -- on the tcgen05 / split-fp16 product path against the CUDA-core fp32 kernels, an independent code path (own
    graph walk, fp32 state, FFMA products).  int32 positions x row strides, 25 k tiles per CTA and the L2-evicting
    working set only exist at this size."""
    n, m, H, L = 1_000_000, 6_000_000, 128, 8
    src, dst, n, x, e = _graph(n, m, seed=0)
    torch.manual_seed(0)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').cuda().eval()
    with torch.no_grad():
        gi = gnb.GraphIndex(src, dst, n)
        xd, ed = x.cuda(), e.cuda()
        a = model(gi, xd, ed)
        gnb.set_backend('ffma')
        try:
            b = model(gi, xd, ed)
        finally:
            gnb.set_backend('tc2')
    assert _prob_err(a, b) <= 2e-5


def test_model_vs_oracle_h256_l8_at_1m2_edges(gnb):
    """The largest oracle comparison the host can hold at the benchmark's model shape: H=256, L=8 on 200k nodes /
    1.2M edges (~17 GB of host RAM on the reference path), probabilities within 1e-4 (measured ~3e-6)."""
    n, m, H, L = 200_000, 1_200_000, 256, 8
    src, dst, n, x, e = _graph(n, m, seed=1)
    torch.manual_seed(0)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        out = model.cuda()((src, dst, n), x.cuda(), e.cuda())
        ref = R.model_forward(sd, src, dst, n, x, e, faithful=False)
    assert _prob_err(out, ref) <= PROB_TOL


def test_cfg1_standin_shipped_weights_ecoli_sized(gnb, shipped_weights):
    """BASELINE config 1 stand-in: the real E. coli graph needs hifiasm + Biopython + DGL;
    an E. coli-sized synthetic assembly graph (15k nodes / 100k edges) with the reference's shipped weights."""
    src, dst, n, x, e = _graph(15_000, 100_000, seed=7)
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch')
    model.load_state_dict(shipped_weights)
    model = model.cuda().eval()
    with torch.no_grad():
        out = model((src, dst, n), x.cuda(), e.cuda())
        ref = R.model_forward(shipped_weights, src, dst, n, x, e, faithful=True)
    assert _prob_err(out, ref) <= PROB_TOL
