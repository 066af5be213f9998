"""GPU parity of the graph hand-off: raw assembly graph -> device-built features ->
model -> ``{idx}_predicts.pt`` / training losses, against fixtures produced by the reference's OWN functions
(oracle/make_golden_handoff.py).  Tolerances: features 5e-6 (the reference accumulates mean / std in fp32, the kernels in
fp64), edge probabilities 1e-4, loss 1e-5 relative, gradients as in test_gpu_train.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gnb():
    import gnnome_b200
    from gnnome_b200 import _lib
    _lib.load()
    return gnnome_b200


def _graph(rec):
    from gnnome_b200.assembly import AssemblyGraph
    r = rec['raw']
    return AssemblyGraph(r['src'], r['dst'], r['num_nodes'],
                         dict(overlap_length=r['overlap_length'], overlap_similarity=r['overlap_similarity'], y=r['y']))


def test_degree_rows_exact(gnb):
    from gnnome_b200 import ops
    rng = np.random.default_rng(0)
    n, m = 5000, 40000
    src = torch.from_numpy(rng.integers(0, n - 7, m).astype(np.int32))       # the last nodes stay isolated
    dst = torch.from_numpy(rng.integers(0, n - 7, m).astype(np.int32))
    gi = gnb.GraphIndex(src, dst, n)
    ind = torch.bincount(dst.long(), minlength=n).float()
    outd = torch.bincount(src.long(), minlength=n).float()
    assert torch.equal(ops.degree_rows(gi).cpu(), torch.stack((ind, outd), 1))
    assert torch.equal(ops.degree_rows(gi, swap=True).cpu(), torch.stack((outd, ind), 1))
    assert torch.equal(ops.degree_rows(gi.reversed()).cpu(), torch.stack((outd, ind), 1))


@pytest.mark.parametrize('rows,cols', [(1000, 1), (9001, 2), (3_000_000, 2), (70_000, 4)])
def test_zscore_cols_vs_fp64(gnb, rows, cols):
    from gnnome_b200 import ops
    rng = np.random.default_rng(rows)
    v = (rng.standard_normal((rows, cols)) * [3.0, 0.03, 700.0, 1.0][:cols] + [1.0e4, 0.95, -5.0, 0.0][:cols]).astype(np.float32)
    mask = (1 << cols) - 1 if cols != 2 else 0b01
    x, stats = ops.zscore_cols(torch.from_numpy(v).cuda(), mask, want_stats=True)
    mean, std = v.astype(np.float64).mean(0), v.astype(np.float64).std(0, ddof=1)
    np.testing.assert_allclose(stats.cpu().numpy()[:, 0], mean, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(stats.cpu().numpy()[:, 1], std, rtol=1e-10)
    want = (v - mean.astype(np.float32)) / std.astype(np.float32)          # fp32 normalise, like the reference
    got = x.cpu().numpy()
    for c in range(cols):
        if (mask >> c) & 1:
            np.testing.assert_allclose(got[:, c], want[:, c], rtol=1e-6, atol=1e-6)
        else:
            assert np.array_equal(got[:, c], v[:, c])                      # unmasked columns are not touched
    # and against what torch does on the host in fp32 (the reference).  Column 0 sits at 1e4 with std 3, where one fp32
    # ulp of the mean is 3e-4 of a z-score, so this bounds the reference's own rounding, not the kernel's
    if rows <= 10 ** 5:
        t = torch.from_numpy(v[:, 0])
        np.testing.assert_allclose(got[:, 0], ((t - t.mean()) / t.std()).numpy(), rtol=0, atol=5e-3)


def test_zscore_degenerate_inputs_like_torch(gnb):
    from gnnome_b200 import ops
    const = ops.zscore_cols(torch.full((100, 2), 3.0, device='cuda'), 0b01)
    assert torch.isnan(const[:, 0]).all() and (const[:, 1] == 3.0).all()      # 0 / 0, as (v - v.mean()) / v.std()
    one = ops.zscore_cols(torch.tensor([[2.0, 5.0]], device='cuda'), 0b11)
    assert torch.isnan(one).all()                                            # unbiased std of one row
    empty = ops.zscore_cols(torch.empty((0, 2), device='cuda'), 0b11)
    assert empty.shape == (0, 2)
    with pytest.raises(ValueError):
        ops.zscore_cols(torch.zeros((4, 5), device='cuda'), 1)
    with pytest.raises(RuntimeError, match='col_mask'):
        ops.zscore_cols(torch.zeros((4, 2), device='cuda'), 0b100)


def test_features_vs_reference_functions(gnb, golden):
    from gnnome_b200 import assembly as A
    g = golden('handoff_scores')
    ag = _graph(g)
    A.add_positional_encoding(A.preprocess_graph(ag))
    assert ag.edata['e'].is_cuda and ag.edata['e'].shape == g['e'].shape
    torch.testing.assert_close(ag.edata['e'].cpu(), g['e'], rtol=5e-6, atol=5e-6)
    assert torch.equal(ag.edata['e'][:, 1].cpu(), g['e'][:, 1])              # the similarity column passes through
    x, e = A.get_full_ne_features(ag)
    torch.testing.assert_close(x.cpu(), g['x'], rtol=5e-6, atol=5e-6)
    x_rev, _ = A.get_full_ne_features(ag.reversed(), reverse=True)           # ndata carried over from the original
    torch.testing.assert_close(x_rev.cpu(), g['x_rev'], rtol=5e-6, atol=5e-6)
    assert torch.equal(x_rev, x.flip(1))
    one_col = A.preprocess_graph(_graph(g), use_similarities=False).edata['e']
    assert one_col.shape == (ag.num_edges(), 1) and torch.equal(one_col[:, 0], ag.edata['e'][:, 0])


def test_compute_scores_writes_predicts_like_inference(gnb, golden, shipped_weights, tmp_path):
    from gnnome_b200 import assembly as A
    g = golden('handoff_scores')
    ag = _graph(g)
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=0.2)
    model.load_state_dict(shipped_weights, strict=True)
    with pytest.raises(RuntimeError, match='eval'):
        A.compute_scores(model.cuda().train(), ag)
    model.eval()
    out_dir = tmp_path / 'decode'
    scores = A.compute_scores(model, ag, idx=3, inference_dir=str(out_dir))
    want = g['predicts']
    assert scores.device.type == 'cpu' and scores.dtype == torch.float32 and scores.shape == want.shape
    assert (torch.sigmoid(scores.double()) - torch.sigmoid(want.double())).abs().max().item() <= 1e-4
    assert ((scores - want).abs() <= 1e-3 * want.abs().clamp_min(1.0)).all()
    saved = torch.load(out_dir / '3_predicts.pt', weights_only=True)
    assert torch.equal(saved, scores) and torch.equal(ag.edata['score'], scores)
    # an existing predictions file wins over the model (inference.py:429-431)
    torch.save(torch.arange(5.0), out_dir / '3_predicts.pt')
    assert torch.equal(A.compute_scores(None, ag, idx=3, inference_dir=str(out_dir)), torch.arange(5.0))


@pytest.mark.parametrize('kind', ['bce', 'sym'])
def test_training_losses_vs_reference_functions(gnb, golden, shipped_weights, kind):
    from gnnome_b200 import assembly as A
    g = golden('handoff_losses')
    ag = _graph(g)
    model = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    model.load_state_dict(shipped_weights, strict=True)
    model.cuda().train()
    pw = torch.tensor([g['pos_weight']], device='cuda')
    ref = g[kind]
    seen = []
    hook = model.register_forward_hook(lambda mod, inp, out: seen.append(out.squeeze(-1)))
    if kind == 'bce':
        loss, logits = A.get_bce_loss_full(ag, model, pw)
        loss.backward()
    else:
        loss, logits = A.get_symmetry_loss_full(ag, model, pw, g['alpha'])
        # |org - rev| has a kink where the two forwards agree to fp32 noise (28 of the 3600 edges are within 1e-4, the
        # closest at 2e-6): there the reference's sign(org - rev) is a coin toss, and one flipped edge moves the first
        # layers' gradients by more than the tolerance.  Compare gradients for the reference's choice of subgradient.
        org, rev = seen
        ref_org, ref_rev = (t.cuda() for t in ref['forwards'])
        assert (torch.sigmoid(rev.detach().double()) - torch.sigmoid(ref_rev.double())).abs().max().item() <= 1e-4
        sign_ref = torch.sign(ref_org - ref_rev)
        flips = int((torch.sign(org.detach() - rev.detach()) != sign_ref).sum())
        print('edges whose sign(org - rev) differs from the reference:', flips)
        assert flips <= 8
        y = ag.edata['y'].cuda()
        bce = lambda s: torch.nn.functional.binary_cross_entropy_with_logits(s, y, pos_weight=pw, reduction='none')  # noqa: E731
        pinned = (bce(org) + bce(rev) + g['alpha'] * (org - rev) * sign_ref).mean()
        assert abs(pinned.item() - loss.item()) <= 1e-6
        pinned.backward()
    hook.remove()
    assert logits.shape == ref['logits'].shape
    assert (torch.sigmoid(logits.detach().cpu().double()) - torch.sigmoid(ref['logits'].double())).abs().max().item() <= 1e-4
    assert abs(loss.item() - ref['loss'].item()) <= 1e-5 * max(1.0, abs(ref['loss'].item()))
    # Gradients.  The network has ~2 M ReLU inputs per forward and a handful of them sit within fp32 noise of zero; when
    # an fp32 evaluation lands on the other side of such a kink than the reference's did, it returns a different (equally
    # valid) subgradient.  On this fixture one flipped element of layer 1 (|ehat| = 3e-7, tests/diag/grad_dump_gate.py)
    # moves the first layers' gradients by 1e-3 .. 6e-2 of their scale although every kernel reproduces its own inputs
    # to 1e-7.  So: the strict per-parameter bound holds for the single-forward case (no flip on that fixture); with two
    # forwards the per-parameter bound is loose and the overall direction is checked in the L2 norm.
    per_param = 2e-3 if kind == 'bce' else 1e-1
    num = den = 0.0
    for k, p in model.named_parameters():
        r = ref['grads'][k]
        err = (p.grad.cpu() - r).abs().max().item()
        scale = max(r.abs().max().item(), 1e-6)
        assert err <= per_param * scale + 5e-6, f'{k}: grad err {err} (scale {scale})'
        num += (p.grad.cpu().double() - r.double()).square().sum().item()
        den += r.double().square().sum().item()
    print('relative L2 error of the whole gradient:', (num / den) ** 0.5)
    assert (num / den) ** 0.5 <= 1e-2
    for k, b in model.named_buffers():
        r = ref['buffers'][k]
        if r.dtype == torch.long:
            assert int(b) == int(r), k            # bn_e: +2 per forward; the symmetric loss runs two forwards
        else:
            torch.testing.assert_close(b.cpu(), r, rtol=1e-4, atol=1e-5, msg=k)
    if kind == 'sym':                             # the reversed graph is staged once and reused
        g_rev = ag._gnb_reversed
        A.get_symmetry_loss_full(ag, model, pw, g['alpha'])
        assert ag._gnb_reversed is g_rev and g_rev._gnb_index_cache is not None


def test_dataset_directory_like_graph_dataset(gnb, golden, tmp_path):
    from gnnome_b200 import assembly as A
    proc = tmp_path / 'hifiasm' / 'processed'
    os.makedirs(proc)
    ga, gb = _graph(golden('handoff_losses')), _graph(golden('handoff_scores'))
    gb.save(proc / '10.pt')
    ga.save(proc / '2.pt')
    (proc / 'notes.txt').write_text('ignored')
    ds = A.AssemblyGraphDataset(str(tmp_path), 'hifiasm')
    assert len(ds) == 2 and [i for i, _ in ds] == [2, 10]
    idx, g = ds[1]
    assert g.num_edges() == gb.num_edges() and g.edata['e'].is_cuda and g.ndata['in_deg'].shape == (gb.num_nodes(),)
    torch.testing.assert_close(g.edata['e'].cpu(), golden('handoff_scores')['e'], rtol=5e-6, atol=5e-6)


# ---- induced subgraphs: strand-wise masking and mini-batches -----------------------------

def _check_losses(model_ctor, fn, ref, tol=1e-5):
    """Run ``fn(model) -> (loss, logits)`` on a fresh train-mode model; compare loss / every forward with the fixture."""
    model = model_ctor()
    seen = []
    hook = model.register_forward_hook(lambda mod, inp, out: seen.append(out.detach().squeeze(-1)))
    loss, logits = fn(model)
    hook.remove()
    assert len(seen) == len(ref['forwards']) and logits.shape == ref['logits'].shape
    for ours, theirs in zip(seen, ref['forwards']):
        assert (torch.sigmoid(ours.cpu().double()) - torch.sigmoid(theirs.double())).abs().max().item() <= 1e-4
    assert abs(loss.item() - ref['loss'].item()) <= tol * max(1.0, abs(ref['loss'].item()))
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_node_subgraph_bit_exact_vs_reference_masking(gnb, golden):
    from gnnome_b200 import assembly as A
    g = golden('handoff_subgraphs')
    ag = _graph(g)
    A.add_positional_encoding(A.preprocess_graph(ag))
    sub = A.node_subgraph(ag, g['mask_keep'])
    assert sub.num_nodes() == g['mask_node_id'].numel() and sub.num_edges() == g['mask_edge_id'].numel()
    assert torch.equal(sub.ndata['_ID'].cpu(), g['mask_node_id']) and torch.equal(sub.edata['_ID'].cpu(), g['mask_edge_id'])
    assert torch.equal(sub.edges()[0].cpu(), g['mask_src']) and torch.equal(sub.edges()[1].cpu(), g['mask_dst'])
    assert torch.equal(sub.ndata['in_deg'].cpu(), g['mask_in_deg'])            # sliced from the parent, not recomputed
    assert torch.equal(sub.edata['y'].cpu(), g['mask_y'])
    torch.testing.assert_close(sub.edata['e'].cpu(), g['mask_e'], rtol=5e-6, atol=5e-6)
    batch = A.node_subgraph(ag, g['batch_keep'])
    assert torch.equal(batch.ndata['_ID'].cpu(), g['batch_node_id']) and torch.equal(batch.edata['_ID'].cpu(), g['batch_edge_id'])
    x, e = A.get_partition_ne_features(batch, ag)
    torch.testing.assert_close(x.cpu(), g['batch_x'], rtol=5e-6, atol=5e-6)
    torch.testing.assert_close(e.cpu(), g['batch_e'], rtol=5e-6, atol=5e-6)


@pytest.mark.parametrize('n,m,frac', [(1, 0, 1.0), (7, 30, 0.5), (1024, 4096, 0.0), (1024, 4096, 1.0),
                                      (50_001, 300_007, 0.8), (2_000_000, 12_000_003, 0.85)])
def test_node_subgraph_properties(gnb, n, m, frac):
    """Sizes straddling the 1024-item blocks, empty and full masks, and a 12 M-edge graph: against torch on the device."""
    from gnnome_b200 import ops
    gen = torch.Generator(device='cuda').manual_seed(n + m)
    src = torch.randint(0, n, (m,), device='cuda', generator=gen, dtype=torch.int32)
    dst = torch.randint(0, n, (m,), device='cuda', generator=gen, dtype=torch.int32)
    keep = torch.rand(n, device='cuda', generator=gen) < frac
    node_id, edge_id, s, d = ops.node_subgraph(keep, src, dst, n)
    want_nodes = torch.nonzero(keep).flatten()
    ek = keep[src.long()] & keep[dst.long()]
    want_edges = torch.nonzero(ek).flatten()
    assert torch.equal(node_id.long(), want_nodes) and torch.equal(edge_id.long(), want_edges)
    assert node_id.dtype == torch.int32 and s.dtype == torch.int32
    if want_edges.numel():
        assert torch.equal(node_id[s.long()], src[want_edges]) and torch.equal(node_id[d.long()], dst[want_edges])


def test_mask_graph_strandwise_keeps_strand_pairs(gnb, golden):
    from gnnome_b200 import assembly as A
    ag = _graph(golden('handoff_subgraphs'))
    gen = torch.Generator(device='cuda').manual_seed(3)
    sub = A.mask_graph_strandwise(ag, 0.8, generator=gen)
    ids = sub.ndata['_ID'].cpu()
    assert ids.numel() % 2 == 0 and torch.equal(ids[0::2] + 1, ids[1::2]) and bool((ids[0::2] % 2 == 0).all())
    assert 0.7 < ids.numel() / ag.num_nodes() < 0.9
    assert torch.equal(sub.edata['y'].cpu(), ag.edata['y'][sub.edata['_ID'].cpu().long()])
    assert A.mask_graph_strandwise(ag, 1.0).num_edges() == ag.num_edges()


def test_losses_on_masked_graph_and_mini_batch(gnb, golden, shipped_weights):
    from gnnome_b200 import assembly as A
    g = golden('handoff_subgraphs')
    ag = _graph(g)
    A.add_positional_encoding(A.preprocess_graph(ag))
    pw = torch.tensor([g['pos_weight']], device='cuda')

    def ctor():
        m = gnb.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
        m.load_state_dict(shipped_weights, strict=True)
        return m.cuda().train()

    masked = A.node_subgraph(ag, g['mask_keep'])
    _check_losses(ctor, lambda m: A.get_symmetry_loss_full(masked, m, pw, g['alpha']), g['mask_sym'])
    batch = A.node_subgraph(ag, g['batch_keep'])
    _check_losses(ctor, lambda m: A.get_bce_loss_partition(batch, ag, m, pw), g['batch_bce'])
    _check_losses(ctor, lambda m: A.get_symmetry_loss_partition(batch, ag, m, pw, g['alpha']), g['batch_sym'])
