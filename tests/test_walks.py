"""Greedy decoder walks: the C++ walker and the oracle's restatement against walks produced
by the reference's OWN ``greedy_forwards`` / ``greedy_backwards_rc`` / ``run_greedy_both_ways`` (oracle/make_golden_walks.py).
Host code: runs without a GPU.  Bit-exact: the same nodes in the same order, the same float32 sums."""
import numpy as np
import pytest
import torch

from oracle import restatement as R


@pytest.fixture(scope='module')
def fx(golden):
    g = golden('walks')
    src, dst, n = g['src'], g['dst'], g['num_nodes']
    succs, preds, edges = {}, {}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs.setdefault(u, []).append(v)
        preds.setdefault(v, []).append(u)
        edges[(u, v)] = k
    g.update(succs=succs, preds=preds, edges=edges, log_probs=torch.log(torch.sigmoid(g['scores'])))
    return g


@pytest.mark.parametrize('key,use_visited', [('results', True), ('results_nothing_visited', False)])
def test_oracle_walks_match_reference_functions(fx, key, use_visited):
    visited = set(fx['visited']) if use_visited else set()
    for (s, d), ref in zip(fx['candidates'], fx[key]):
        walk_f, walk_b, sum_f, sum_b = R.run_greedy_both_ways(s, d, fx['log_probs'], fx['succs'], fx['edges'], visited)
        assert walk_f == ref['walk_f'] and walk_b == ref['walk_b']
        assert torch.equal(sum_f, ref['sum_f']) and torch.equal(sum_b, ref['sum_b'])


@pytest.mark.parametrize('build', ['edge_list', 'dicts'])
@pytest.mark.parametrize('key,use_visited', [('results', True), ('results_nothing_visited', False)])
@pytest.mark.parametrize('threads', [1, 4])
def test_walker_bit_exact_vs_reference_functions(fx, build, key, use_visited, threads):
    from gnnome_b200.decode import WalkGraph
    n = fx['num_nodes']
    wg = (WalkGraph.from_edge_list(fx['src'], fx['dst'], n) if build == 'edge_list'
          else WalkGraph.from_dicts(n, fx['succs'], fx['edges'], fx['preds']))
    visited = set(fx['visited']) if use_visited else None
    res = wg.run_greedy_both_ways(fx['candidates'], fx['log_probs'], visited, threads=threads)
    assert len(res) == len(fx[key])
    for (walk_f, walk_b, sum_f, sum_b), ref in zip(res, fx[key]):
        assert walk_f == ref['walk_f'] and walk_b == ref['walk_b']
        assert np.float32(sum_f) == ref['sum_f'].numpy()[0] and np.float32(sum_b) == ref['sum_b'].numpy()[0]   # bit-equal
        walk = walk_b + walk_f
        assert len(set(walk) | {w ^ 1 for w in walk}) == ref['n_visited']
        assert wg.get_contig_length(walk, fx['prefix_length'], fx['read_length']) == ref['contig_length']
        assert sorted(wg.jumped_nodes(walk)) == ref['jumped']


def test_walker_edge_cases(fx):
    from gnnome_b200.decode import WalkGraph
    # 0 -> 2 -> 4, 0 -> 6 (dead end), 4 has no successor; strand twins 1, 3, 5, 7 isolated
    wg = WalkGraph.from_edge_list([0, 2, 0], [2, 4, 6], 8)
    logp = torch.log(torch.tensor([0.9, 0.8, 0.1]))
    (walk_f, walk_b, sum_f, sum_b), = wg.run_greedy_both_ways([(0, 2)], logp)
    assert walk_f == [2, 4] and walk_b == [0] and sum_b == 0.0          # backward: start 0 ^ 1 = 1 has no successors
    assert np.float32(sum_f) == np.float32(np.log(np.float32(0.8)))
    assert wg.run_greedy_both_ways([], logp) == []
    # visited as a mask; a visited successor stops the walk
    mask = torch.zeros(8, dtype=torch.bool)
    mask[4] = True
    (walk_f, _, sum_f, _), = wg.run_greedy_both_ways([(0, 2)], logp, mask)
    assert walk_f == [2] and sum_f == 0.0
    # exact tie: the first successor in list order wins, as torch.topk does
    wt = WalkGraph.from_edge_list([0, 0, 0], [2, 4, 6], 8)
    (walk_f, _, _, _), = wt.run_greedy_both_ways([(1, 0)], torch.log(torch.tensor([0.5, 0.5, 0.5])))
    assert walk_f == [0, 2]
    with pytest.raises(ValueError, match='strand pairs'):
        WalkGraph.from_edge_list([0], [1], 3)
    with pytest.raises((RuntimeError, IndexError), match='out of range|outside'):
        wg.run_greedy_both_ways([(0, 9)], logp)
    # a score vector shorter than the edge ids the lists refer to (stale {idx}_predicts.pt): the reference raises
    # IndexError on logProbs[edges[...]]; reading past the end must never happen silently
    with pytest.raises(IndexError, match='log_probs has 1 entries'):
        wg.run_greedy_both_ways([(0, 2)], logp[:1])
    with pytest.raises(IndexError, match='prefix_length has 2 entries'):
        wg.get_contig_length([0, 2, 4], torch.ones(2, dtype=torch.int64), torch.ones(8, dtype=torch.int64))
    with pytest.raises(IndexError, match='neighbour id outside'):
        WalkGraph(4, (np.array([0, 1, 1, 1, 1]), np.array([7], dtype=np.int32), np.array([0], dtype=np.int32)))
    with pytest.raises(RuntimeError, match='no edge'):
        wg.get_contig_length([0, 4], torch.ones(3, dtype=torch.int64), torch.ones(8, dtype=torch.int64))
    assert wg.get_contig_length([0, 2, 4], torch.tensor([5, 7, 11]), torch.arange(8) * 100) == 5 + 7 + 400


def test_walker_long_walks_grow_the_buffer():
    """A 20 000-node chain per strand: one candidate walks all of it (longer than the first buffer guess)."""
    from gnnome_b200.decode import WalkGraph
    n = 40_000
    fwd = np.arange(0, n - 2, 2)                       # 0 -> 2 -> 4 ... on the even strand
    rev = np.arange(n - 1, 1, -2)                      # n-1 -> n-3 -> ... on the odd strand (reverse complements)
    src, dst = np.concatenate((fwd, rev)), np.concatenate((fwd + 2, rev - 2))
    wg = WalkGraph.from_edge_list(src, dst, n)
    logp = torch.zeros(src.size)
    mid = n // 2
    (walk_f, walk_b, _, _), = wg.run_greedy_both_ways([(mid, mid + 2)], logp)
    assert walk_b + walk_f == list(range(0, n, 2))


@pytest.mark.parametrize('run', [0, 1])
def test_get_contigs_greedy_matches_reference_run(golden, run, tmp_path):
    """The whole decoder loop against a run of the reference's own get_contigs_greedy (oracle/make_golden_contigs.py):
    same torch seed -> the same start edges are sampled -> the same contigs, walk for walk."""
    from gnnome_b200.assembly import AssemblyGraph
    from gnnome_b200.decode import get_contigs_greedy
    g = golden('contigs')
    src, dst, n = g['src'], g['dst'], g['num_nodes']
    succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs[u].append(v)
        preds[v].append(u)
        edges[(u, v)] = k
    ag = AssemblyGraph(src, dst, n, dict(score=g['score'], prefix_length=g['prefix_length']), dict(read_length=g['read_length']))
    ref = g['runs'][run]
    torch.manual_seed(ref['seed'])
    walks = get_contigs_greedy(ag, succs, preds, edges, ref['len_threshold'], ref['nb_paths'], checkpoint_dir=str(tmp_path))
    assert walks == ref['walks']
    if len(walks) >= 10:                       # a checkpoint was written after the 10th contig, in the reference's format
        import pickle
        ck = pickle.load(open(tmp_path / 'checkpoint.pkl', 'rb'))
        assert ck['walks'][:10] == ref['walks'][:10] and set(ck) == {'walks', 'visited', 'all_walks_len', 'all_contigs_len'}
        # resuming from it finishes with the same contigs when the remaining draws are replayed... the RNG state is
        # not part of the reference's checkpoint, so only the restored prefix is checked
        resumed = get_contigs_greedy(ag, succs, preds, edges, 10 ** 12, ref['nb_paths'], checkpoint_dir=str(tmp_path),
                                     load_checkpoint=True)
        assert resumed == ck['walks']           # threshold too high for anything new: the restored walks come back


def test_get_contigs_greedy_with_labels(golden):
    from gnnome_b200.assembly import AssemblyGraph
    from gnnome_b200.decode import get_contigs_greedy
    g = golden('contigs')
    n = g['num_nodes']
    src, dst = g['src'][:400], g['dst'][:400]
    succs, preds, edges = {}, {}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs.setdefault(u, []).append(v)
        preds.setdefault(v, []).append(u)
        edges[(u, v)] = k
    y = (torch.arange(400) % 3 != 0).float()
    ag = AssemblyGraph(src, dst, n, dict(y=y, prefix_length=g['prefix_length'][:400]), dict(read_length=g['read_length']))
    torch.manual_seed(1)
    for fast in (False, True):
        walks = get_contigs_greedy(ag, succs, preds, edges, 0, nb_paths=8, use_labels=True, fast_sampling=fast)
        assert walks and all(len(w) >= 2 for w in walks)
        used = [v for w in walks for v in w]
        assert len(used) == len(set(used))          # a node is used by at most one contig
        assert not ({v ^ 1 for v in used} & set(used))   # ... and never together with its reverse complement


def test_inference_driver_from_existing_predictions(golden, tmp_path):
    """gnnome_b200.inference.inference on a dataset directory laid out like the reference's: with {idx}_predicts.pt in
    place (inference.py:429-431) no device is needed; walks are pickled like the reference's and equal a direct decode."""
    import pickle
    from gnnome_b200.assembly import AssemblyGraph
    from gnnome_b200.decode import get_contigs_greedy
    from gnnome_b200.inference import inference
    g = golden('contigs')
    src, dst, n = g['src'], g['dst'], g['num_nodes']
    succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs[u].append(v)
        preds[v].append(u)
        edges[(u, v)] = k
    data, save = tmp_path / 'data', tmp_path / 'out'
    (data / 'hifiasm' / 'processed').mkdir(parents=True)
    (data / 'hifiasm' / 'info').mkdir()
    prefix = g['prefix_length'].clone()
    prefix[:5] = -3                                   # negative prefixes are clamped to 0 before decoding (inference.py:461)
    AssemblyGraph(src, dst, n, dict(prefix_length=prefix, overlap_length=g['prefix_length'],
                                    overlap_similarity=torch.ones(src.numel())), dict(read_length=g['read_length'])
                  ).save(data / 'hifiasm' / 'processed' / '0.pt')
    for name, obj in (('succ', succs), ('pred', preds), ('edges', edges)):
        pickle.dump(obj, open(data / 'hifiasm' / 'info' / f'0_{name}.pkl', 'wb'))
    (save / 'decode').mkdir(parents=True)
    torch.save(g['score'], save / 'decode' / '0_predicts.pt')
    hp = dict(num_decoding_paths=20, len_threshold=60_000, seed=5, load_checkpoint=False)
    out = inference(str(data), 'no-model-needed.pt', 'hifiasm', str(save), hyperparameters=hp)
    walks = pickle.load(open(save / 'decode' / '0_walks.pkl', 'rb'))
    assert out == {0: walks} and len(walks) >= 3
    ag = AssemblyGraph(src, dst, n, dict(score=g['score'], prefix_length=prefix.clamp_min(0)), dict(read_length=g['read_length']))
    torch.manual_seed(5)
    assert walks == get_contigs_greedy(ag, succs, preds, edges, 60_000, 20)
    with pytest.raises(ValueError, match='strategy'):
        inference(str(data), 'x', 'hifiasm', str(save), hyperparameters=dict(strategy='beam'))


def test_walker_vs_oracle_random_graphs_with_ties_and_multi_edges():
    """Property test on small random graphs (self-loops, duplicate pairs, dead ends, exact score ties, random visited
    sets): the C++ walker against the oracle's pure-Python restatement -- walks, float32 sums, contig lengths, jumped-over
    nodes."""
    from hypothesis import given, settings, strategies as st
    from gnnome_b200.decode import WalkGraph

    @settings(max_examples=150, deadline=None)
    @given(st.integers(2, 12), st.integers(0, 60), st.integers(0, 2 ** 31 - 1))
    def check(half_nodes, m, seed):
        n = 2 * half_nodes
        rng = np.random.default_rng(seed)
        src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
        succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
        for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
            succs[u].append(v)
            preds[v].append(u)
            edges[(u, v)] = k
        log_probs = torch.log(torch.from_numpy(rng.choice([0.1, 0.5, 0.5, 0.9], max(m, 1)).astype(np.float32)))
        visited = set(np.nonzero(np.repeat(rng.random(half_nodes) < 0.3, 2))[0].tolist())
        prefix, reads = torch.from_numpy(rng.integers(0, 50, max(m, 1))), torch.from_numpy(rng.integers(1, 90, n))
        wg = WalkGraph.from_edge_list(src, dst, n)
        assert all(np.array_equal(a, b) for a, b in zip(wg._succ, WalkGraph.from_dicts(n, succs, edges, preds)._succ))
        cands = [(int(src[k]), int(dst[k])) for k in range(min(m, 6))]
        for (s, d), (walk_f, walk_b, sum_f, sum_b) in zip(cands, wg.run_greedy_both_ways(cands, log_probs, visited, threads=2)):
            rf, rb, rsf, rsb = R.run_greedy_both_ways(s, d, log_probs, succs, edges, visited)
            assert (walk_f, walk_b) == (rf, rb)
            assert np.float32(sum_f) == rsf.numpy()[0] and np.float32(sum_b) == rsb.numpy()[0]
            walk = walk_b + walk_f
            if all((u, v) in edges for u, v in zip(walk[:-1], walk[1:])):
                assert wg.get_contig_length(walk, prefix, reads) == R.contig_length(walk, edges, prefix, reads)
            assert wg.jumped_nodes(walk) == R.jumped_nodes(walk, succs, preds)

    check()
