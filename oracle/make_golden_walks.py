"""Generate ``tests/golden/walks.pt``: greedy decoder walks from the reference's OWN unmodified functions
(``inference.py``: ``greedy_forwards``, ``greedy_backwards_rc``, ``run_greedy_both_ways``), executed through
``reference_runner.load_functions`` (the module itself needs DGL, Biopython ...).

TEST INFRASTRUCTURE.  Run in the build container only:  ``python -m oracle.make_golden_walks``.
``get_contig_length`` (:30-37) indexes a DGLGraph by node pairs and the jumped-over nodes (:316-322) are inline code of
``get_contigs_greedy``: both are stored from the restatement in ``oracle/restatement.py``.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_runner as rr  # noqa: E402
from oracle import restatement as R  # noqa: E402
from gnnome_b200 import synth  # noqa: E402


def main():
    ns = dict(torch=torch, math=math, RANDOM=False, early_stopping=False, p_threshold=0.06, DEBUG=False)   # inference.py:25-28
    rr.load_functions('inference.py', {'greedy_forwards', 'greedy_backwards_rc', 'run_greedy_both_ways'}, ns)
    n, m = 3000, 18000
    src, dst = synth.make_assembly_graph(n, m, seed=41)
    rng = np.random.default_rng(41)
    succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs[u].append(v)
        preds[v].append(u)
        edges[(u, v)] = k
    scores = torch.from_numpy(rng.normal(0.0, 3.0, m).astype(np.float32))
    log_probs = torch.log(torch.sigmoid(scores))                                       # inference.py:184
    pairs = rng.random(n // 2) < 0.25
    visited = set(np.nonzero(np.repeat(pairs, 2))[0].tolist())                         # whole strand pairs, as the decoder adds them
    cand = rng.choice(m, size=64, replace=False)
    cands = [(int(src[k]), int(dst[k])) for k in cand]
    prefix_length = torch.from_numpy(rng.integers(100, 9000, m))
    read_length = torch.from_numpy(rng.integers(8000, 25000, n))
    def run(vis):
        out = []
        for s, d in cands:
            walk_f, walk_b, vis_f, vis_b, sum_f, sum_b = ns['run_greedy_both_ways'](s, d, log_probs, succs, preds, edges, vis)
            walk = walk_b + walk_f
            out.append(dict(walk_f=walk_f, walk_b=walk_b, sum_f=sum_f.clone(), sum_b=sum_b.clone(),
                            n_visited=len(vis_f | vis_b), contig_length=R.contig_length(walk, edges, prefix_length, read_length),
                            jumped=sorted(R.jumped_nodes(walk, succs, preds))))
        lens = [len(o['walk_f']) + len(o['walk_b']) for o in out]
        print('candidates', len(out), 'walk lengths min/mean/max', min(lens), sum(lens) / len(lens), max(lens))
        return out

    torch.save(dict(src=torch.from_numpy(src), dst=torch.from_numpy(dst), num_nodes=n, scores=scores, visited=sorted(visited),
                    candidates=cands, prefix_length=prefix_length, read_length=read_length,
                    results=run(visited), results_nothing_visited=run(set())),
               os.path.join(ROOT, 'tests', 'golden', 'walks.pt'))


if __name__ == '__main__':
    main()
