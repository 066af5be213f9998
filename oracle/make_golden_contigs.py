"""Generate ``tests/golden/contigs.pt``: a whole run of the reference's OWN unmodified ``get_contigs_greedy``
(``inference.py:167-361`` with ``get_subgraph``, ``sample_edges``, ``get_contig_length``, the walk functions) over
``oracle/dgl_shim`` on a seeded synthetic assembly graph.

TEST INFRASTRUCTURE.  Run in the build container only:  ``python -m oracle.make_golden_contigs``.
The graph is simple (duplicate node pairs removed): ``graph.edges[u, v]`` resolves a pair to ONE edge id and DGL does not
say which when there are several.
"""
import contextlib
import io
import math
import os
import pickle
import sys
import types
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_runner as rr  # noqa: E402
from gnnome_b200 import synth  # noqa: E402

NAMES = {'get_contig_length', 'get_subgraph', 'sample_edges', 'greedy_forwards', 'greedy_backwards_rc',
         'run_greedy_both_ways', 'get_contigs_greedy'}


def build_inputs(n, m, seed):
    src, dst = synth.make_assembly_graph(n, m, seed=seed)
    _, first = np.unique(src.astype(np.int64) * n + dst, return_index=True)
    keep = np.sort(first)
    src, dst = src[keep], dst[keep]
    rng = np.random.default_rng(seed)
    return dict(src=torch.from_numpy(src), dst=torch.from_numpy(dst), num_nodes=n,
                score=torch.from_numpy(rng.normal(0.5, 3.0, src.size).astype(np.float32)),
                prefix_length=torch.from_numpy(rng.integers(100, 9000, src.size)),
                read_length=torch.from_numpy(rng.integers(8000, 25000, n)))


def dicts(src, dst, n):
    succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs[u].append(v)
        preds[v].append(u)
        edges[(u, v)] = k
    return succs, preds, edges


def main():
    dgl, _, _ = rr.load()
    utils = types.SimpleNamespace(timedelta_to_str=lambda d: str(d))
    ns = dict(torch=torch, dgl=dgl, os=os, pickle=pickle, math=math, datetime=datetime, utils=utils, psutil=None,
              ThreadPoolExecutor=ThreadPoolExecutor, RANDOM=False, DEBUG=False, early_stopping=False, p_threshold=0.06)
    rr.load_functions('inference.py', NAMES, ns)
    rec = build_inputs(1600, 9600, seed=53)
    n = rec['num_nodes']
    succs, preds, edges = dicts(rec['src'], rec['dst'], n)
    g = dgl.graph((rec['src'], rec['dst']), num_nodes=n)
    g.edata['score'], g.edata['prefix_length'], g.ndata['read_length'] = rec['score'], rec['prefix_length'], rec['read_length']
    runs = []
    for seed, nb_paths, len_threshold in ((0, 20, 60_000), (7, 50, 150_000)):
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            walks = ns['get_contigs_greedy'](g, succs, preds, edges, len_threshold, nb_paths, False, '/tmp', False)
        runs.append(dict(seed=seed, nb_paths=nb_paths, len_threshold=len_threshold, walks=walks))
        print(f'seed {seed}: {len(walks)} contigs, walk lengths {[len(w) for w in walks][:12]}')
    rec['runs'] = runs
    torch.save(rec, os.path.join(ROOT, 'tests', 'golden', 'contigs.pt'))


if __name__ == '__main__':
    main()
