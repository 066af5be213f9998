"""Wall time of a whole greedy decode (get_contigs_greedy) on a seeded synthetic graph: gnnome_b200.decode against the
reference's own function (build container only).  Usage: python tests/diag/decode_timing.py [N E nb_paths]"""
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import make_golden_contigs as M  # noqa: E402  (timing baseline only)
from oracle import reference_runner as rr  # noqa: E402
from gnnome_b200.assembly import AssemblyGraph  # noqa: E402
from gnnome_b200.decode import get_contigs_greedy  # noqa: E402


def main():
    n, m, nb_paths = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (20_000, 120_000, 50)
    rec = M.build_inputs(n, m, seed=5)
    succs, preds, edges = M.dicts(rec['src'], rec['dst'], n)
    ag = AssemblyGraph(rec['src'], rec['dst'], n, dict(score=rec['score'], prefix_length=rec['prefix_length']),
                       dict(read_length=rec['read_length']))
    threshold = 70_000
    torch.manual_seed(3)
    t0 = time.perf_counter()
    ours = get_contigs_greedy(ag, succs, preds, edges, threshold, nb_paths)
    t_ours = time.perf_counter() - t0
    print(f'gnnome_b200.decode.get_contigs_greedy: {t_ours:8.2f} s, {len(ours)} contigs, {sum(map(len, ours))} nodes in walks')
    torch.manual_seed(3)
    t0 = time.perf_counter()
    fast = get_contigs_greedy(ag, succs, preds, edges, threshold, nb_paths, fast_sampling=True)
    print(f'   with fast_sampling=True:            {time.perf_counter() - t0:8.2f} s, {len(fast)} contigs, '
          f'{sum(map(len, fast))} nodes in walks (different random draws)')
    if rr.available():
        import math, pickle, types
        from concurrent.futures import ThreadPoolExecutor
        from datetime import datetime
        dgl, _, _ = rr.load()
        ns = dict(torch=torch, dgl=dgl, os=os, pickle=pickle, math=math, datetime=datetime, psutil=None,
                  utils=types.SimpleNamespace(timedelta_to_str=str), ThreadPoolExecutor=ThreadPoolExecutor,
                  RANDOM=False, DEBUG=False, early_stopping=False, p_threshold=0.06)
        rr.load_functions('inference.py', M.NAMES, ns)
        g = dgl.graph((rec['src'], rec['dst']), num_nodes=n)
        g.edata['score'], g.edata['prefix_length'], g.ndata['read_length'] = rec['score'], rec['prefix_length'], rec['read_length']
        torch.manual_seed(3)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ref = ns['get_contigs_greedy'](g, succs, preds, edges, threshold, nb_paths, False, '/tmp', False)
        t_ref = time.perf_counter() - t0
        print(f'reference get_contigs_greedy (dgl shim):  {t_ref:8.2f} s, {len(ref)} contigs; identical walks: {ref == ours}; '
              f'{t_ref / t_ours:.1f} x')


if __name__ == '__main__':
    main()
