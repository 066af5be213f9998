"""Time of one decoder iteration's candidate walks (100 start edges, nothing visited yet): the C++ walker against the
reference's own Python functions (build container only: they are read from /root/reference through
oracle/reference_runner.py) and against the oracle restatement.  Usage: python tests/diag/walk_timing.py [N E]"""
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnnome_b200 import synth  # noqa: E402
from gnnome_b200.decode import WalkGraph  # noqa: E402
from oracle import reference_runner as rr  # noqa: E402  (timing baseline only)


def main():
    n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200_000, 1_200_000)
    src, dst = synth.make_assembly_graph(n, m, seed=7)
    rng = np.random.default_rng(7)
    log_probs = torch.log(torch.sigmoid(torch.from_numpy(rng.normal(0, 3, m).astype(np.float32))))
    cands = [(int(src[k]), int(dst[k])) for k in rng.choice(m, 100, replace=False)]
    t0 = time.perf_counter()
    wg = WalkGraph.from_edge_list(src, dst, n)
    t_build = time.perf_counter() - t0
    for threads in (1, 0):
        t0 = time.perf_counter()
        res = wg.run_greedy_both_ways(cands, log_probs, None, threads=threads)
        dt = time.perf_counter() - t0
        steps = sum(len(a) + len(b) for a, b, _, _ in res)
        print(f'C++ walker, threads={threads or os.cpu_count()}: {dt * 1e3:9.2f} ms for {len(cands)} candidates, {steps} walk steps '
              f'({steps / dt / 1e6:.2f} M steps/s); CSR build {t_build * 1e3:.0f} ms')
    if rr.available():
        ns = dict(torch=torch, math=math, RANDOM=False, early_stopping=False, p_threshold=0.06, DEBUG=False)
        rr.load_functions('inference.py', {'greedy_forwards', 'greedy_backwards_rc', 'run_greedy_both_ways'}, ns)
        succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
        for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
            succs[u].append(v)
            preds[v].append(u)
            edges[(u, v)] = k
        t0 = time.perf_counter()
        ref = [ns['run_greedy_both_ways'](s, d, log_probs, succs, preds, edges, set()) for s, d in cands]
        dt = time.perf_counter() - t0
        steps_ref = sum(len(r[0]) + len(r[1]) for r in ref)
        same = all(r[0] == a and r[1] == b for r, (a, b, _, _) in zip(ref, res))
        print(f'reference Python functions: {dt * 1e3:9.2f} ms, {steps_ref} walk steps ({steps_ref / dt / 1e6:.3f} M steps/s); '
              f'walks identical: {same}')


if __name__ == '__main__':
    main()
