"""Device time of the hand-off kernels at the cfg3 graph size (10 M nodes / 60 M edges): degree rows, z-scores, induced
subgraph.  Usage: python tools/handoff_timing.py [N E]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnome_b200  # noqa: E402
from gnnome_b200 import ops, synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def main():
    n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10_000_000, 60_000_000)
    src, dst = synth.make_assembly_graph(n, m, seed=0)
    gi = gnnome_b200.GraphIndex(torch.from_numpy(src), torch.from_numpy(dst), n)
    e = torch.randn(m, 2, device='cuda')
    keep = torch.rand(n // 2, device='cuda').lt(0.85).repeat_interleave(2)
    ms, _ = timed(lambda: ops.degree_rows(gi))
    print(f'gnb_degree_rows      N={n}: {ms:8.3f} ms  {(16 * n) / ms / 1e6:7.1f} GB/s (8 B read + 8 B written per node)')
    ms, _ = timed(lambda: ops.zscore_cols(e, 0b01))
    print(f'gnb_zscore_cols      E={m}: {ms:8.3f} ms  {(32 * m) / ms / 1e6:7.1f} GB/s (3 reads + 1 write of 8 B per edge)')
    ms, out = timed(lambda: ops.node_subgraph(keep, gi.src, gi.dst, n), reps=3)
    print(f'node_subgraph        E={m}: {ms:8.3f} ms  kept {out[0].numel()} nodes / {out[1].numel()} edges '
          f'({(16 * m + 12 * out[1].numel()) / ms / 1e6:7.1f} GB/s on 2 x 8 B read per edge + 12 B per induced edge; '
          f'includes the host read of the sizes and the output allocation)')


if __name__ == '__main__':
    main()
