"""Bring-up check of the tcgen05 linear kernel against fp64 (GPU box).  Run under `timeout`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnome_b200 import ops

def main():
    torch.manual_seed(0)
    for rows, K, M in ((64, 32, 128), (64, 64, 128), (200, 128, 128), (1000, 256, 128), (5000, 256, 1280), (777, 64, 320), (100000, 128, 640)):
        a = torch.randn(rows, K) * 3
        w = torch.randn(M, K) / K ** 0.5
        b = torch.randn(M)
        ref = a.double() @ w.double().t() + b.double()
        wp = ops.pack_linear_tc(w.cuda())
        out = ops.node_linear_tc(a.cuda(), wp, b.cuda(), M)
        torch.cuda.synchronize()
        err = (out.cpu().double() - ref).abs().max().item()
        ref32 = ops.node_linear(a.cuda(), w.t().contiguous().cuda(), b.cuda())
        err32 = (ref32.cpu().double() - ref).abs().max().item()
        print(f'rows={rows} K={K} M={M}: tc err {err:.3g}  ffma err {err32:.3g}  max|ref| {ref.abs().max():.3g}', flush=True)
        if err > 1e-3:
            d = (out.cpu().double() - ref).abs()
            bad = (d > 1e-3).nonzero()
            print('  bad entries', bad.shape[0], 'first', bad[:8].tolist(), 'vals', out.cpu()[bad[0, 0], bad[0, 1]].item(), ref[bad[0, 0], bad[0, 1]].item())

if __name__ == '__main__':
    main()
