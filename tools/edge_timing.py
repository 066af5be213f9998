"""Cycle accounting of the gnb_edge_forward_tc2 epilogue warps (gnb_debug_edge_timing) on one layer."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnome_b200
from gnnome_b200 import ops, synth, _lib
from gnnome_b200.layers.encoders import encode_rows2

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n, m = (2_000_000, 12_000_000)
s, d = synth.make_assembly_graph(n, m, seed=0)
x, e = synth.make_features(s, d, n, seed=0)
s, d, x, e = map(torch.from_numpy, (s, d, x, e))
torch.manual_seed(0)
model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, 2, 64, 'batch').cuda().eval()
gi = gnnome_b200.GraphIndex(s, d, n)
lib = _lib.load()
with torch.no_grad():
    h16, h32 = encode_rows2(x.cuda(), None, model.linear1_node, model.linear2_node, gi.N, want32=True)
    e16, _ = encode_rows2(e.cuda(), gi.in_eid, model.linear1_edge, model.linear2_edge, gi.E)
    ws = {}
    conv = model.gnn.convs[0]
    conv.forward_positions16(gi, h32, h16, e16, ws)      # warm-up
    torch.cuda.synchronize()
    buf = torch.zeros((148, 32, 5), dtype=torch.int64, device='cuda')
    lib.gnb_debug_edge_timing(ctypes.c_void_p(buf.data_ptr()))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pk = conv._pack(e16.device)
    ev0.record()
    ops.edge_forward_tc2(gi, H, ws['P'], pk['We_t'], e16, ws['F'], ws['carry'], conv._flags())
    ev1.record()
    torch.cuda.synchronize()
    lib.gnb_debug_edge_timing(None)
ms = ev0.elapsed_time(ev1)
b = buf.cpu().double()
role = b[:, :4, :].clone()   # warp 0: producer [empty wait, -, -, -, tiles]; warp 1: MMA issue [full wait, dempty wait, issue, -, tiles]
b[:, :4, :] = 0
live = b[:, :, 4] > 0
tiles = b[:, :, 4][live]
print(f'H={H} E={m}: kernel {ms:.2f} ms; epilogue warps with work: {int(live.sum())}; tiles per warp {tiles.mean():.1f}')
names = ['wait full (TMA + idx)', 'wait dfull (MMA)', 'gather + compute', 'flush + hand-off']
tot = sum(b[:, :, k][live].sum() for k in range(4))
for k in range(4):
    per_tile = (b[:, :, k][live] / tiles).mean()
    print(f'  {names[k]:24s} {per_tile:9.0f} cycles / tile / warp   ({100 * b[:, :, k][live].sum() / tot:.1f} %)')
print(f'  total {sum((b[:, :, k][live] / tiles).mean() for k in range(4)):.0f} cycles per tile per warp; kernel cycles/tile-visit at 1.9 GHz: {ms * 1e-3 * 1.9e9 / tiles.mean():.0f}')
for w in (4, 5, 8, 12, 16):
    print(f'  CTA 0 warp {w}:', [int(v) for v in (b[0, w, :4] / max(b[0, w, 4], 1)).tolist()], 'tiles', int(b[0, w, 4]))
pt, mt = role[:, 0, 4].clamp(min=1), role[:, 1, 4].clamp(min=1)
print(f'  producer warp: {(role[:, 0, 0] / pt).mean():.0f} cycles / tile waiting for an empty stage ({pt.mean():.0f} tiles per CTA)')
for w in (1, 3):
    mt = role[:, w, 4].clamp(min=1)
    print(f'  MMA warp {w}: per tile {(role[:, w, 0] / mt).mean():.0f} cycles waiting for the stage (TMA), {(role[:, w, 1] / mt).mean():.0f} for the accumulator set, '
          f'{(role[:, w, 2] / mt).mean():.0f} issuing ({mt.mean():.0f} tiles)')
