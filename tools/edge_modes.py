"""L2 management of the aggregation kernel (gnb_debug_edge_mode): launch time of each setting at a bench size.
  python tools/edge_modes.py [workload] [launches per setting]
bit 1: evict-first TMA loads of e, 2: evict-first TMA stores of e', 4: node rows prefetched when the stage is free,
8: no prefetch.  Every setting is timed twice, interleaved, so that a drifting clock shows."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gnnome_b200
from gnnome_b200 import ops, _lib
from gnnome_b200.layers.encoders import encode_rows2

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
modes = [int(m) for m in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 1, 2, 3, 4, 7, 8, 11]
n, m, H, L, _ = bench.WORKLOADS[wl]
dev = torch.device('cuda', 0)
model = bench.make_model(H, L, dev)
src, dst, x, e = bench.make_inputs(n, m, seed=0)
gi = gnnome_b200.GraphIndex(src, dst, n, dev)
lib = _lib.load()
with torch.no_grad():
    h16, h32 = encode_rows2(x.to(dev), None, model.linear1_node, model.linear2_node, gi.N, want32=True)
    e16, _ = encode_rows2(e.to(dev), gi.in_eid, model.linear1_edge, model.linear2_edge, gi.E)
    ws = {}
    conv = model.gnn.convs[0]
    conv.forward_positions16(gi, h32, h16, e16, ws)
    pk = conv._pack(dev, 'tc2')
    P, Fb, carry, flags = ws['P'], ws['F'], ws['carry'], conv._flags()
    run = lambda: ops.edge_forward_tc2(gi, H, P, pk['We_t'], e16, Fb, carry, flags)
    for _ in range(8):
        run()
    torch.cuda.synchronize()
    res = {md: [] for md in modes}
    for rnd in range(2):
        for md in modes:
            lib.gnb_debug_edge_mode(md)
            run(); torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(reps):
                run()
            ev1.record(); torch.cuda.synchronize()
            res[md].append(ev0.elapsed_time(ev1) / reps)
    lib.gnb_debug_edge_mode(-1)   # back to the library's default
print(f'{wl}: N={n} E={m} H={H}, {reps} launches per setting and round')
for md in modes:
    print(f'  mode {md:2d}: ' + '  '.join(f'{t:7.2f} ms' for t in res[md]))
