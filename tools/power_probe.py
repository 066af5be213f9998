"""Which of the layer's kernels run into the board's power cap?  Each kernel of one layer is launched back to back for
a few seconds on a bench-sized graph while NVML is sampled every 20 ms: mean launch time, SM clock and board power.
A kernel whose clock sits below the maximum with `sw_power_cap` active is bound by energy per edge, not by cycles.
  python tools/power_probe.py [workload] [seconds per kernel]"""
import os, sys, threading, time, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import gnnome_b200
from gnnome_b200 import ops
from gnnome_b200.layers.encoders import encode_rows2

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
only = sys.argv[3].split(',') if len(sys.argv) > 3 else None   # substrings of the kernel names to probe
n, m, H, L, _ = bench.WORKLOADS[wl]
dev = torch.device('cuda', 0)
model = bench.make_model(H, L, dev)
src, dst, x, e = bench.make_inputs(n, m, seed=0)
gi = gnnome_b200.GraphIndex(src, dst, n, dev)

import pynvml
pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], threading.Event()


def sampler():
    while not stop.is_set():
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(nv) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(nv)))
        time.sleep(0.02)


with torch.no_grad():
    h16, h32 = encode_rows2(x.to(dev), None, model.linear1_node, model.linear2_node, gi.N, want32=True)
    e16, _ = encode_rows2(e.to(dev), gi.in_eid, model.linear1_edge, model.linear2_edge, gi.E)
    ws = {}
    conv = model.gnn.convs[0]
    conv.forward_positions16(gi, h32, h16, e16, ws)      # fills P, F, carry
    torch.cuda.synchronize()
    pk = conv._pack(dev, 'tc2')
    P, Fb, carry = ws['P'], ws['F'], ws['carry']
    h_out, h16_out = torch.empty_like(h32), torch.empty_like(h16)
    flags = conv._flags()
    kernels = {
        'gnb_edge_forward_tc2': lambda: ops.edge_forward_tc2(gi, H, P, pk['We_t'], e16, Fb, carry, flags),
        'gnb_node_update2': lambda: ops.node_update2(gi, H, P, e16, Fb, carry, h32, pk['scale_h'], pk['shift_h'], h_out,
                                                     h16_out, flags, gi.chunk(H, 'tc2')),
        'gnb_node_linear_tc2': lambda: ops.node_linear_tc2(h16, pk['Wn_t'], pk['bn'], P.shape[1], out=P),
    }
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    print(f'{wl}: N={n} E={m} H={H}; max SM clock {pynvml.nvmlDeviceGetMaxClockInfo(nv, pynvml.NVML_CLOCK_SM)} MHz, '
          f'power limit {pynvml.nvmlDeviceGetPowerManagementLimit(nv) / 1000:.0f} W')
    for name, fn in kernels.items():
        if only and not any(o in name for o in only):
            continue
        fn(); torch.cuda.synchronize()
        time.sleep(1.0)                                   # let the board cool down to idle power
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        t0 = time.time(); launches = 0
        ev[0].record()
        while time.time() - t0 < secs:
            for _ in range(4):
                fn(); launches += 1
            torch.cuda.synchronize()
        ev[1].record(); torch.cuda.synchronize()
        t1 = time.time()
        win = [s for s in samples if t0 + 0.5 <= s[0] <= t1]   # skip the ramp
        clk = statistics.median(s[1] for s in win); pw = statistics.mean(s[2] for s in win)
        capped = sum(1 for s in win if s[3] & pynvml.nvmlClocksThrottleReasonSwPowerCap) / max(len(win), 1)
        ms = ev[0].elapsed_time(ev[1]) / launches
        print(f'  {name:24s} {ms:8.2f} ms/launch  SM {clk:6.0f} MHz  {pw:6.0f} W  sw_power_cap in {100 * capped:.0f} % of samples '
              f'({ms * clk * 1e3 / 1e6:.1f} M cycles)')
    stop.set()
