#!/usr/bin/env python
"""bench.py -- edges/s of the GatedGCN forward + edge-score pass (BASELINE.json metric).

One "step" = one ``model(graph, x, e)`` pass (encoders -> 8 SymGatedGCN layers -> ScorePredictor,
eval mode) over one synthetic power-law assembly graph of the shape BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|...] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): the graph is partitioned by destination-node
range, every rank runs the same kernels on its shard and exchanges halo rows over NCCL
(gnnome_b200.partition); timing is max-over-ranks, value = global edges / that time.

Keys of the JSON line (rank 0 prints exactly one):
  value      edges/s with the staged graph, x and e already resident in HBM (CUDA events)
  e2e        edges/s through the public model call with PINNED HOST src/dst/x/e: H2D copies, graph
             staging (two device radix sorts), forward, D2H of the (E,1) scores -- all timed
  roofline   the dominant kernel (gnb_edge_forward): its algorithmic bytes / its mean launch time
             measured with CUDA events inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restatement of the reference's algorithm) on a bounded sample of
             the same workload, all host threads, plus the GPU-vs-oracle parity on that sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, E, H, L, description)
    'cfg3': (10_000_000, 60_000_000, 256, 8, 'synthetic power-law assembly graph 10M nodes / 60M edges, hidden=256, L=8'),
    'cfg2': (1_000_000, 6_000_000, 128, 8, 'synthetic power-law assembly graph 1M nodes / 6M edges, hidden=128, L=8'),
    'cfg2s': (1_000_000, 6_000_000, 64, 8, 'synthetic 1M nodes / 6M edges, hidden=64 (shipped-model width), L=8'),
    'mid': (2_000_000, 12_000_000, 256, 8, 'synthetic 2M nodes / 12M edges, hidden=256, L=8 (profiling size)'),
    'small': (100_000, 600_000, 256, 8, 'synthetic 100k nodes / 600k edges, hidden=256, L=8 (debug size)'),
    # training workloads (bench_train.py): one optimisation step = forward + BCE + backward + gradient all-reduce + Adam
    'cfg5': (4_000_000, 24_000_000, 256, 8, 'train.py full forward+backward (BCE on edge scores), synthetic 4M nodes / 24M edges, hidden=256, L=8'),
    'cfg5s': (500_000, 3_000_000, 256, 8, 'training step on the per-GPU share of config 5 at 8 GPUs: 0.5M nodes / 3M edges, hidden=256, L=8'),
    'cfg5t': (50_000, 300_000, 256, 8, 'training step, debug size: 50k nodes / 300k edges, hidden=256, L=8'),
}
TRAIN_WORKLOADS = ('cfg5', 'cfg5s', 'cfg5t')
CPU_TRAIN_SAMPLE = (10_000, 60_000)  # bounded sample for the CPU arm of the training workloads
CPU_SAMPLE = (200_000, 1_200_000)   # bounded sample of the workload for the CPU arm (same H, L, generator): ~17 GB of host RAM
CPU_PASSES = 1                      # timed passes of the CPU arm inside the default GPU run (after one warm-up pass; ~15 s each)
if os.environ.get('GNB_BENCH_CPU_SAMPLE'):   # tests shrink it
    CPU_SAMPLE = tuple(int(v) for v in os.environ['GNB_BENCH_CPU_SAMPLE'].split(','))
HIDDEN_NE, HIDDEN_SCORES = 16, 64   # configs/hyperparameters.py:24-25 of the reference


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The driver reads exactly ONE JSON line from stdout.  Libraries (NCCL prints its version banner on stdout) must
# not get in the way: file descriptor 1 points at stderr for the whole run and the result line goes to the saved
# original descriptor.
_RESULT_FD = None


def _protect_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


STEP_BUDGET_S = float(os.environ.get('GNB_BENCH_STEP_BUDGET_S', 30))   # host-side watchdog: seconds allowed per step


def error_line(args, what):
    """The JSON line of a run that did not finish: the driver learns WHY instead of timing out on silence."""
    hang = ''
    try:
        from gnnome_b200 import _lib
        hang = _lib.hang_report()
    except Exception:   # noqa: BLE001
        pass
    return {'metric': 'edges/s', 'value': None, 'unit': 'edges/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'error': what, 'hang_report': hang or None,
            'config': {'workload': WORKLOADS[args.workload][4]}}


class Watchdog:
    """A device hang must not be a silent 30-minute burn: every phase of the run is armed with a wall-clock budget
    (STEP_BUDGET_S per step); when it expires -- the main thread is then stuck in a synchronize -- rank 0 prints a
    JSON line carrying "error" (and the kernels' own spin-watchdog record, if one fired) and the process exits.
    The device-side watchdog (gnb_hang_report) normally fires first, as a CUDA error that main() reports."""

    def __init__(self, args, rank):
        self.args, self.rank, self.deadline, self.label = args, rank, None, ''
        self.lock = threading.Lock()
        threading.Thread(target=self._run, daemon=True).start()

    def arm(self, label, steps):
        with self.lock:
            self.label, self.deadline = label, time.time() + 60.0 + STEP_BUDGET_S * steps

    def disarm(self):
        with self.lock:
            self.deadline = None

    def _run(self):
        while True:
            time.sleep(0.5)
            with self.lock:
                expired = self.deadline is not None and time.time() > self.deadline
                label = self.label
            if expired:
                log(f'[bench] watchdog: phase {label!r} exceeded its budget')
                if self.rank == 0:
                    emit(error_line(self.args, f'phase {label!r} exceeded its wall-clock budget '
                                               f'({STEP_BUDGET_S:.0f} s per step): device hang'))
                os._exit(3)


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def make_model(H, L, device=None):
    import gnnome_b200
    torch.manual_seed(0)
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, HIDDEN_NE, L, HIDDEN_SCORES, 'batch', dropout=None)
    model.eval()
    return model.to(device) if device is not None else model


def make_inputs(n, m, seed=0):
    from gnnome_b200 import synth
    t0 = time.time()
    src, dst = synth.make_assembly_graph(n, m, seed=seed)
    x, e = synth.make_features(src, dst, n, seed=seed)
    log(f'[bench] generated graph N={n} E={m} in {time.time() - t0:.1f}s')
    return tuple(torch.from_numpy(a) for a in (src, dst, x, e))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], 0.0, set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ln in self.lines:
            f = [t.strip() for t in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': smax, 'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_edge_traffic(workload, world):
    """DRAM bytes of one launch of the aggregation kernel from the committed ncu capture (N=1 only)."""
    if world != 1:
        return None
    try:
        with open(os.path.join(ROOT, 'profiles', 'edge_kernel_dram_traffic.json')) as f:
            return json.load(f)['per_launch_bytes'].get(workload)
    except (OSError, KeyError, ValueError):
        return None


def edge_kernel_bytes(n_nodes, n_edges, H):
    """Algorithmic (compulsory) bytes of ONE gnb_edge_forward launch, fp32 state, int32 indices:
    read e + write e' (2*E*H*4), read src/dst (8E), read the B1h/A2h/B2h node rows once each and
    write F (4*N*H*4).  See DESIGN.md section 4."""
    return n_edges * (2 * H * 4 + 8) + n_nodes * 4 * H * 4


def forward_bytes(n_nodes, n_edges, H, L):
    """B_fwd = (2L+2)(E*H*s_e + N*H*s_h) + (8L+20)E + 8N with fp32 state."""
    return (2 * L + 2) * (n_edges * H * 4 + n_nodes * H * 4) + (8 * L + 20) * n_edges + 8 * n_nodes


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's algorithm) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_run(state_dict, H, L, steps, warmup, sample=CPU_SAMPLE, seed=1):
    from oracle import restatement as R   # allowed here: cpu_baseline / --impl reference legs only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, m = sample
    src, dst, x, e = make_inputs(n, m, seed=seed)
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = R.model_forward(state_dict, src, dst, n, x, e, faithful=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return dict(n=n, m=m, times=times, out=out, inputs=(src, dst, x, e), cores=cores)


def cross_backend_parity(device):
    """Parity at size where the CPU oracle cannot go: the tcgen05 / split-fp16 product path against the CUDA-core
    fp32 kernels (an independent code path: own graph walk, own arithmetic) at BASELINE config 2's FULL size."""
    import gnnome_b200
    n, m, H, L, _ = WORKLOADS['cfg2']
    model = make_model(H, L, device)
    src, dst, x, e = make_inputs(n, m, seed=0)
    with torch.no_grad():
        gi = gnnome_b200.GraphIndex(src, dst, n, device)
        x_d, e_d = x.to(device), e.to(device)
        a = model(gi, x_d, e_d)
        gnnome_b200.set_backend('ffma')
        try:
            b = model(gi, x_d, e_d)
        finally:
            gnnome_b200.set_backend('tc2')
        err = (torch.sigmoid(a.double()) - torch.sigmoid(b.double())).abs().max().item()
    return {'workload': f'cfg2 full size (N={n} E={m} H={H} L={L})', 'backends': 'tc2 (tcgen05, split fp16 state) vs ffma (CUDA cores, fp32)',
            'max_prob_diff': err}


def cfg1_standin_parity(device):
    """BASELINE config 1 (the E. coli example through hifiasm -> DGL -> inference.py) cannot be produced here: hifiasm,
    Biopython and DGL are absent.  Its stand-in, as SURVEY.md section 8(c)(iv) prescribes: an E. coli-sized synthetic
    assembly graph scored with the reference's SHIPPED weights, GPU path vs the CPU oracle."""
    import gnnome_b200
    from oracle import restatement as R   # checker only (cpu_baseline leg)
    sd = torch.load(os.path.join(ROOT, 'tests', 'golden', 'weights.pt'), weights_only=True)
    n, m = 15_000, 100_000
    src, dst, x, e = make_inputs(n, m, seed=7)
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch')
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(device)
    with torch.no_grad():
        out = model((src, dst, n), x.to(device), e.to(device))
        ref = R.model_forward(sd, src, dst, n, x, e, faithful=True)
    err = (torch.sigmoid(out.double().cpu()) - torch.sigmoid(ref.double())).abs().max().item()
    return {'what': f'stand-in for BASELINE config 1: E. coli-sized synthetic assembly graph (N={n} E={m}), the reference\'s '
                    f'shipped weights/weights.pt (H=64, L=8); the real example needs hifiasm + Biopython + DGL',
            'max_prob_err_vs_oracle': err}


def run_reference_arm(args, wl):
    rank, _, world = dist_env()
    if rank != 0:
        return
    n, m, H, L, desc = wl
    model = make_model(H, L)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    r = cpu_oracle_run(sd, H, L, steps=args.steps, warmup=min(args.warmup, 1))
    t = float(np.mean(r['times']))
    value = r['m'] / t
    sample = (f'N={r["n"]} E={r["m"]} (same generator/H={H}/L={L}; the full workload needs ~14 KB/edge of host RAM '
              f'on the reference path)')
    line = {
        'impl': 'reference', 'metric': 'edges/s', 'value': value, 'unit': 'edges/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': t * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'N': n, 'E': m, 'H': H, 'L': L, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'edges/s', 'cores': r['cores'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args, wl):
    import gnnome_b200
    from gnnome_b200 import ops, _lib
    rank, local_rank, world = dist_env()
    n, m, H, L, desc = wl
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torchrun')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    _lib.load()
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)

    dog = Watchdog(args, rank)
    model = make_model(H, L, device)
    src, dst, x, e = make_inputs(n, m, seed=0)

    if world > 1:
        from gnnome_b200 import partition
        runner = partition.ShardedForward(model, src, dst, n, x, e, rank, world, device)
    else:
        runner = None

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement ---------------------------------------------------------
    if runner is None:
        gi = gnnome_b200.GraphIndex(src, dst, n, device)
        x_d, e_d = x.to(device), e.to(device)
        step = lambda: model(gi, x_d, e_d)
    else:
        step = runner.step
    with torch.no_grad():
        dog.arm('warm-up', args.warmup)
        for _ in range(args.warmup):
            out = step()
        barrier()
        dog.arm('timed steps', args.steps)
        ops.LaunchLog.reset(enabled=True, timing=True)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            out = step()
        ev1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
    dog.disarm()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = ops.LaunchLog.total()
    ktimes = ops.LaunchLog.times_ms()
    ops.LaunchLog.reset(enabled=False)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = m / (ms * 1e-3)
    breakdown = {k: {'launches_per_step': len(v) / args.steps, 'ms_per_step': float(np.sum(v)) / args.steps}
                 for k, v in sorted(ktimes.items())}

    # ---- roofline of the dominant kernel -------------------------------------------------------
    peak, peak_src = measured_peak()
    ek_name = next((k for k in ('gnb_edge_forward_tc2', 'gnb_edge_forward_tc', 'gnb_edge_forward') if k in ktimes),
                   'gnb_edge_forward')
    ek = ktimes.get(ek_name, [])
    roofline = None
    if ek:
        n_loc, m_loc = (runner.local_sizes() if runner is not None else (n, m))
        kb = edge_kernel_bytes(n_loc, m_loc, H)
        achieved = kb / (float(np.mean(ek)) * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': ek_name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': measured_edge_traffic(args.workload, world), 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': kb, 'ms_per_launch': float(np.mean(ek)),
                    'whole_forward_frac': forward_bytes(n, m, H, L) / (ms * 1e-3) / 1e9 / peak / world}

    # ---- end-to-end through the public call with pinned host buffers ---------------------------
    e2e = None
    dog.arm('end-to-end steps', 2 * (1 + max(1, min(args.steps, 3))))
    if runner is None:
        hs, hd, hx, he = (t.pin_memory() for t in (src, dst, x, e))
        e2e_steps = max(1, min(args.steps, 3))
        host_out = torch.empty((m, 1), dtype=torch.float32).pin_memory()
        with torch.no_grad():
            def e2e_step():
                o = model((hs.to(device, non_blocking=True), hd.to(device, non_blocking=True), n),
                          hx.to(device, non_blocking=True), he.to(device, non_blocking=True))
                host_out.copy_(o, non_blocking=True)
            e2e_step()
            barrier()
            ev0.record()
            for _ in range(e2e_steps):
                e2e_step()
            ev1.record()
            barrier()
        e2e_ms = ev0.elapsed_time(ev1) / e2e_steps
        e2e = {'value': m / (e2e_ms * 1e-3), 'unit': 'edges/s', 'ms_per_step': e2e_ms, 'steps': e2e_steps,
               'h2d_bytes_per_step': int(sum(t.numel() * t.element_size() for t in (hs, hd, hx, he))),
               'd2h_bytes_per_step': int(host_out.numel() * 4),
               'includes': 'H2D of src/dst/x/e, graph staging (2 radix sorts), forward, D2H of scores'}
    else:
        import torch.distributed as dist
        h2d, d2h = runner.e2e_prepare()
        e2e_steps = max(1, min(args.steps, 3))
        with torch.no_grad():
            runner.e2e_step()
            barrier()
            ev0.record()
            for _ in range(e2e_steps):
                runner.e2e_step()
            ev1.record()
            barrier()
        tt = torch.tensor([ev0.elapsed_time(ev1) / e2e_steps], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        bb = torch.tensor([h2d, d2h], dtype=torch.int64, device=device)
        dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        e2e_ms = float(tt.item())
        e2e = {'value': m / (e2e_ms * 1e-3), 'unit': 'edges/s', 'ms_per_step': e2e_ms, 'steps': e2e_steps,
               'h2d_bytes_per_step': int(bb[0].item()), 'd2h_bytes_per_step': int(bb[1].item()),
               'includes': 'per rank: H2D of its shard (local src/dst, x, e), graph staging, forward with halo '
                           'exchanges, D2H of its scores; max over ranks'}

    # ---- N > 1: the sharded forward against the single-GPU forward of the same model on a sample graph ----------
    parity_multi = None
    if world > 1:
        import torch.distributed as dist
        from gnnome_b200 import partition
        dog.arm('sharded-vs-single parity', 4)
        sn, sm = 200_000, 1_200_000
        s_src, s_dst, s_x, s_e = make_inputs(sn, sm, seed=1)
        with torch.no_grad():
            r2 = partition.ShardedForward(model, s_src, s_dst, sn, s_x, s_e, rank, world, device)
            full = partition.gather_scores(r2, r2.step(), sm)
            if rank == 0:
                single = model((s_src, s_dst, sn), s_x.to(device), s_e.to(device))
                parity_multi = {'sample': f'N={sn} E={sm} H={H} L={L}, same generator',
                                'max_prob_diff': (torch.sigmoid(full.double()) - torch.sigmoid(single.double())).abs().max().item(),
                                'max_logit_diff': (full - single).abs().max().item()}
        del r2, full
    dog.arm('teardown', 2)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    dog.disarm()
    if rank != 0:
        return
    # ---- CPU baseline + parity on the bounded sample (rank 0, N=1 only) ------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        del out
        torch.cuda.empty_cache()
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        r = cpu_oracle_run(sd, H, L, steps=CPU_PASSES, warmup=1)    # ~3 s per pass at H=256: 10-15 s of CPU work
        s_src, s_dst, s_x, s_e = r['inputs']
        dog.arm('parity sample', 2)
        with torch.no_grad():
            ours = model((s_src, s_dst, r['n']), s_x, s_e)
        torch.cuda.synchronize()
        dog.disarm()
        perr = (torch.sigmoid(ours.double().cpu()) - torch.sigmoid(r['out'].double())).abs().max().item()
        t_cpu = float(np.mean(r['times']))
        cpu = {'value': r['m'] / t_cpu, 'unit': 'edges/s', 'cores': r['cores'], 'kind': 'port',
               'sample': f'N={r["n"]} E={r["m"]} H={H} L={L}, same generator, mean of {len(r["times"])} passes after one '
                         f'warm-up ({t_cpu:.1f}s each); torch {torch.__version__} CPU threads={r["cores"]}',
               'parity_max_prob_err_on_sample': perr}
        del r, ours
        dog.arm('parity extras', 8)
        cpu['parity_cross_backend'] = cross_backend_parity(device)
        cpu['cfg1_standin'] = cfg1_standin_parity(device)
        dog.disarm()

    line = {
        'metric': 'edges/s', 'value': value, 'unit': 'edges/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32 (state held as split fp16 hi+lo pairs, 22 significant bits; fp16x2-split tcgen05 products, f32 accumulate)', 'data': 'synthetic',
        'config': {'workload': desc, 'N': n, 'E': m, 'H': H, 'L': L, 'model': 'SymGatedGCNModel eval, seed-0 init',
                   'graph': 'make_assembly_graph(seed=0, band=64, alpha=2.2, p_long=0.01)',
                   'l2': 'inputs (>= 1 GB of edge state per layer) exceed the 126 MB L2; no flush needed',
                   'parallelism': f'dst-range x{world}' if world > 1 else 'single GPU'},
        'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
        'kernels': breakdown,
    }
    if parity_multi is not None:
        line['parity_vs_single_gpu'] = parity_multi
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='gnnome_b200', choices=['gnnome_b200', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    _protect_stdout()
    wl = WORKLOADS[args.workload]
    train = args.workload in TRAIN_WORKLOADS
    if train:
        import bench_train
    if args.impl == 'reference':
        bench_train.run_reference_arm(args, wl, emit) if train else run_reference_arm(args, wl)
        return
    try:
        bench_train.run_gpu_arm(args, wl, emit) if train else run_gpu_arm(args, wl)
    except Exception as exc:   # noqa: BLE001 -- a CUDA error (e.g. the kernels' spin watchdog trapped): say so in the line
        import traceback
        traceback.print_exc()
        if dist_env()[0] == 0:
            emit(error_line(args, f'{type(exc).__name__}: {exc}'[:2000]))
        os._exit(3)


if __name__ == '__main__':
    main()
