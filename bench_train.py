"""bench.py's training workloads (BASELINE.json config 5: ``train.py`` full forward + backward, BCE on the edge scores,
synthetic 4M nodes / 24M edges, hidden 256, L = 8, 8 x B200).  Imported by bench.py; not a script of its own.

One "step" = one optimisation step of ``SymGatedGCNModel`` in train mode on the whole graph: forward (batch-statistics
BatchNorm over all E / N rows), ``binary_cross_entropy_with_logits(pos_weight)`` (train.py:143-144), backward, gradient
all-reduce, Adam update (train.py:259) -- ``gnnome_b200.train_dist.ShardedTrainer``, the graph partitioned by
destination-node range over the ranks.  ``value`` = global edges / max-over-ranks step time.

Algorithmic bytes of a training step (the ``roofline`` object, whole step): SURVEY.md section 8(d) counts
``B_train = 3 * B_fwd`` -- the forward's stage-compulsory traffic, once more to read the saved activations and once more
to write the gradients, with ``B_fwd = (2L+2)(E*H*4 + N*H*4) + (8L+20)E + 8N`` (fp32 state).
"""
import os
import time

import numpy as np
import torch

LABEL_P, POS_WEIGHT, LR = 0.75, 1.0 / 3.0, 1e-4     # SURVEY.md section 8(d); configs/hyperparameters.py (lr)


def make_labels(m, seed=0):
    return (torch.rand(m, generator=torch.Generator().manual_seed(seed + 17)) < LABEL_P).float()


def cpu_oracle_train(state_dict, n, m, seed, steps, warmup):
    """The reference's training step (forward in train mode, BCE, backward) restated on the host cores."""
    import __main__ as bench
    from oracle import restatement as R   # allowed here: cpu_baseline / --impl reference legs only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    src, dst, x, e = bench.make_inputs(n, m, seed=seed)
    y = make_labels(m, seed)
    times, loss = [], None
    for i in range(warmup + steps):
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
             for k, v in state_dict.items()}
        t0 = time.perf_counter()
        out = R.model_forward(p, src, dst, n, x, e, training=True, cast=False).squeeze(-1)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y, pos_weight=torch.tensor(POS_WEIGHT))
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return dict(n=n, m=m, times=times, loss=float(loss.detach()), cores=cores, inputs=(src, dst, x, e, y))


def run_reference_arm(args, wl, emit):
    import __main__ as bench
    rank, _, _ = bench.dist_env()
    if rank != 0:
        return
    n, m, H, L, desc = wl
    model = bench.make_model(H, L)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sn, sm = bench.CPU_TRAIN_SAMPLE
    r = cpu_oracle_train(sd, sn, sm, 1, steps=args.steps, warmup=min(args.warmup, 1))
    t = float(np.mean(r['times']))
    value = sm / t
    sample = f'N={sn} E={sm} (same generator/H={H}/L={L}): one forward + BCE + backward of the restated reference'
    emit({'impl': 'reference', 'metric': 'edges/s (training step: fwd + bwd)', 'value': value, 'unit': 'edges/s',
          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': t * 1e3,
          'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': desc, 'N': n, 'E': m, 'H': H, 'L': L, 'sample': sample},
          'cpu_baseline': {'value': value, 'unit': 'edges/s', 'cores': r['cores'], 'kind': 'port', 'sample': sample},
          'e2e': {'value': value, 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0})


def run_gpu_arm(args, wl, emit):
    import __main__ as bench
    from gnnome_b200 import _lib, ops, train_dist
    rank, local_rank, world = bench.dist_env()
    n, m, H, L, desc = wl
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torchrun')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback)')
    _lib.load()
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    # activations of one check-pointed layer + the per-layer (h, e) inputs: ~20 E x H fp32 tensors per rank
    need_gb = 20 * (m / world) * H * 4 / 1e9
    have_gb = torch.cuda.get_device_properties(device).total_memory / 1e9
    if need_gb > 0.9 * have_gb:
        raise SystemExit(f'{args.workload} on {world} GPU(s) needs ~{need_gb:.0f} GB per GPU for activations '
                         f'(have {have_gb:.0f} GB): run it on more GPUs (BASELINE config 5 is an 8-GPU configuration)')
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    dog = bench.Watchdog(args, rank)
    model = bench.make_model(H, L, device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=LR)
    src, dst, x, e = bench.make_inputs(n, m, seed=0)
    y = make_labels(m, 0)
    tr = train_dist.ShardedTrainer(model, src, dst, n, x, e, y, rank, world, device, pos_weight=POS_WEIGHT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dog.arm('warm-up', 2 * args.warmup)
    losses = []
    for _ in range(args.warmup):
        losses.append(tr.step(opt))
    barrier()
    dog.arm('timed steps', 2 * args.steps)
    ops.LaunchLog.reset(enabled=True, timing=True)
    sampler = bench.ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        losses.append(tr.step(opt))
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    dog.disarm()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches, ktimes = ops.LaunchLog.total(), ops.LaunchLog.times_ms()
    ops.LaunchLog.reset(enabled=False)
    peak_mem = torch.cuda.max_memory_allocated(device) / 1e9
    if world > 1:
        t = torch.tensor([ms, peak_mem], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, peak_mem = float(t[0]), float(t[1])
    value = m / (ms * 1e-3)
    breakdown = {k: {'launches_per_step': len(v) / args.steps, 'ms_per_step': float(np.sum(v)) / args.steps}
                 for k, v in sorted(ktimes.items())}

    # ---- end to end: the shard's inputs come from pinned host memory every step, the loss goes back -------------
    host = {k: getattr(tr, k).cpu().pin_memory() for k in ('x_own', 'e_pos', 'y_pos')}
    e2e_steps = max(1, min(args.steps, 3))
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def e2e_step():
        for k, v in host.items():
            getattr(tr, k).copy_(v, non_blocking=True)
        loss_host.copy_(tr.step(opt), non_blocking=True)

    dog.arm('end-to-end steps', 2 * (1 + e2e_steps))
    e2e_step()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    barrier()
    dog.disarm()
    e2e_ms = ev0.elapsed_time(ev1) / e2e_steps
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    if world > 1:
        tt = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        bb = torch.tensor([h2d], dtype=torch.int64, device=device)
        dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        e2e_ms, h2d = float(tt.item()), int(bb.item())
    e2e = {'value': m / (e2e_ms * 1e-3), 'unit': 'edges/s', 'ms_per_step': e2e_ms, 'steps': e2e_steps,
           'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 4 * world,
           'includes': 'per rank: H2D of its shard (x, e, labels) from pinned memory, forward, loss, backward, gradient '
                       'all-reduce, Adam step, D2H of the loss; max over ranks'}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = bench.measured_peak()
    b_train = 3 * bench.forward_bytes(n, m, H, L)
    achieved = b_train / (ms * 1e-3) / 1e9 / world
    top = max(breakdown.items(), key=lambda kv: kv[1]['ms_per_step'])[0] if breakdown else None
    roofline = {'bound': 'hbm', 'kernel': 'whole training step (per GPU)', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                'algorithmic_bytes_per_step': b_train, 'formula': 'B_train = 3 * B_fwd; B_fwd = (2L+2)(4EH + 4NH) + (8L+20)E + 8N',
                'top_kernel_by_time': top}
    cpu = None
    if world == 1 and not args.no_cpu:
        sd = {k: v.detach().cpu().clone() for k, v in bench.make_model(H, L).state_dict().items()}
        sn, sm = bench.CPU_TRAIN_SAMPLE
        r = cpu_oracle_train(sd, sn, sm, 1, steps=2, warmup=1)
        # parity on the sample: the first training step's loss of a fresh model, GPU (sharded trainer, world 1) vs oracle
        s_src, s_dst, s_x, s_e, s_y = r['inputs']
        fresh = bench.make_model(H, L, device)
        fresh.train()
        tr2 = train_dist.ShardedTrainer(fresh, s_src, s_dst, sn, s_x, s_e, s_y, 0, 1, device, pos_weight=POS_WEIGHT)
        gl = float(tr2.step(None))
        t_cpu = float(np.mean(r['times']))
        cpu = {'value': sm / t_cpu, 'unit': 'edges/s', 'cores': r['cores'], 'kind': 'port',
               'sample': f'N={sn} E={sm} H={H} L={L}: forward + BCE + backward of the restated reference, mean of '
                         f'{len(r["times"])} steps after one warm-up ({t_cpu:.1f}s each)',
               'parity_first_step_loss': {'gpu': gl, 'oracle': r['loss'], 'abs_diff': abs(gl - r['loss'])}}
    emit({'metric': 'edges/s (training step: fwd + bwd)', 'value': value, 'unit': 'edges/s', 'n_gpus': world,
          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': desc, 'N': n, 'E': m, 'H': H, 'L': L,
                     'model': 'SymGatedGCNModel train mode (batch-statistics BatchNorm), seed-0 init, Adam lr 1e-4',
                     'labels': f'Bernoulli({LABEL_P}), pos_weight {POS_WEIGHT:.4f}',
                     'graph': 'make_assembly_graph(seed=0, band=64, alpha=2.2, p_long=0.01)',
                     'l2': 'per-layer edge state (>= 0.4 GB per rank) exceeds the 126 MB L2; no flush needed',
                     'parallelism': f'dst-range x{world}, layer check-pointing'},
          'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
          'kernels': breakdown, 'loss_first_last': [float(losses[0]), float(losses[-1])],
          'peak_memory_gb_per_gpu': peak_mem})
