"""Multi-GPU equivalence check (run under torchrun, one rank per GPU): the destination-range sharded forward
must reproduce the single-GPU forward of the same model/graph.  Prints one line per case on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnnome_b200
from gnnome_b200 import partition, synth

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
ok = True
for H, L, n, m, p_long in [(64, 8, 20000, 120000, 0.01), (256, 3, 30000, 180000, 0.2), (128, 2, 5000, 30000, 1.0)]:
    s, d = synth.make_assembly_graph(n, m, seed=H, p_long=p_long)
    x, e = synth.make_features(s, d, n, seed=H)
    s, d, x, e = map(torch.from_numpy, (s, d, x, e))
    torch.manual_seed(0)
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').to(dev).eval()
    with torch.no_grad():
        runner = partition.ShardedForward(model, s, d, n, x, e, rank, world, dev)
        full = partition.gather_scores(runner, runner.step(), m)
        if rank == 0:
            ref = model((s, d, n), x.to(dev), e.to(dev))
            err = (torch.sigmoid(full.double()) - torch.sigmoid(ref.double())).abs().max().item()
            lerr = (full - ref).abs().max().item()
            print(f'[dist x{world}] H={H} L={L} E={m} p_long={p_long}: halo {runner.shard.n_halo} of {runner.shard.n_own} own; '
                  f'max prob diff vs 1 GPU {err:.3g}, logit diff {lerr:.3g}', flush=True)
            ok = ok and err < 1e-5
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print('DIST_OK' if ok else 'DIST_FAIL', flush=True)
