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
# ---- training step: sharded trainer (all ranks) against the single-GPU training path (rank 0) ------------------
import copy
import torch.nn.functional as F
from gnnome_b200 import train_dist
for H, L, n, m, p_long in [(64, 3, 20000, 120000, 0.05), (256, 2, 30000, 180000, 0.2)]:
    s, d = synth.make_assembly_graph(n, m, seed=H + 1, p_long=p_long)
    x, e = synth.make_features(s, d, n, seed=H + 1)
    s, d, x, e = map(torch.from_numpy, (s, d, x, e))
    y = (torch.rand(m, generator=torch.Generator().manual_seed(4)) < 0.75).float()
    torch.manual_seed(1)
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').to(dev).train()
    ref = copy.deepcopy(model)
    tr = train_dist.ShardedTrainer(model, s, d, n, x, e, y, rank, world, dev, pos_weight=1 / 3)
    loss = tr.step(None)
    if rank == 0:
        out = ref((s, d, n), x.to(dev), e.to(dev)).squeeze(-1)
        ref_loss = F.binary_cross_entropy_with_logits(out, y.to(dev), pos_weight=torch.tensor(1 / 3, device=dev))
        ref_loss.backward()
        gerr = max(float((p_.grad - q_.grad).abs().max() / max(float(q_.grad.abs().max()), 1e-6))
                   for p_, q_ in zip(model.parameters(), ref.parameters()))
        berr = max(float((a - b).abs().max()) for a, b in zip(model.buffers(), ref.buffers()) if a.is_floating_point())
        print(f'[dist x{world}] train H={H} L={L} E={m}: loss {float(loss):.7f} vs 1 GPU {float(ref_loss):.7f}; '
              f'max relative gradient diff {gerr:.3g}; max BatchNorm buffer diff {berr:.3g}', flush=True)
        ok = ok and abs(float(loss) - float(ref_loss)) < 1e-5 and gerr < 1e-3 and berr < 1e-4
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print('DIST_OK' if ok else 'DIST_FAIL', flush=True)
