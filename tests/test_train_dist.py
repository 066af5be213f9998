"""Sharded training step (BASELINE config 5) on CPU: world-2/3 ``gloo`` processes against the oracle's autograd.

The arithmetic is a plain-torch implementation of the primitive contracts (tests/_emul.py); what is under test is the
distributed logic of ``gnnome_b200.train_dist``: partition + halo exchanges and their adjoints, all-reduced BatchNorm
statistics (forward and backward), the double running-statistics update of ``bn_e``, loss scaling, the gradient
all-reduce and the layer checkpoint."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gnnome_b200 import synth  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _inputs(n, m, seed, p_long):
    src, dst = synth.make_assembly_graph(n, m, seed=seed, p_long=p_long)
    x, e = synth.make_features(src, dst, n, seed=seed)
    src, dst, x, e = map(torch.from_numpy, (src, dst, x, e))
    y = (torch.rand(m, generator=torch.Generator().manual_seed(seed + 1)) < 0.7).double()
    return src, dst, x.double(), e.double(), y


def _worker(rank, world, port, kind, norm, n, m, p_long, checkpoint, out_path):
    import gnnome_b200
    from gnnome_b200 import train_dist
    from _emul import TorchPrims
    from oracle import restatement as R
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        src, dst, x, e, y = _inputs(n, m, 3, p_long)
        H, L = 32, 3
        if kind == 'sym':
            model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, L, 64, norm)
        else:
            model = gnnome_b200.models.GatedGCNModel(2, 2, H, 16, L, 64, norm, directed=(kind != 'gated_undirected'))
        model = model.double().train()
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        tr = train_dist.ShardedTrainer(model, src, dst, n, x, e, y, rank, world, torch.device('cpu'), pos_weight=0.4,
                                       prims=TorchPrims(), dtype=torch.float64, checkpoint=checkpoint)
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        loss = tr.step(opt)
        with torch.no_grad():
            scores = tr.scores_in_edge_order(tr.forward().detach())      # a second forward: the stats move once more
        sizes = [None] * world
        dist.all_gather_object(sizes, (tr.shard.n_own, tr.shard.n_halo, tr.shard.num_edges))
        if rank == 0:
            # oracle: the same step on one process through the restatement under torch autograd
            p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
                 for k, v in sd.items()}
            okind, odir = ('gated', False) if kind == 'gated_undirected' else (kind, True)
            ref = R.model_forward(p, src, dst, n, x, e, model=okind, directed=odir, normalization=norm, training=True, cast=False, dtype=torch.float64).squeeze(-1)
            ref_loss = F.binary_cross_entropy_with_logits(ref, y, pos_weight=torch.tensor(0.4, dtype=torch.float64))
            ref_loss.backward()
            gerr = {}
            for k, prm in model.named_parameters():
                g_ref = p[k].grad
                want = sd[k] - 0.05 * g_ref                                  # the SGD step every rank must have taken
                gerr[k] = float((prm.detach() - want).abs().max() / max(float(want.abs().max()), 1e-12))
            berr = {}
            if norm == 'batch':
                # buffers after ONE training forward of the oracle + a second one with the updated parameters
                p2 = {k: v.detach().clone() for k, v in p.items()}
                for k, prm in model.named_parameters():
                    p2[k] = prm.detach().clone()
                R.model_forward(p2, src, dst, n, x, e, model=okind, directed=odir, normalization=norm, training=True, cast=False, dtype=torch.float64)
                for k, b in model.named_buffers():
                    berr[k] = float((b.double() - p2[k].double()).abs().max())
            torch.save({'loss': float(loss), 'ref_loss': float(ref_loss), 'gerr': gerr, 'berr': berr, 'sizes': sizes,
                        'finite': bool(torch.isfinite(scores).all())}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize('world,kind,norm,p_long,checkpoint', [
    (2, 'sym', 'batch', 0.05, True), (3, 'sym', 'batch', 0.3, False), (2, 'gated', 'batch', 0.05, True),
    (2, 'sym', 'layer', 0.1, True), (3, 'gated_undirected', 'batch', 0.1, True)])
def test_sharded_training_step_matches_oracle_autograd_gloo(tmp_path, world, kind, norm, p_long, checkpoint):
    n, m = 400, 2400
    out = str(tmp_path / 'res.pt')
    mp.spawn(_worker, args=(world, _free_port(), kind, norm, n, m, p_long, checkpoint, out), nprocs=world, join=True)
    res = torch.load(out)
    assert res['finite']
    assert abs(res['loss'] - res['ref_loss']) <= 1e-10 * max(1.0, abs(res['ref_loss'])), res
    worst = max(res['gerr'].values())
    assert worst <= 1e-8, sorted(res['gerr'].items(), key=lambda kv: -kv[1])[:5]
    if res['berr']:
        assert max(res['berr'].values()) <= 1e-8, sorted(res['berr'].items(), key=lambda kv: -kv[1])[:5]
    assert sum(s[0] for s in res['sizes']) == n and sum(s[2] for s in res['sizes']) == (2 * m if kind == 'gated_undirected' else m)
    assert all(s[1] > 0 for s in res['sizes'])
