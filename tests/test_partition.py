"""Multi-GPU host logic on CPU: destination-range partitioning + halo exchanges over ``gloo``
(world_size 2 and 3) must reproduce the single-graph oracle; the arithmetic of each rank is the torch
emulation of the kernel contracts (tests/_emul.py), so only the partition / exchange logic is under test."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from gnnome_b200 import partition, synth  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _inputs(n, m, seed, p_long):
    src, dst = synth.make_assembly_graph(n, m, seed=seed, p_long=p_long)
    x, e = synth.make_features(src, dst, n, seed=seed)
    return tuple(torch.from_numpy(a) for a in (src, dst, x, e))


def _worker(rank, world, port, kind, n, m, p_long, out_path):
    import gnnome_b200
    from _emul import EmulKernels
    from oracle import restatement as R
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        src, dst, x, e = _inputs(n, m, 3, p_long)
        if kind == 'sym':
            model = gnnome_b200.models.SymGatedGCNModel(2, 2, 32, 16, 3, 64, 'batch')
        else:
            model = gnnome_b200.models.GatedGCNModel(2, 2, 32, 16, 3, 64, 'batch', directed=(kind != 'gated_undirected'))
        with torch.no_grad():
            for mod in model.modules():                      # non-trivial eval-mode BatchNorm statistics
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.running_mean.normal_(0, 0.3)
                    mod.running_var.uniform_(0.05, 2.0)
                    mod.weight.uniform_(0.5, 1.5)
                    mod.bias.normal_(0, 0.2)
        model.eval()
        runner = partition.ShardedForward(model, src, dst, n, x, e, rank, world, torch.device('cpu'),
                                          kernels=EmulKernels(torch.float64), dtype=torch.float64)
        with torch.no_grad():
            local = runner.step()
            full = partition.gather_scores(runner, local, m)
        sizes = [None] * world
        dist.all_gather_object(sizes, (runner.shard.n_own, runner.shard.n_halo, int(runner.owned_edge_ids.numel())))
        if rank == 0:
            sd = {k: v.clone() for k, v in model.state_dict().items()}
            ref = R.model_forward(sd, src, dst, n, x, e, model='sym' if kind == 'sym' else 'gated',
                                  directed=(kind != 'gated_undirected'), dtype=torch.float64, faithful=False)
            err = (full.double() - ref).abs().max().item()
            torch.save({'err': err, 'sizes': sizes}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,kind,p_long', [(2, 'sym', 0.01), (3, 'sym', 0.3), (2, 'gated', 0.05),
                                                  (3, 'gated_undirected', 0.05)])
def test_sharded_forward_matches_oracle_gloo(tmp_path, world, kind, p_long):
    n, m = 600, 3600
    out = str(tmp_path / 'res.pt')
    mp.spawn(_worker, args=(world, _free_port(), kind, n, m, p_long, out), nprocs=world, join=True)
    res = torch.load(out)
    assert res['err'] < 1e-5, res
    assert sum(s[0] for s in res['sizes']) == n and sum(s[2] for s in res['sizes']) == m
    assert all(s[1] > 0 for s in res['sizes'])           # every rank really had a halo to exchange


def _degenerate_worker(rank, world, port, out_path):
    import gnnome_b200
    from _emul import EmulKernels
    from oracle import restatement as R
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # every edge but one enters node 5: ranks 1 and 2 own no node and no edge, rank 3 owns one edge and has no halo
        src = torch.tensor([0, 1, 2, 3, 6, 6], dtype=torch.int32)
        dst = torch.tensor([5, 5, 5, 5, 5, 7], dtype=torch.int32)
        n, m = 8, 6
        x, e = torch.randn(n, 2), torch.randn(m, 2)
        model = gnnome_b200.models.SymGatedGCNModel(2, 2, 32, 16, 2, 64, 'batch').eval()
        runner = partition.ShardedForward(model, src, dst, n, x, e, rank, world, torch.device('cpu'),
                                          kernels=EmulKernels(torch.float64), dtype=torch.float64)
        with torch.no_grad():
            full = partition.gather_scores(runner, runner.step(), m)
        sizes = [None] * world
        dist.all_gather_object(sizes, (runner.shard.n_own, runner.shard.n_halo, runner.shard.num_edges))
        if rank == 0:
            ref = R.model_forward({k: v.clone() for k, v in model.state_dict().items()}, src, dst, n, x, e,
                                  dtype=torch.float64, faithful=False)
            torch.save({'err': (full.double() - ref).abs().max().item(), 'sizes': sizes}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ranks_without_work_still_enter_the_exchanges(tmp_path):
    """A rank that owns nothing (and one without a halo) must still take part in every all_to_all: the exchanges are
    collectives.  (They used to be guarded by the rank's own counts, which dead-locked this case.)"""
    out = str(tmp_path / 'res.pt')
    mp.spawn(_degenerate_worker, args=(4, _free_port(), out), nprocs=4, join=True)
    res = torch.load(out)
    assert res['err'] < 1e-12, res
    assert [s[2] for s in res['sizes']] == [5, 0, 0, 1] and res['sizes'][1][0] == 0


def test_node_bounds_balance_in_edges():
    src, dst, _, _ = _inputs(2000, 12000, 1, 0.01)
    b = partition.node_bounds(dst, 2000, 4)
    assert b[0] == 0 and b[-1] == 2000 and all(b[i] <= b[i + 1] for i in range(4))
    cnt = [int(((dst >= b[i]) & (dst < b[i + 1])).sum()) for i in range(4)]
    assert sum(cnt) == 12000 and max(cnt) - min(cnt) < 0.05 * 12000


def test_shard_local_numbering():
    src, dst, _, _ = _inputs(1000, 6000, 2, 0.2)
    sh = partition.Shard(src, dst, 1000, rank=1, world=3)
    glob = torch.cat((torch.arange(sh.lo, sh.hi), sh.halo_nodes))
    assert torch.equal(glob[sh.src_local.long()], src[sh.edge_ids].long())
    assert torch.equal(sh.dst_local.long() + sh.lo, dst[sh.edge_ids].long())
    assert sum(sh.recv_counts) == sh.n_halo and sh.recv_counts[1] == 0
    assert ((sh.halo_nodes < sh.lo) | (sh.halo_nodes >= sh.hi)).all()
