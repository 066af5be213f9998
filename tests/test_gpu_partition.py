"""The multi-GPU path's REAL kernels on one GPU (SURVEY.md section 4(5) / 8(e) equivalence test): k destination-range
shards run through ``gnnome_b200.partition.CudaKernels`` -- gnb_gather_rows on the projection table, the halo exchange,
gnb_reverse_partial2 and the ``xp_ptr / xp_row / xp_buf`` branch of gnb_node_update2 -- one host thread per rank on the
same device, with the collectives replaced by an in-process exchange.  The result must equal the single-graph CUDA
forward (to summation order: only the reverse aggregation's partial sums are added in a different order) and the
oracle."""
import threading

import pytest
import torch

from oracle import restatement as R
from gnnome_b200 import synth

pytestmark = pytest.mark.gpu


class ThreadComm:
    """In-process stand-in for ``partition.DistComm``: ``world`` host threads, one per rank, meet at a barrier."""

    class Hub:
        def __init__(self, world):
            self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def __init__(self, hub, rank):
        self.hub, self.rank = hub, rank

    def _exchange(self, payload):
        torch.cuda.synchronize()
        self.hub.slots[self.rank] = payload
        self.hub.barrier.wait()
        got = list(self.hub.slots)
        self.hub.barrier.wait()
        return got

    def all_to_all(self, out, inp, out_splits=None, in_splits=None):
        world = self.hub.world
        if in_splits is None:
            in_splits = [inp.shape[0] // world] * world
        got = self._exchange((inp, in_splits))
        pos = 0
        for r in range(world):
            src, splits = got[r]
            off = sum(splits[:self.rank])
            n = splits[self.rank]
            out[pos:pos + n].copy_(src[off:off + n])
            pos += n
        assert pos == out.shape[0]
        torch.cuda.synchronize()
        self.hub.barrier.wait()          # nobody reuses its input buffer before every peer has copied from it

    def all_reduce_max(self, t):
        got = self._exchange(t.clone())
        t.copy_(torch.stack(got).max(0).values)


def _run_sharded(gnb, model, src, dst, n, x, e, world, overlap=True):
    from gnnome_b200 import partition
    hub = ThreadComm.Hub(world)
    out = torch.empty((src.numel(), 1), dtype=torch.float32, device='cuda')
    info, errors = [None] * world, []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            with torch.no_grad():
                runner = partition.ShardedForward(model, src, dst, n, x, e, rank, world, torch.device('cuda', 0),
                                                  comm=ThreadComm(hub, rank))
                runner.overlap = overlap
                scores = runner.step()
                scores2 = runner.step()                  # buffers are recycled between steps: same answer again
                assert torch.equal(scores, scores2)
                out[runner.owned_edge_ids.to('cuda')] = scores
                info[rank] = (runner.shard.n_own, runner.shard.n_halo, runner.shard.num_edges, runner.plan.n_send)
            torch.cuda.synchronize()
        except BaseException as exc:   # noqa: BLE001
            errors.append((rank, exc))
            hub.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    return out, info


@pytest.mark.parametrize('world,H,L,n,m,p_long', [(2, 64, 8, 20_000, 120_000, 0.01), (3, 256, 3, 30_000, 180_000, 0.2),
                                                  (8, 128, 2, 24_000, 144_000, 0.05), (8, 256, 2, 6_000, 36_000, 1.0)])
def test_k_shards_on_one_gpu_match_single_graph_and_oracle(world, H, L, n, m, p_long, shipped_weights):
    import gnnome_b200 as gnb
    src, dst = synth.make_assembly_graph(n, m, seed=H + world, p_long=p_long)
    x, e = synth.make_features(src, dst, n, seed=H + world)
    src, dst, x, e = map(torch.from_numpy, (src, dst, x, e))
    torch.manual_seed(0)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch')
    if H == 64:
        model.load_state_dict(shipped_weights)
    model = model.cuda().eval()
    with torch.no_grad():
        single = model((src, dst, n), x.cuda(), e.cuda())
    sharded, info = _run_sharded(gnb, model, src, dst, n, x, e, world)
    assert sum(i[0] for i in info) == n and sum(i[2] for i in info) == m
    assert sum(i[1] for i in info) > 0 and sum(i[3] for i in info) > 0          # rows really moved between shards
    prob = lambda t: torch.sigmoid(t.double().cpu())  # noqa: E731
    assert (prob(sharded) - prob(single)).abs().max().item() <= 1e-5
    assert (sharded - single).abs().max().item() <= 2e-4 * max(1.0, single.abs().max().item())
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = R.model_forward(sd, src, dst, n, x, e, faithful=False)
    assert (prob(sharded) - prob(ref)).abs().max().item() <= 1e-4


@pytest.mark.parametrize('world,H,p_long', [(3, 256, 0.05), (8, 128, 0.3)])
def test_overlapped_schedule_is_bit_identical_to_the_serial_one(world, H, p_long):
    """The multi-GPU forward projects the halo rows ahead of the full projection (CudaKernels.project_rows: the h rows
    a peer waits for are gathered and projected on their own) so that the exchange travels under the projection of
    the whole table: the same products on the same inputs, hence the same bits as the serial schedule
    (ShardedForward.overlap = False)."""
    import gnnome_b200 as gnb
    n, m, L = 18_000, 108_000, 3
    src, dst = synth.make_assembly_graph(n, m, seed=7, p_long=p_long)
    x, e = synth.make_features(src, dst, n, seed=7)
    src, dst, x, e = map(torch.from_numpy, (src, dst, x, e))
    torch.manual_seed(1)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').cuda().eval()
    a, info = _run_sharded(gnb, model, src, dst, n, x, e, world, overlap=True)
    b, _ = _run_sharded(gnb, model, src, dst, n, x, e, world, overlap=False)
    assert sum(i[3] for i in info) > 0
    assert torch.equal(a, b)


@pytest.mark.parametrize('world,H', [(3, 64), (4, 128)])
def test_undirected_model_sharded_over_its_doubled_graph(world, H):
    """GatedGCNModel(directed=False): the graph with every edge doubled by its reverse is what gets partitioned; each
    rank scores the edges it owns and returns the original ones (models/full_graph.py:47-52)."""
    import gnnome_b200 as gnb
    n, m, L = 9_000, 54_000, 2
    src, dst = synth.make_assembly_graph(n, m, seed=11, p_long=0.1)
    x, e = synth.make_features(src, dst, n, seed=11)
    src, dst, x, e = map(torch.from_numpy, (src, dst, x, e))
    sd = R.init_state_dict(model='gated', hidden=H, num_layers=L, seed=H)
    model = gnb.models.GatedGCNModel(2, 2, H, 16, L, 64, 'batch', directed=False)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    with torch.no_grad():
        single = model((src, dst, n), x.cuda(), e.cuda())
    sharded, info = _run_sharded(gnb, model, src, dst, n, x, e, world)
    assert sum(i[0] for i in info) == n and sum(i[2] for i in info) == 2 * m     # the doubled graph is what is split
    prob = lambda t: torch.sigmoid(t.double().cpu())  # noqa: E731
    assert (prob(sharded) - prob(single)).abs().max().item() <= 1e-5
    ref = R.model_forward(sd, src, dst, n, x, e, model='gated', directed=False)
    assert (prob(sharded) - prob(ref)).abs().max().item() <= 1e-4
