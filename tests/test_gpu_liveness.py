"""Liveness of the persistent tensor-core kernels (the round-1 driver bench dead-locked inside gnb_edge_forward_tc2).

* fault injection: the TMA-store thread of the edge kernel is stalled after every tile -- the schedule must terminate
  with bit-identical results however slow that thread is (the old per-group hand-over barrier aliased its phase parity
  exactly then);
* soak: hundreds of forwards at the profiling size with one checksum;
* watchdog: a deliberately dead-locked launch must surface as a CUDA error with a record of who waited for what, in
  seconds (run in a throw-away process: the context does not survive the trap)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from gnnome_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_and_graph(gnb, n, m, H, L=3, seed=5):
    torch.manual_seed(seed)
    model = gnb.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch').cuda().eval()
    src, dst = synth.make_assembly_graph(n, m, seed=seed)
    x, e = synth.make_features(src, dst, n, seed=seed)
    src, dst, x, e = (torch.from_numpy(a) for a in (src, dst, x, e))
    return model, gnb.GraphIndex(src, dst, n), x.cuda(), e.cuda()


@pytest.mark.parametrize('H,n,m', [(256, 60_000, 360_002), (128, 30_000, 180_002), (64, 30_000, 180_030)])
@pytest.mark.parametrize('delay_ns', [3_000, 40_000])
def test_edge_kernel_survives_a_stalled_store_thread(H, n, m, delay_ns):
    import gnnome_b200 as gnb
    from gnnome_b200 import _lib
    lib = _lib.load()
    model, gi, x, e = _model_and_graph(gnb, n, m, H)
    with torch.no_grad():
        ref = model(gi, x, e)
        torch.cuda.synchronize()
        lib.gnb_debug_store_delay_ns(delay_ns)
        try:
            out = model(gi, x, e)
            torch.cuda.synchronize()
        finally:
            lib.gnb_debug_store_delay_ns(0)
    assert torch.equal(out, ref)


def test_soak_300_forwards_one_checksum():
    """300 forwards of the 8-layer H=256 model on the 2M-node / 12M-edge profiling graph: every one terminates and
    returns the same bits (fixed summation order everywhere)."""
    import gnnome_b200 as gnb
    model, gi, x, e = _model_and_graph(gnb, 2_000_000, 12_000_000, 256, L=8, seed=0)
    with torch.no_grad():
        first = model(gi, x, e).clone()
        sums = []
        for it in range(300):
            out = model(gi, x, e)
            sums.append(out.double().sum())
            if it % 50 == 49:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        assert torch.equal(out, first)
    sums = torch.stack(sums).cpu()
    assert torch.isfinite(sums).all() and (sums == sums[0]).all()


_DEADLOCK = textwrap.dedent('''
    import sys, torch
    sys.path.insert(0, {root!r})
    import gnnome_b200 as gnb
    from gnnome_b200 import _lib, synth
    lib = _lib.load()
    lib.gnb_set_spin_timeout_ms(300)
    torch.manual_seed(0)
    model = gnb.models.SymGatedGCNModel(2, 2, {H}, 16, 1, 64, 'batch').cuda().eval()
    src, dst = synth.make_assembly_graph(20000, 120000, seed=1)
    x, e = synth.make_features(src, dst, 20000, seed=1)
    src, dst, x, e = (torch.from_numpy(a) for a in (src, dst, x, e))
    lib.gnb_debug_store_delay_ns(-1)          # the store thread never releases a stage: a real dead-lock
    try:
        with torch.no_grad():
            model((src, dst, 20000), x.cuda(), e.cuda())
        torch.cuda.synchronize()
        print('NO_ERROR')
    except Exception as exc:                  # noqa: BLE001
        print('CUDA_ERROR', type(exc).__name__)
    print('REPORT', _lib.hang_report())
''')


@pytest.mark.parametrize('H', [256, 128])
def test_watchdog_turns_a_deadlock_into_an_error_with_a_record(H):
    r = subprocess.run([sys.executable, '-c', _DEADLOCK.format(root=ROOT, H=H)], capture_output=True, text=True,
                       timeout=300)
    out = r.stdout
    assert 'CUDA_ERROR' in out and 'NO_ERROR' not in out, (out, r.stderr[-2000:])
    report = [ln for ln in out.splitlines() if ln.startswith('REPORT')][0]
    assert 'gnb_edge_forward_tc2' in report and 'waited' in report and 'barrier' in report, report
