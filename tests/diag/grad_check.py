"""Gradient accuracy of the training path against an fp64 evaluation of the oracle restatement (CPU), next to the
accuracy of the reference's own fp32 gradients (the fixture).  Prints, per parameter, max |g - g64| / max |g64|.
Usage: python tests/diag/grad_check.py [bce|sym]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import restatement as R  # noqa: E402  (checker only)
import gnnome_b200  # noqa: E402
from gnnome_b200 import assembly as A  # noqa: E402


def combine(kind, org_fn, rev_fn, y, pw, alpha):
    """bce: BCE(org) | rev: BCE(rev), reversed forward only | both: BCE(org) + BCE(rev) | abs: mean |org - rev| |
    sym: the symmetry loss of train.py:103-109"""
    bce = lambda s: F.binary_cross_entropy_with_logits(s, y, pos_weight=pw)  # noqa: E731
    if kind == 'bce':
        return bce(org_fn())
    if kind == 'rev':
        return bce(rev_fn())
    org, rev = org_fn(), rev_fn()
    if kind == 'both':
        return bce(org) + bce(rev)
    if kind == 'abs':
        return (org - rev).abs().mean()
    return (F.binary_cross_entropy_with_logits(org, y, pos_weight=pw, reduction='none') +
            F.binary_cross_entropy_with_logits(rev, y, pos_weight=pw, reduction='none') + alpha * (org - rev).abs()).mean()


def truth(g, sd, kind):
    r = g['raw']
    src, dst, n = r['src'], r['dst'], r['num_nodes']
    dt = torch.float64
    p = {}
    for k, v in sd.items():
        p[k] = v.clone().to(dt).requires_grad_('running' not in k) if v.is_floating_point() else v.clone()
    e = R.edge_input_features(r['overlap_length'], r['overlap_similarity']).to(dt)
    y, pw = r['y'].to(dt), torch.tensor([g['pos_weight']], dtype=dt)
    fwd = lambda s, d, rev: R.model_forward(p, s, d, n, R.node_input_features(src, dst, n, reverse=rev).to(dt), e,  # noqa: E731
                                            training=True, cast=False, dtype=dt).squeeze(-1)
    loss = combine(kind, lambda: fwd(src, dst, False), lambda: fwd(dst, src, True), y, pw, g['alpha'])
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items() if v.is_floating_point() and v.requires_grad}


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else 'sym'
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'handoff_losses.pt'), weights_only=True)
    sd = torch.load(os.path.join(ROOT, 'tests', 'golden', 'weights.pt'), weights_only=True)
    l64, g64 = truth(g, sd, kind)
    r = g['raw']
    ag = A.AssemblyGraph(r['src'], r['dst'], r['num_nodes'], dict(overlap_length=r['overlap_length'],
                         overlap_similarity=r['overlap_similarity'], y=r['y']))
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    pw = torch.tensor([g['pos_weight']], device='cuda')
    y = r['y'].cuda()
    g_rev = ag.reversed()

    def org_fn():
        x, e = A.get_full_ne_features(ag, reverse=False)
        return model(ag, x, e).squeeze(-1)

    def rev_fn():
        g_rev.edata, g_rev.ndata = dict(ag.edata), dict(ag.ndata)
        x, e = A.get_full_ne_features(g_rev, reverse=True)
        return model(g_rev, x, e).squeeze(-1)

    A.get_full_ne_features(ag)                       # degrees / e on the parent before it is reversed
    loss = combine(kind, org_fn, rev_fn, y, pw, g['alpha'])
    loss.backward()
    have_ref = kind in g
    print(f'{kind}: loss fp64 {l64:.9f}  ours {loss.item():.9f}' + (f'  reference {g[kind]["loss"].item():.9f}' if have_ref else ''))
    rows = []
    for k, p in model.named_parameters():
        t = g64[k]
        scale = max(t.abs().max().item(), 1e-12)
        ours = (p.grad.cpu().double() - t).abs().max().item() / scale
        ref = (g[kind]['grads'][k].double() - t).abs().max().item() / scale if have_ref else float('nan')
        rows.append((ours, ref, scale, k))
    t, o = g64['predictor.W1.weight'], model.predictor.W1.weight.grad.cpu().double()
    H = t.shape[1] // 3
    for name, sl in (('W1[:, src block]', slice(0, H)), ('W1[:, dst block]', slice(H, 2 * H)), ('W1[:, edge block]', slice(2 * H, 3 * H))):
        sc = t[:, sl].abs().max().item()
        print(f'predictor.{name:20s} err {(o[:, sl] - t[:, sl]).abs().max().item() / sc:.2e} scale {sc:.2e}')
    print(f'{"parameter":36s} {"ours":>10s} {"reference":>10s} {"scale":>10s}')
    for ours, ref, scale, k in rows:
        flag = '  <--' if ours > 2e-4 and scale > 1e-5 else ''
        print(f'{k:36s} {ours:10.2e} {ref:10.2e} {scale:10.2e}{flag}')


if __name__ == '__main__':
    main()
