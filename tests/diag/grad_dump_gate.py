"""Inputs / outputs of Gate.backward and Agg.backward for layers 1 and 2 (training path, BCE loss, swapped degree
columns), for offline comparison with an fp64 evaluation.  Output: gpurun_out/grad_gate.pt (rows in position order)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gnnome_b200  # noqa: E402
from gnnome_b200 import assembly as A  # noqa: E402
import gnnome_b200.autograd as ag  # noqa: E402


def main():
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'handoff_losses.pt'), weights_only=True)
    sd = torch.load(os.path.join(ROOT, 'tests', 'golden', 'weights.pt'), weights_only=True)
    r = g['raw']
    agr = A.AssemblyGraph(r['src'], r['dst'], r['num_nodes'], dict(overlap_length=r['overlap_length'],
                          overlap_similarity=r['overlap_similarity'], y=r['y']))
    x, e = A.get_full_ne_features(agr)
    x = x.flip(1).contiguous()
    store = {}
    calls = {'gate': 0, 'agg': 0}
    gate_bwd, agg_bwd, bn_bwd = ag.Gate.backward, ag.Agg.backward, ag.BatchNormTrain.backward

    def gate_wrapped(ctx, g_e, g_sigma):
        layer = 7 - calls['gate']          # backward visits the layers last to first
        calls['gate'] += 1
        out = gate_bwd(ctx, g_e, g_sigma)
        if layer in (1, 2):
            ehat, sigma = ctx.saved_tensors
            store[f'gate{layer}'] = dict(g_e=g_e.cpu(), g_sigma=g_sigma.cpu(), ehat=ehat.cpu(), sigma=sigma.cpu(),
                                         g_ehat=out[0].cpu(), g_ein=None if out[1] is None else out[1].cpu())
        return out

    def agg_wrapped(ctx, gout):
        k = calls['agg']                   # per layer: the out-edge aggregation (mode 1) comes back first, then mode 0
        calls['agg'] += 1
        layer = 7 - k // 2
        out = agg_bwd(ctx, gout)
        if layer in (1, 2):
            store[f'agg{layer}_mode{ctx.mode}'] = dict(gout=gout.cpu(), gA=out[1].cpu(), gsigma=out[2].cpu(),
                                                        den=ctx.saved_tensors[2].cpu(), out=ctx.saved_tensors[3].cpu(),
                                                        A=ctx.saved_tensors[0].cpu())
        return out

    bn_calls = [0]

    def bn_wrapped(ctx, gy, gm, gv):
        k = bn_calls[0]                    # per layer: bn_h comes back first, then bn_e
        bn_calls[0] += 1
        layer, which = 7 - k // 2, ('bn_h', 'bn_e')[k % 2]
        out = bn_bwd(ctx, gy, gm, gv)
        if layer in (1, 2) and which == 'bn_e':
            store[f'bn_e{layer}'] = dict(g=gy.cpu(), x=ctx.saved_tensors[0].cpu(), gx=out[0].cpu(), gw=out[1].cpu(), gb=out[2].cpu())
        return out

    ag.Gate.backward = staticmethod(gate_wrapped)
    ag.Agg.backward = staticmethod(agg_wrapped)
    ag.BatchNormTrain.backward = staticmethod(bn_wrapped)
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    graph = gnnome_b200.GraphIndex(r['src'], r['dst'], r['num_nodes'])
    logits = model(graph, x, e).squeeze(-1)
    loss = F.binary_cross_entropy_with_logits(logits, r['y'].cuda(), pos_weight=torch.tensor([g['pos_weight']], device='cuda'))
    loss.backward()
    store['in_eid'] = graph.in_eid[:graph.E].cpu()
    store['in_src'], store['in_dst'] = graph.in_src[:graph.E].cpu(), graph.in_dst[:graph.E].cpu()
    store['calls'] = dict(calls, bn=bn_calls[0])
    store['grads'] = {k: p.grad.cpu() for k, p in model.named_parameters()}
    torch.save(store, os.path.join(ROOT, 'gpurun_out', 'grad_gate.pt'))
    print('loss', loss.item(), store['calls'], sorted(k for k in store if k[0] in 'gab'))


if __name__ == '__main__':
    main()
