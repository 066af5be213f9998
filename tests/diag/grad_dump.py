"""Dump logits and parameter gradients of the training path for a few graph / feature variants of the hand-off fixture
(BCE loss), for offline comparison with the fp64 oracle.  Output: gpurun_out/grad_dump.pt"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gnnome_b200  # noqa: E402
from gnnome_b200 import assembly as A  # noqa: E402


def main():
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'handoff_losses.pt'), weights_only=True)
    sd = torch.load(os.path.join(ROOT, 'tests', 'golden', 'weights.pt'), weights_only=True)
    r = g['raw']
    ag = A.AssemblyGraph(r['src'], r['dst'], r['num_nodes'], dict(overlap_length=r['overlap_length'],
                         overlap_similarity=r['overlap_similarity'], y=r['y']))
    x, e = A.get_full_ne_features(ag)
    y, pw = r['y'].cuda(), torch.tensor([g['pos_weight']], device='cuda')
    fwd, rev = (r['src'], r['dst'], r['num_nodes']), (r['dst'], r['src'], r['num_nodes'])
    out = {}
    for name, graph, xx in (('org_x', fwd, x), ('rev_xswap', rev, x.flip(1).contiguous()), ('rev_x', rev, x),
                            ('org_xswap', fwd, x.flip(1).contiguous())):
        model = gnnome_b200.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
        model.load_state_dict(sd, strict=True)
        model.cuda().train()
        logits = model(graph, xx, e).squeeze(-1)
        loss = F.binary_cross_entropy_with_logits(logits, y, pos_weight=pw)
        loss.backward()
        out[name] = dict(loss=loss.item(), logits=logits.detach().cpu(),
                         grads={k: p.grad.cpu() for k, p in model.named_parameters()})
        print(name, loss.item())
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    torch.save(out, os.path.join(ROOT, 'gpurun_out', 'grad_dump.pt'))


if __name__ == '__main__':
    main()
