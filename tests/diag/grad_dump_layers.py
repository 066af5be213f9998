"""Per-layer forward states and their gradients (training path, BCE loss, swapped degree columns) for offline comparison
with the fp64 oracle.  Output: gpurun_out/grad_layers.pt (edge rows in position order + in_eid)."""
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
    store, count = {}, [0]
    orig = ag.layer_forward

    def wrapped(conv, gi, h, e_pos):
        h2, e2 = orig(conv, gi, h, e_pos)
        i = count[0]
        count[0] += 1
        store[f'h{i}'], store[f'e{i}'] = h2.detach().cpu(), e2.detach().cpu()
        h2.register_hook(lambda gr, i=i: store.__setitem__(f'gh{i}', gr.detach().cpu()))
        e2.register_hook(lambda gr, i=i: store.__setitem__(f'ge{i}', gr.detach().cpu()))
        return h2, e2

    ag.layer_forward = wrapped
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    graph = gnnome_b200.GraphIndex(r['src'], r['dst'], r['num_nodes'])
    logits = model(graph, x, e).squeeze(-1)
    loss = F.binary_cross_entropy_with_logits(logits, r['y'].cuda(), pos_weight=torch.tensor([g['pos_weight']], device='cuda'))
    loss.backward()
    store['in_eid'] = graph.in_eid[:graph.E].cpu()
    store['loss'] = loss.item()
    store['grads'] = {k: p.grad.cpu() for k, p in model.named_parameters()}
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    torch.save(store, os.path.join(ROOT, 'gpurun_out', 'grad_layers.pt'))
    print('loss', loss.item(), 'layers', count[0])


if __name__ == '__main__':
    main()
