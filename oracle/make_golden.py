"""Generate ``tests/golden/*.pt`` by running the reference's OWN unmodified model code
(``/root/reference/{layers,models}``) on CPU over ``oracle/dgl_shim``.

TEST INFRASTRUCTURE.  Run in the build container only:  ``python -m oracle.make_golden``.
The fixtures (inputs + reference outputs) travel to the GPU box; ``/root/reference`` does not.
``weights.pt`` is the reference's shipped ``weights/weights.pt`` (a data fixture, 0.9 MB), kept
so that parity is checked on real trained parameters (BN running_var spans 5e-5 .. 460).
"""
import os
import shutil
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_runner as rr  # noqa: E402
from gnnome_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _graph_inputs(kind, seed):
    if kind == 'tiny':
        src, dst, n = synth.tiny_adversarial_graph()
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((n, 2)).astype(np.float32)
        e = rng.standard_normal((src.size, 2)).astype(np.float32)
    else:
        n, m = kind
        src, dst = synth.make_assembly_graph(n, m, seed=seed)
        x, e = synth.make_features(src, dst, n, seed=seed)
    return (torch.from_numpy(src), torch.from_numpy(dst), n, torch.from_numpy(x), torch.from_numpy(e))


def _hook_layers(model):
    outs = []
    hooks = [c.register_forward_hook(lambda m, i, o: outs.append((o[0].clone(), o[1].clone())))
             for c in model.gnn.convs]
    return outs, hooks


def main():
    os.makedirs(OUT, exist_ok=True)
    dgl, layers, models = rr.load()
    shutil.copyfile(rr.weights_path(), os.path.join(OUT, 'weights.pt'))
    sd = torch.load(rr.weights_path(), weights_only=True)

    # (i) shipped SymGatedGCNModel, eval: tiny adversarial graph (+ per-layer states) and a 2k/12k graph
    m = models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    m.load_state_dict(sd, strict=True)
    m.eval()
    for name, kind, seed in (('sym_shipped_tiny', 'tiny', 7), ('sym_shipped_2k', (2000, 12000), 11)):
        src, dst, n, x, e = _graph_inputs(kind, seed)
        g = dgl.graph((src, dst), num_nodes=n)
        outs, hooks = _hook_layers(m)
        with rr.quiet(), torch.no_grad():
            logits = m(g, x, e)
        for hk in hooks:
            hk.remove()
        rec = dict(src=src, dst=dst, num_nodes=n, x=x, e=e, logits=logits)
        if kind == 'tiny':
            rec['layers'] = outs
        torch.save(rec, os.path.join(OUT, name + '.pt'))
        print(name, tuple(logits.shape), float(logits.min()), float(logits.max()))

    # (ii) GatedGCNModel (one-direction layers), directed and undirected, seeded init; LayerNorm Sym model
    src, dst, n, x, e = _graph_inputs((600, 3600), 5)
    g = dgl.graph((src, dst), num_nodes=n)
    for name, ctor in (
            ('gated_directed', lambda: models.GatedGCNModel(2, 2, 32, 16, 3, 64, 'batch', directed=True)),
            ('gated_undirected', lambda: models.GatedGCNModel(2, 2, 32, 16, 3, 64, 'batch', directed=False)),
            ('sym_layernorm', lambda: models.SymGatedGCNModel(2, 2, 32, 16, 3, 64, 'layer')),
            ('sym_h128', lambda: models.SymGatedGCNModel(2, 2, 128, 16, 2, 64, 'batch'))):
        torch.manual_seed(3)
        mm = ctor()
        # non-trivial BN statistics so that the eval-mode fold is exercised
        for k, v in mm.state_dict().items():
            if k.endswith('running_mean'):
                v.copy_(torch.randn_like(v) * 0.3)
            if k.endswith('running_var'):
                v.copy_(torch.rand_like(v) * 2 + 0.05)
        mm.eval()
        with rr.quiet(), torch.no_grad():
            logits = mm(g, x, e)
        torch.save(dict(src=src, dst=dst, num_nodes=n, x=x, e=e, logits=logits,
                        state_dict={k: v.clone() for k, v in mm.state_dict().items()}),
                   os.path.join(OUT, name + '.pt'))
        print(name, tuple(logits.shape))

    # (iii) one training step of the shipped model (train.py:138-145: BCE-with-logits, pos_weight),
    #       dropout=0 so the step is deterministic: loss, all gradients, BN buffers afterwards
    src, dst, n, x, e = _graph_inputs((600, 3600), 9)
    g = dgl.graph((src, dst), num_nodes=n)
    y = (torch.rand(src.numel(), generator=torch.Generator().manual_seed(1)) < 0.75).float()
    mt = models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
    mt.load_state_dict(sd, strict=True)
    mt.train()
    with rr.quiet():
        logits = mt(g, x, e).squeeze(-1)
    loss = F.binary_cross_entropy_with_logits(logits, y, pos_weight=torch.tensor(1.0 / 3.0))
    loss.backward()
    torch.save(dict(src=src, dst=dst, num_nodes=n, x=x, e=e, y=y, pos_weight=1.0 / 3.0,
                    logits=logits.detach(), loss=loss.detach(),
                    grads={k: p.grad.clone() for k, p in mt.named_parameters()},
                    buffers={k: b.clone() for k, b in mt.named_buffers()}),
               os.path.join(OUT, 'sym_shipped_trainstep.pt'))
    print('trainstep loss', float(loss))


if __name__ == '__main__':
    main()
