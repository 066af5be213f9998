"""Generate ``tests/golden/handoff_*.pt``: the callers' code either side of ``model(g, x, e)``, run from the reference's
OWN unmodified functions (``utils/data_utils.py:preprocess_graph``, ``train.py:get_full_ne_features``,
``symmetry_loss``, ``get_bce_loss_full``, ``get_symmetry_loss_full``) over ``oracle/dgl_shim`` on CPU.

TEST INFRASTRUCTURE.  Run in the build container only:  ``python -m oracle.make_golden_handoff``.
Those modules cannot be imported whole here (Biopython, ``dgl.data``), so ``reference_runner.load_functions`` executes
the functions' source with the module-level names supplied by hand.  ``add_positional_encoding`` needs scipy/``dgl.backend``
only past its early return (``nb_pos_enc`` = 0, configs/hyperparameters.py:26); its two live lines
(utils/data_utils.py:50-51) are applied directly.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_runner as rr  # noqa: E402
from gnnome_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def raw_graph(n, m, seed):
    """An assembly graph as graph_parser leaves it: int64 overlap lengths, fp32 similarities, 0/1 labels."""
    src, dst = synth.make_assembly_graph(n, m, seed=seed)
    rng = np.random.default_rng(seed)
    ol_len = torch.from_numpy(rng.integers(500, 25000, size=src.size))             # int64
    ol_sim = torch.from_numpy(rng.uniform(0.9, 1.0, size=src.size).astype(np.float32))
    y = torch.from_numpy((rng.random(src.size) < 0.75).astype(np.float32))
    return torch.from_numpy(src), torch.from_numpy(dst), n, ol_len, ol_sim, y


def main():
    dgl, layers, models = rr.load()
    from configs.hyperparameters import get_hyperparameters
    ns = dict(torch=torch, F=F, dgl=dgl, get_hyperparameters=get_hyperparameters)
    rr.load_functions('utils/data_utils.py', {'preprocess_graph'}, ns)
    rr.load_functions('train.py', {'symmetry_loss', 'get_full_ne_features', 'get_bce_loss_full',
                                   'get_symmetry_loss_full'}, ns)
    hp = get_hyperparameters()
    sd = torch.load(rr.weights_path(), weights_only=True)

    def build(n, m, seed):
        src, dst, n, ol_len, ol_sim, y = raw_graph(n, m, seed)
        g = dgl.graph((src, dst), num_nodes=n)
        g.edata['overlap_length'], g.edata['overlap_similarity'], g.edata['y'] = ol_len, ol_sim, y
        g = ns['preprocess_graph'](g)
        g.ndata['in_deg'] = g.in_degrees().float()     # utils/data_utils.py:50
        g.ndata['out_deg'] = g.out_degrees().float()   # utils/data_utils.py:51
        raw = dict(src=src, dst=dst, num_nodes=n, overlap_length=ol_len, overlap_similarity=ol_sim, y=y)
        return g, raw

    # (a) scoring hand-off: raw graph -> features -> shipped model -> what inference.py:441-442 saves
    g, raw = build(1500, 9000, 21)
    x, e = ns['get_full_ne_features'](g, reverse=False)
    x_rev, _ = ns['get_full_ne_features'](dgl.reverse(g, True, True), reverse=True)
    m = models.SymGatedGCNModel(hp['node_features'], hp['edge_features'], hp['dim_latent'],
                                hp['hidden_ne_features'], hp['num_gnn_layers'], hp['hidden_edge_scores'],
                                hp['normalization'], dropout=None)   # inference.py:364-375,435
    m.load_state_dict(sd, strict=True)
    m.eval()
    with rr.quiet(), torch.no_grad():
        predicts = m(g, x, e).squeeze()
    torch.save(dict(raw=raw, x=x, x_rev=x_rev, e=e, predicts=predicts), os.path.join(OUT, 'handoff_scores.pt'))
    print('handoff_scores', tuple(predicts.shape), float(predicts.min()), float(predicts.max()))

    # (b) training losses on the full graph (train.py:138-145 and :158-170), dropout off so the step is deterministic
    g, raw = build(600, 3600, 23)
    pos_weight, alpha = torch.tensor([1.0 / 3.0]), hp['alpha']
    rec = dict(raw=raw, pos_weight=float(pos_weight), alpha=alpha)
    for name, fn, args in (('bce', 'get_bce_loss_full', (pos_weight, 'cpu')),
                           ('sym', 'get_symmetry_loss_full', (pos_weight, alpha, 'cpu'))):
        mt = models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
        mt.load_state_dict(sd, strict=True)
        mt.train()
        seen = []   # every forward's output: [org] for BCE, [org, rev] for the symmetry loss
        hook = mt.register_forward_hook(lambda mod, inp, out: seen.append(out.detach().squeeze(-1).clone()))
        with rr.quiet():
            loss, logits = ns[fn](g, mt, *args)
        hook.remove()
        loss.backward()
        rec[name] = dict(loss=loss.detach(), logits=logits.detach(), forwards=seen,
                         grads={k: p.grad.clone() for k, p in mt.named_parameters()},
                         buffers={k: b.clone() for k, b in mt.named_buffers()})
        print(name, 'loss', float(loss))
    torch.save(rec, os.path.join(OUT, 'handoff_losses.pt'))


if __name__ == '__main__':
    main()
