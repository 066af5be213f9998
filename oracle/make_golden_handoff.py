"""Generate ``tests/golden/handoff_*.pt``: the callers' code either side of ``model(g, x, e)``, run from the reference's
OWN unmodified functions (``utils/data_utils.py:preprocess_graph``, ``train.py:get_full_ne_features``,
``symmetry_loss``, ``get_bce_loss_full``, ``get_symmetry_loss_full``) over ``oracle/dgl_shim`` on CPU.

TEST INFRASTRUCTURE.  Run in the build container only:  ``python -m oracle.make_golden_handoff``.
Re-running reproduces every forward value bit for bit; the stored gradients move by ~1e-7 of their scale between runs
(torch's threaded CPU backward of the index ops does not fix its summation order).
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
                                   'get_symmetry_loss_full', 'mask_graph_strandwise', 'get_partition_ne_features',
                                   'get_bce_loss_partition', 'get_symmetry_loss_partition'}, ns)
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

    # (c) strand-wise masking (train.py:91-100) and a mini-batch (train.py:125-135, 148-155, 173-185): induced
    #     subgraphs with '_ID' maps, their features and losses.  dgl.node_subgraph is the shim's restatement of DGL's
    #     documented behaviour; METIS is not involved -- the mini-batch is a contiguous block of nodes.
    g, raw = build(1200, 7200, 29)
    torch.manual_seed(5)
    with rr.quiet():
        masked = ns['mask_graph_strandwise'](g, 0.8, 'cpu')
    keep = torch.zeros(g.num_nodes(), dtype=torch.bool)
    keep[masked.ndata['_ID'].long()] = True
    rec = dict(raw=raw, pos_weight=float(pos_weight), alpha=alpha, mask_keep=keep,
               mask_node_id=masked.ndata['_ID'], mask_edge_id=masked.edata['_ID'],
               mask_src=masked.edges()[0], mask_dst=masked.edges()[1],
               mask_in_deg=masked.ndata['in_deg'], mask_e=masked.edata['e'], mask_y=masked.edata['y'])

    def losses(fn, *args):
        mt = models.SymGatedGCNModel(2, 2, 64, 16, 8, 64, 'batch', dropout=None)
        mt.load_state_dict(sd, strict=True)
        mt.train()
        seen = []
        hook = mt.register_forward_hook(lambda mod, inp, out: seen.append(out.detach().squeeze(-1).clone()))
        with rr.quiet():
            loss, logits = ns[fn](*args[:-1], mt, *args[-1])
        hook.remove()
        return dict(loss=loss.detach(), logits=logits.detach(), forwards=seen)

    rec['mask_sym'] = losses('get_symmetry_loss_full', masked, (pos_weight, alpha, 'cpu'))
    batch_keep = torch.zeros(g.num_nodes(), dtype=torch.bool)
    batch_keep[300:700] = True
    sub = dgl.node_subgraph(g, batch_keep, store_ids=True)
    x_b, e_b = ns['get_partition_ne_features'](sub, g, reverse=False)
    rec.update(batch_keep=batch_keep, batch_node_id=sub.ndata['_ID'], batch_edge_id=sub.edata['_ID'],
               batch_x=x_b, batch_e=e_b,
               batch_bce=losses('get_bce_loss_partition', sub, g, (pos_weight, 'cpu')),
               batch_sym=losses('get_symmetry_loss_partition', sub, g, (pos_weight, alpha, 'cpu')))
    torch.save(rec, os.path.join(OUT, 'handoff_subgraphs.pt'))
    print('masked', masked.num_nodes(), masked.num_edges(), 'loss', float(rec['mask_sym']['loss']),
          '| batch', sub.num_nodes(), sub.num_edges(), float(rec['batch_bce']['loss']), float(rec['batch_sym']['loss']))


if __name__ == '__main__':
    main()
