"""``inference()`` of the reference up to the walks (inference.py:364-475): score every graph of a dataset directory with a
trained model and decode it greedily.  Per graph ``idx`` it leaves the reference's artefacts:

    {savedir}/decode/{idx}_predicts.pt    edge scores (torch.save, (E,) fp32) -- reused when present (:429-431)
    {savedir}/decode/{idx}_walks.pkl      list of walks (pickle)                (:472-473)
    {savedir}/checkpoint/checkpoint.pkl   decoder checkpoint every 10 contigs   (:344-359)

The graphs are ``AssemblyGraph`` files ``{data_path}/{assembler}/processed/{idx}.pt`` (INTEGRATION.md section 4 has the
exporter from ``.dgl``); successors / predecessors / edge ids come from the reference's own pickles
``{data_path}/{assembler}/info/{idx}_succ.pkl``, ``_pred.pkl``, ``_edges.pkl`` (:447-455).  Turning walks into contig
sequences (``utils/evaluate.py``; needs the reads and Biopython) is not part of this package."""
import os
import shutil
import pickle
import random

import numpy as np
import torch

from . import models
from .assembly import AssemblyGraphDataset, add_positional_encoding, compute_scores, preprocess_graph
from .decode import get_contigs_greedy

# configs/hyperparameters.py of the reference (the values inference() reads, :364-384)
HYPERPARAMETERS = dict(seed=1, num_gnn_layers=8, dim_latent=64, normalization='batch', node_features=2, edge_features=2,
                       hidden_ne_features=16, hidden_edge_scores=64, dropout=0.2, strategy='greedy',
                       num_decoding_paths=100, len_threshold=70_000, decode_with_labels=False, load_checkpoint=True)


def set_seed(seed=42):
    """utils/utils.py:10-29: seed python, numpy and torch (the decoder's start edges are drawn with torch's generator)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def load_model(model_path, hp, dropout=None):
    """inference.py:435-438: build, ``load_state_dict`` (strict), ``eval()``; the parameters may stay on the CPU."""
    model = models.SymGatedGCNModel(hp['node_features'], hp['edge_features'], hp['dim_latent'], hp['hidden_ne_features'],
                                    hp['num_gnn_layers'], hp['hidden_edge_scores'], hp['normalization'], dropout=dropout)
    model.load_state_dict(torch.load(model_path, map_location='cpu', weights_only=True))
    return model.eval()


def inference(data_path, model_path, assembler, savedir, device=None, dropout=None, hyperparameters=None, threads=0,
              fast_sampling=False):
    """Scores + walks for every graph under ``{data_path}/{assembler}`` -> ``{idx: walks}``.  Same positional arguments
    as the reference's ``inference`` (``device`` selects the CUDA device for the scoring pass, not ``'cpu'``)."""
    hp = dict(HYPERPARAMETERS, **(hyperparameters or {}))
    if hp['strategy'] != 'greedy':
        raise ValueError('Invalid decoding strategy')                   # :466-468
    set_seed(hp['seed'])                                                # :389
    inference_dir, checkpoint_dir = os.path.join(savedir, 'decode'), os.path.join(savedir, 'checkpoint')
    os.makedirs(inference_dir, exist_ok=True)
    os.makedirs(checkpoint_dir, exist_ok=True)
    ds = AssemblyGraphDataset(data_path, assembler, preprocess=False)
    model, all_walks = None, {}
    for idx, g in ds:
        if not hp['decode_with_labels']:
            if not os.path.isfile(os.path.join(inference_dir, f'{idx}_predicts.pt')):
                if model is None:
                    model = load_model(model_path, hp, dropout)
                add_positional_encoding(preprocess_graph(g, device=device), device=device)
            compute_scores(model, g, idx, inference_dir, device)        # :408-442
        info = os.path.join(ds.info_dir, str(idx))
        succs, preds, edges = (pickle.load(open(f'{info}_{name}.pkl', 'rb')) for name in ('succ', 'pred', 'edges'))  # :447-455
        prefix = g.edata['prefix_length']
        g.edata['prefix_length'] = prefix.masked_fill(prefix < 0, 0)    # :461: negative prefixes break the contig lengths
        # One check-point directory PER GRAPH, removed once the graph's walks are written: the reference keeps a single
        # {savedir}/checkpoint/checkpoint.pkl for the whole dataset (inference.py:398), so graph k + 1 (or a re-run)
        # would resume from graph k's walks and visited set.
        graph_ckpt = os.path.join(checkpoint_dir, str(idx))
        os.makedirs(graph_ckpt, exist_ok=True)
        walks = get_contigs_greedy(g, succs, preds, edges, hp['len_threshold'], hp['num_decoding_paths'],
                                   hp['decode_with_labels'], graph_ckpt, hp['load_checkpoint'], threads, fast_sampling)
        with open(os.path.join(inference_dir, f'{idx}_walks.pkl'), 'wb') as f:
            pickle.dump(walks, f)                                        # :472-473
        shutil.rmtree(graph_ckpt, ignore_errors=True)
        all_walks[idx] = walks
    return all_walks
