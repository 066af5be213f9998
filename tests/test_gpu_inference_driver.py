"""GPU leg of gnnome_b200.inference.inference: no predictions on disk -> the model scores the graph, then the decode."""
import pickle

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_inference_driver_scores_and_decodes(golden, tmp_path):
    from gnnome_b200.assembly import AssemblyGraph
    from gnnome_b200.inference import inference
    import os
    g = golden('handoff_scores')
    r = g['raw']
    src, dst, n = r['src'], r['dst'], r['num_nodes']
    succs, preds, edges = {i: [] for i in range(n)}, {i: [] for i in range(n)}, {}
    for k, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        succs[u].append(v)
        preds[v].append(u)
        edges[(u, v)] = k
    data, save = tmp_path / 'data', tmp_path / 'out'
    (data / 'hifiasm' / 'processed').mkdir(parents=True)
    (data / 'hifiasm' / 'info').mkdir()
    gen = torch.Generator().manual_seed(0)
    AssemblyGraph(src, dst, n, dict(overlap_length=r['overlap_length'], overlap_similarity=r['overlap_similarity'],
                                    prefix_length=torch.randint(100, 9000, (src.numel(),), generator=gen)),
                  dict(read_length=torch.randint(8000, 25000, (n,), generator=gen))).save(data / 'hifiasm' / 'processed' / '4.pt')
    for name, obj in (('succ', succs), ('pred', preds), ('edges', edges)):
        pickle.dump(obj, open(data / 'hifiasm' / 'info' / f'4_{name}.pkl', 'wb'))
    weights = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'weights.pt')
    out = inference(str(data), weights, 'hifiasm', str(save),
                    hyperparameters=dict(num_decoding_paths=10, len_threshold=0, load_checkpoint=False))
    scores = torch.load(save / 'decode' / '4_predicts.pt', weights_only=True)
    assert scores.shape == g['predicts'].shape
    assert (torch.sigmoid(scores.double()) - torch.sigmoid(g['predicts'].double())).abs().max().item() <= 1e-4
    walks = pickle.load(open(save / 'decode' / '4_walks.pkl', 'rb'))
    assert out == {4: walks} and walks and all(len(w) >= 2 for w in walks)
    used = [v for w in walks for v in w]
    assert len(used) == len(set(used))
