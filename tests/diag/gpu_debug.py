"""Layer-by-layer error report of the CUDA path against the oracle (debug aid, GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import gnnome_b200  # noqa: E402
from gnnome_b200 import GraphIndex, ops, synth  # noqa: E402
from gnnome_b200.layers.encoders import encode_rows  # noqa: E402
from oracle import restatement as R  # noqa: E402


def main():
    sd = torch.load(os.path.join(ROOT, 'tests', 'golden', 'weights.pt'), weights_only=True)
    for H, L, n, m in ((64, 8, 3000, 18000), (32, 2, 500, 3000), (128, 2, 2000, 12000), (256, 2, 1000, 6000)):
        src, dst = synth.make_assembly_graph(n, m, seed=H)
        x, e = synth.make_features(src, dst, n, seed=H)
        src, dst, x, e = map(torch.from_numpy, (src, dst, x, e))
        p = sd if H == 64 else R.init_state_dict(hidden=H, num_layers=L, seed=1)
        model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch')
        model.load_state_dict(p)
        model.eval()
        with torch.no_grad():
            ref_scores, ref_layers = R.model_forward(p, src, dst, n, x, e, dtype=torch.float64, return_layers=True)
            gi = GraphIndex(src, dst, n)
            xd, ed = x.cuda(), e.cuda()
            h = encode_rows(xd, None, model.linear1_node, model.linear2_node, n)
            ep = encode_rows(ed, gi.in_eid, model.linear1_edge, model.linear2_edge, m)
            eid = gi.in_eid[:m].long().cpu()
            ws = {}
            print(f'== H={H} L={L} N={n} E={m}')
            for i, conv in enumerate(model.gnn.convs):
                h, ep = conv.forward_positions(gi, h, ep, ws)
                hr, er = ref_layers[i]
                eh = (h.cpu().double() - hr).abs().max().item()
                ee = (ep.cpu().double() - er[eid]).abs().max().item()
                print(f'  layer {i}: |dh|={eh:.3g} (max |h|={hr.abs().max():.3g})  |de|={ee:.3g} (max |e|={er.abs().max():.3g})')
            sc = model.predictor.forward_positions(gi, h, ep)
            print(f'  scores: |d|={(sc.cpu().double() - ref_scores).abs().max().item():.3g}  prob err='
                  f'{(torch.sigmoid(sc.cpu().double()) - torch.sigmoid(ref_scores)).abs().max().item():.3g}')


if __name__ == '__main__':
    main()
