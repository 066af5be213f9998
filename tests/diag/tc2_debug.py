"""GPU bring-up of the split16 / TMA kernels: each stage against a torch reference, errors printed (not asserted)."""
import sys, os, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gnnome_b200
from gnnome_b200 import ops, synth
from oracle import restatement as R

torch.manual_seed(0)
dev = 'cuda'


def stage(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        print(f'[{name}] EXCEPTION')
        traceback.print_exc()


def t_split():
    for rows, K in [(5, 64), (1000, 256), (77, 128)]:
        x = torch.randn(rows, K, device=dev) * 3
        x16 = ops.split_rows(x)
        y = ops.merge_rows(x16)
        print(f'[split] rows={rows} K={K} roundtrip max rel err {((y - x).abs() / x.abs().clamp_min(1e-3)).max().item():.3g}')
        idx = torch.randperm(rows, device=dev).to(torch.int32)
        y2 = ops.merge_rows(ops.split_rows(x, idx), idx)
        print(f'[split] permuted roundtrip max abs err {(y2 - x).abs().max().item():.3g}')


def t_linear():
    for rows, K, M in [(64, 64, 128), (130, 64, 320), (1000, 128, 640), (777, 256, 1280), (513, 256, 128), (1, 256, 128)]:
        g = torch.Generator().manual_seed(rows)
        a, w, b = torch.randn(rows, K, generator=g), torch.randn(M, K, generator=g) / K ** 0.5, torch.randn(M, generator=g)
        ref = a.double() @ w.double().t() + b.double()
        out = ops.node_linear_tc2(ops.split_rows(a.cuda()), ops.pack_linear_tc(w.cuda()), b.cuda(), M)
        torch.cuda.synchronize()
        err = (out.cpu().double() - ref).abs()
        print(f'[linear_tc2] rows={rows} K={K} M={M} max err {err.max().item():.3g} (mean {err.mean().item():.3g})')
        if err.max().item() > 1e-3:
            bad = (err > 1e-3).nonzero()
            print('   first bad (row, col):', bad[:5].tolist(), ' bad rows mod 8:', sorted(set((bad[:, 0] % 8).tolist()))[:8],
                  ' bad cols/64:', sorted(set((bad[:, 1] // 64).tolist()))[:8])


def t_encode():
    for rows, H in [(5, 64), (1000, 256), (333, 128)]:
        g = torch.Generator().manual_seed(H)
        x = torch.randn(rows + 7, 2, generator=g)
        idx = torch.randint(0, rows + 7, (rows,), generator=g).to(torch.int32)
        W1, b1 = torch.randn(16, 2, generator=g), torch.randn(16, generator=g)
        W2, b2 = torch.randn(H, 16, generator=g), torch.randn(H, generator=g)
        ref = torch.relu(x[idx.long()].double() @ W1.double().t() + b1.double()) @ W2.double().t() + b2.double()
        o16, o32 = ops.encode2(x.cuda(), idx.cuda(), W1.cuda(), b1.cuda(), W2.t().contiguous().cuda(), b2.cuda(), rows,
                               want16=True, want32=True)
        print(f'[encode2] rows={rows} H={H} fp32 err {(o32.cpu().double() - ref).abs().max().item():.3g} '
              f'split16 err {(ops.merge_rows(o16).cpu().double() - ref).abs().max().item():.3g}')


def _hub_graph(n, m, hub_deg, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n // 2, size=m)
    dst = rng.integers(0, n // 2, size=m)
    hub = n // 3
    src = np.concatenate([src, rng.integers(0, n, size=hub_deg), np.full(hub_deg, hub), [5, 5, 5]])
    dst = np.concatenate([dst, np.full(hub_deg, hub), rng.integers(0, n, size=hub_deg), [5, 6, 6]])
    return torch.from_numpy(src.astype(np.int32)), torch.from_numpy(dst.astype(np.int32))


def t_layer():
    for H in (64, 128, 256):
        for sym in (True, False):
            n = 700
            src, dst = _hub_graph(n, 3000, 1500, seed=H)
            m = src.numel()
            torch.manual_seed(H + sym)
            layer = (gnnome_b200.layers.SymGatedGCN if sym else gnnome_b200.layers.GatedGCN)(H, H, 'batch')
            with torch.no_grad():
                for bn in (layer.bn_h, layer.bn_e):
                    bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.05, 2.0)
                    bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
            layer.eval()
            h, e = torch.randn(n, H), torch.randn(m, H)
            p = {'L.' + k: v for k, v in layer.state_dict().items()}
            fn = R.sym_gated_gcn_layer if sym else R.gated_gcn_layer
            with torch.no_grad():
                p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
                h_ref, e_ref = fn(p64, 'L.', src.long(), dst.long(), n, h.double(), e.double())
                res = {}
                for be in ('tc2', 'tc'):
                    gnnome_b200.set_backend(be)
                    h_out, e_out = layer((src, dst, n), h.cuda(), e.cuda())
                    torch.cuda.synchronize()
                    res[be] = ((h_out.cpu().double() - h_ref).abs().max().item(), (e_out.cpu().double() - e_ref).abs().max().item())
                gnnome_b200.set_backend('tc2')
            print(f'[layer] H={H} sym={sym} tc2 (h,e) err {res["tc2"][0]:.3g} {res["tc2"][1]:.3g} | tc {res["tc"][0]:.3g} {res["tc"][1]:.3g}'
                  f' | scale h {h_ref.abs().max().item():.3g} e {e_ref.abs().max().item():.3g}')


def t_model():
    sd = torch.load(os.path.join(os.path.dirname(__file__), '..', 'golden', 'weights.pt'), weights_only=True)
    for H, L, n, m in [(64, 8, 20000, 120000), (128, 4, 10000, 60000), (256, 3, 6000, 36000), (256, 8, 100000, 600000)]:
        s, d = synth.make_assembly_graph(n, m, seed=H)
        x, e = synth.make_features(s, d, n, seed=H)
        s, d, x, e = map(torch.from_numpy, (s, d, x, e))
        sdd = sd if H == 64 else R.init_state_dict(hidden=H, num_layers=L, seed=H)
        model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, L, 64, 'batch')
        model.load_state_dict(sdd, strict=True)
        model.eval()
        with torch.no_grad():
            truth = R.model_forward(sdd, s, d, n, x, e, dtype=torch.float64, faithful=False)
            for be in ('tc2', 'tc'):
                gnnome_b200.set_backend(be)
                out = model((s, d, n), x, e)
                torch.cuda.synchronize()
                err = (torch.sigmoid(out.double().cpu()) - torch.sigmoid(truth)).abs().max().item()
                print(f'[model] H={H} L={L} E={m} backend={be} max prob err vs fp64 {err:.3g}')
            gnnome_b200.set_backend('tc2')
            a = model((s, d, n), x, e); b = model((s, d, n), x, e)
            print(f'[model] deterministic: {torch.equal(a, b)}')


def t_determinism():
    """Each kernel twice on identical inputs (H=64 shipped weights, 300k edges): bitwise equal?"""
    from gnnome_b200.layers.encoders import encode_rows2
    sd = torch.load(os.path.join(os.path.dirname(__file__), '..', 'golden', 'weights.pt'), weights_only=True)
    for H, n, m in [(64, 50000, 300000), (128, 50000, 300000), (256, 50000, 300002)]:
        s, d = synth.make_assembly_graph(n, m, seed=5)
        x, e = synth.make_features(s, d, n, seed=5)
        s, d, x, e = map(torch.from_numpy, (s, d, x, e))
        sdd = sd if H == 64 else R.init_state_dict(hidden=H, num_layers=8, seed=H)
        model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, 8, 64, 'batch')
        model.load_state_dict(sdd, strict=True)
        model.eval()
        gi = gnnome_b200.GraphIndex(s, d, n)
        xd, ed = x.cuda(), e.cuda()
        with torch.no_grad():
            def enc():
                h16, h32 = encode_rows2(xd, None, model.linear1_node, model.linear2_node, gi.N, want32=True)
                e16, _ = encode_rows2(ed, gi.in_eid, model.linear1_edge, model.linear2_edge, gi.E)
                return h16, h32, e16
            a, b = enc(), enc()
            print(f'[det H={H}] encode2 equal:', all(torch.equal(u, v) for u, v in zip(a, b)))
            h16, h32, e16 = a
            conv = model.gnn.convs[0]
            pk = conv._pack(xd.device)
            nb = 5
            P1 = ops.node_linear_tc2(h16, pk['Wn_t'], pk['bn'], nb * H)
            P2 = ops.node_linear_tc2(h16, pk['Wn_t'], pk['bn'], nb * H)
            print(f'[det H={H}] node_linear_tc2 equal:', torch.equal(P1, P2))
            outs = []
            for rep in range(3):
                ec = e16.clone()
                Fb = torch.zeros((gi.N, H), device='cuda')
                carry = torch.zeros((gi.num_chunks(H, 'tc'), 4, H), device='cuda')
                tf, ep = gi.tile_flags(H, 'tc2') if H > 128 else (None, 0)
                ops.edge_forward_tc2(gi, H, P1, pk['We_t'], pk['scale_e'], pk['shift_e'], ec, Fb, carry, tf, ep, conv._flags())
                torch.cuda.synchronize()
                outs.append((ec, Fb, carry))
            for rep in (1, 2):
                eq = [torch.equal(u, v) for u, v in zip(outs[0], outs[rep])]
                print(f'[det H={H}] edge_forward_tc2 run0 vs run{rep} equal (e16, F, carry):', eq)
                if not eq[0]:
                    diff = (outs[0][0] != outs[rep][0]).nonzero()
                    print('    e16 diffs:', diff.shape[0], 'first', diff[:6].tolist(), 'rows/64 set', sorted(set((diff[:, 1] // 64).tolist()))[:10])
                if not eq[1]:
                    diff = (outs[0][1] != outs[rep][1]).nonzero()
                    print('    F diffs:', diff.shape[0], 'first', diff[:6].tolist())
                if not eq[2]:
                    diff = (outs[0][2] != outs[rep][2]).nonzero()
                    print('    carry diffs:', diff.shape[0], 'first', diff[:6].tolist())
            ec, Fb, carry = outs[0]
            # cross-check against the first-generation kernel on fp32 state
            e32 = ops.merge_rows(e16)
            F1 = torch.zeros((gi.N, H), device='cuda'); c1 = torch.zeros_like(carry)
            tf, ep = gi.tile_flags(H) if H > 128 else (None, 0)
            ops.edge_forward_tc(gi, H, P1, pk['We_t'], pk['scale_e'], pk['shift_e'], e32, F1, c1, tf, ep, conv._flags())
            print(f'[det H={H}] tc2 vs tc: e max diff {(ops.merge_rows(ec) - e32).abs().max().item():.3g} F max diff {(Fb - F1).abs().max().item():.3g} carry {(carry - c1).abs().max().item():.3g}')
            res = []
            for rep in range(2):
                ho, h16o = torch.empty_like(h32), torch.empty_like(h16)
                ops.node_update2(gi, H, P1, ec, Fb, carry, h32, pk['scale_h'], pk['shift_h'], ho, h16o, conv._flags(), gi.chunk(H, 'tc'))
                res.append((ho, h16o))
            print(f'[det H={H}] node_update2 equal:', all(torch.equal(u, v) for u, v in zip(*res)))
            S = model.predictor.node_rows16(res[0][1])
            s1 = model.predictor.score_positions16(gi, S, ec); s2 = model.predictor.score_positions16(gi, S, ec)
            print(f'[det H={H}] score_tc2 equal:', torch.equal(s1, s2))
            o1 = model(gi, xd, ed); o2 = model(gi, xd, ed)
            print(f'[det H={H}] model equal:', torch.equal(o1, o2), 'max diff', (o1 - o2).abs().max().item())


def t_layers():
    """Two full passes layer by layer (H=64 shipped weights, 300k edges): first tensor that differs."""
    from gnnome_b200.layers.encoders import encode_rows2
    sd = torch.load(os.path.join(os.path.dirname(__file__), '..', 'golden', 'weights.pt'), weights_only=True)
    H, n, m = 64, 50000, 300000
    s, d = synth.make_assembly_graph(n, m, seed=5)
    x, e = synth.make_features(s, d, n, seed=5)
    s, d, x, e = map(torch.from_numpy, (s, d, x, e))
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, 8, 64, 'batch')
    model.load_state_dict(sd, strict=True)
    model.eval()
    gi = gnnome_b200.GraphIndex(s, d, n)
    xd, ed = x.cuda(), e.cuda()
    with torch.no_grad():
        _, ref_layers = R.model_forward(sd, s, d, n, x, e, dtype=torch.float64, faithful=False, return_layers=True)
        order = gi.in_eid[:gi.E].cpu().long()
        passes = []
        for rep in range(3):
            sync = rep == 2
            h16, h32 = encode_rows2(xd, None, model.linear1_node, model.linear2_node, gi.N, want32=True)
            e16, _ = encode_rows2(ed, gi.in_eid, model.linear1_edge, model.linear2_edge, gi.E)
            ws, snaps = {}, []
            for li, conv in enumerate(model.gnn.convs):
                h32, h16, e16 = conv.forward_positions16(gi, h32, h16, e16, ws)
                if sync:
                    torch.cuda.synchronize()
                snaps.append((h32.clone(), h16.clone(), e16.clone()))
            passes.append(snaps)
        for li in range(8):
            eq01 = [torch.equal(a, b) for a, b in zip(passes[0][li], passes[1][li])]
            eq02 = [torch.equal(a, b) for a, b in zip(passes[0][li], passes[2][li])]
            errs = []
            for p_ in passes:
                hh = (p_[li][0].cpu().double() - ref_layers[li][0]).abs().max().item()
                ee = (ops.merge_rows(p_[li][2]).cpu().double() - ref_layers[li][1][order]).abs().max().item()
                errs.append((hh, ee))
            print(f'[layers] layer {li}: pass0==pass1 (h32,h16,e16) {eq01}  pass0==pass2(sync) {eq02}  err vs fp64 (h,e) per pass: '
                  + ' | '.join(f'{a:.2g},{b:.2g}' for a, b in errs))
            if not eq01[2]:
                diff = (passes[0][li][2] != passes[1][li][2]).nonzero()
                print('    e16 diff count', diff.shape[0], 'first', diff[:4].tolist(), 'last', diff[-2:].tolist(),
                      'tiles', sorted(set((diff[:, 1] // 64).tolist()))[:12])
            if not eq01[0]:
                diff = (passes[0][li][0] != passes[1][li][0]).nonzero()
                print('    h32 diff count', diff.shape[0], 'first', diff[:4].tolist())


def t_badrow():
    """Layer 0 of the H=64 shipped model at 300k edges, workspace pre-filled with NaN: dump the rows that are off."""
    from gnnome_b200.layers.encoders import encode_rows2
    sd = torch.load(os.path.join(os.path.dirname(__file__), '..', 'golden', 'weights.pt'), weights_only=True)
    H, n, m = 64, 50000, 300000
    s, d = synth.make_assembly_graph(n, m, seed=5)
    x, e = synth.make_features(s, d, n, seed=5)
    s, d, x, e = map(torch.from_numpy, (s, d, x, e))
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, H, 16, 8, 64, 'batch')
    model.load_state_dict(sd, strict=True)
    model.eval()
    gi = gnnome_b200.GraphIndex(s, d, n)
    xd, ed = x.cuda(), e.cuda()
    with torch.no_grad():
        _, ref_layers = R.model_forward(sd, s, d, n, x, e, dtype=torch.float64, faithful=False, return_layers=True)
        order = gi.in_eid[:gi.E].cpu().long()
        e_ref = ref_layers[0][1][order]
        for rep in range(3):
            h16, h32 = encode_rows2(xd, None, model.linear1_node, model.linear2_node, gi.N, want32=True)
            e16, _ = encode_rows2(ed, gi.in_eid, model.linear1_edge, model.linear2_edge, gi.E)
            e_in = ops.merge_rows(e16).cpu()
            nan = float('nan')
            ws = {'P': torch.full((gi.N, 5 * H), nan, device='cuda'), 'F': torch.full((gi.N, H), nan, device='cuda'),
                  'carry': torch.full((gi.num_chunks(H, 'tc'), 4, H), nan, device='cuda')}
            conv = model.gnn.convs[0]
            h32o, h16o, e16o = conv.forward_positions16(gi, h32, h16, e16, ws)
            torch.cuda.synchronize()
            got = ops.merge_rows(e16o).cpu().double()
            err = (got - e_ref).abs()
            bad = (err.max(dim=1).values > 1e-2).nonzero().squeeze(1)
            print(f'[badrow rep {rep}] e rows off by > 1e-2: {bad.numel()}  nan in e: {torch.isnan(got).sum().item()}  nan in h: {torch.isnan(h32o).sum().item()}'
                  f'  max |e_ref| {e_ref.abs().max().item():.4g}  max |h| {h32o.abs().max().item():.4g}')
            P = ws['P'].cpu()
            src_pos, dst_pos = gi.in_src[:gi.E].cpu().long(), gi.in_dst[:gi.E].cpu().long()
            for r in bad[:6].tolist():
                ch = (err[r] > 1e-2).nonzero().squeeze(1)
                c = ch[0].item()
                print(f'   row {r} (tile {r // 64} row-in-tile {r % 64}) src {src_pos[r].item()} dst {dst_pos[r].item()} '
                      f'indeg(dst) {(dst_pos == dst_pos[r]).sum().item()} bad channels {ch[:4].tolist()}..{ch[-1].item()} (n={ch.numel()})')
                print(f'      ch {c}: got {got[r, c].item():.6g} ref {e_ref[r, c].item():.6g} e_in {e_in[r, c].item():.6g} '
                      f'B1h[src] {P[src_pos[r], 2 * c].item():.6g} B2h[dst] {P[dst_pos[r], 2 * H + c].item():.6g}')
            hb = ((h32o.cpu().double() - ref_layers[0][0]).abs().max(dim=1).values > 1e-2).nonzero().squeeze(1)
            print(f'   h rows off by > 1e-2: {hb.numel()} first {hb[:8].tolist()}')


which = sys.argv[1:] or ['split', 'linear', 'encode', 'layer', 'model']
for w in which:
    stage(w, globals()['t_' + w])
print('done')
