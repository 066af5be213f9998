"""GPU bring-up of the split16 / TMA kernels: each stage against a torch reference, errors printed (not asserted)."""
import sys, os, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    sd = torch.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'weights.pt'), weights_only=True)
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


which = sys.argv[1:] or ['split', 'linear', 'encode', 'layer', 'model']
for w in which:
    stage(w, globals()['t_' + w])
print('done')
