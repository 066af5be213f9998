"""Aggregate ncu warp-stall samples per CUDA source line.

usage: ncu_lines.py <ncu --page source --csv output> <nvdisasm -g -c output> <kernel name substring> [top]
The ncu CSV carries SASS addresses + samples; nvdisasm -g gives offset -> file:line."""
import csv, re, sys, collections

src_csv, sass_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- offset -> line from nvdisasm
line_of, cur, infn = {}, None, False
for ln in open(sass_path, errors='replace'):
    if ln.startswith('.text.'):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['Address']] != 'Address']
data = data[:len(data) // 2] if len(data) > 2 * len(line_of) - 10 else data   # ncu prints the table twice
base = int(data[0][idx['Address']], 16)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S = lambda r, k: int(float(r[idx[k]] or 0))
agg = collections.defaultdict(lambda: collections.Counter())
for r in data:
    off = int(r[idx['Address']], 16) - base
    key = line_of.get(off, ('?', 0))
    c = agg[key]
    c['samples'] += S(r, '# Samples'); c['exec'] += S(r, 'Instructions Executed'); c['n_inst'] += 1
    for s in stalls:
        c[s] += S(r, s)
tot = sum(c['samples'] for c in agg.values())
print('total samples', tot, 'sass instructions', len(data))
allst = collections.Counter()
for c in agg.values():
    for s in stalls: allst[s] += c[s]
print({k: f'{100*v/tot:.1f}%' for k, v in allst.most_common(8)})
srcs = {}
for (f, l), c in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    if f not in srcs:
        try: srcs[f] = open(f'/root/repo/gnnome_b200/csrc/{f}').read().split('\n')
        except OSError: srcs[f] = []
    text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ''
    st = ' '.join(f'{s[6:]}={100*c[s]/max(c["samples"],1):.0f}%' for s, _ in sorted(((s, c[s]) for s in stalls), key=lambda kv: -kv[1])[:3])
    print(f'{100*c["samples"]/tot:5.1f}% {f}:{l:<4d} inst={c["n_inst"]:<4d} exec={c["exec"]:<10d} {text:70s} {st}')
