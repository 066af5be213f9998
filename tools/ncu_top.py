"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['Address']] != 'Address']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S = lambda r, k: int(float(r[idx[k]] or 0))
tot = sum(S(r, '# Samples') for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {s: sum(S(r, s) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
for r in sorted(data, key=lambda r: -S(r, '# Samples'))[:n]:
    st = sorted(((s, S(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(str(S(r, '# Samples')).rjust(7), r[idx['Address']][-5:], r[idx['Source']][:72].ljust(72), st, 'exec', r[idx['Instructions Executed']])
