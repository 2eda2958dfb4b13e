"""Aggregate the warp-stall samples of an ncu source page by kernel region (prologue / RK4 loop / epilogue)
and list the most-stalled instructions outside the loop.
    ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_regions.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = hdr_idx[0]
i1 = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
hdr, body = rows[i0], rows[i0 + 1:i1]
col = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[col['# Samples']])
ex = [int(r[col['Instructions Executed']]) for r in body]
tot = sum(S(r) for r in body)
mx = max(ex)
loop = [n for n, e in enumerate(ex) if e >= mx * 0.9]
print(rows[i0 - 1][1] if i0 else '', '\ninstructions', len(body), 'samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for name, (a, b) in {'prologue': (0, loop[0]), 'loop': (loop[0], loop[-1] + 1), 'epilogue': (loop[-1] + 1, len(body))}.items():
    s = sum(S(r) for r in body[a:b])
    e = sum(ex[a:b])
    top = sorted(((sum(int(r[col[h]]) for r in body[a:b]), h) for h in stalls), reverse=True)[:5]
    print('%-9s instr %5d..%5d  samples %6d (%4.1f%%)  warp-inst %10d (%4.1f%%)  %s' % (
        name, a, b, s, 100 * s / tot, e, 100 * e / sum(ex), ' '.join('%s=%d' % (h[6:], v) for v, h in top)))
out = sorted(((S(r), n, r[col['Source']].strip()) for n, r in enumerate(body) if n < loop[0] or n > loop[-1]), reverse=True)
for s, n, src in out[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print('%6d  #%d  %s' % (s, n, src))
