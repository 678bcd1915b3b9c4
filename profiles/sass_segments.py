#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): stall samples per code segment,
segments delimited by barrier-like instructions.  Usage: sass_segments.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    try:
        data.append((r[1].strip(), int(r[col['# Samples']]), int(r[col['Instructions Executed']]),
                     {s: int(r[col[s]] or 0) for s in stalls}, int(r[col['L1 Wavefronts Shared Excessive']] or 0)))
    except Exception:
        pass
tot = sum(d[1] for d in data)
print('total samples', tot, 'sass instructions', len(data))
acc = 0; inst = 0; st = {s: 0 for s in stalls}; start = 0; seg = 0; exc = 0
for i, (s, n, e, sd, x) in enumerate(data):
    acc += n; inst += e; exc += x
    for k, v in sd.items(): st[k] += v
    mark = any(m in s for m in ('BAR.', 'DEPBAR', 'UCGABAR', 'MEMBAR', 'ERRBAR', 'SYNCS.')) or i == len(data) - 1
    if mark:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print('seg %2d [%4d..%4d] %5.1f%% winst %10d smem_excess %9d | %s | %s' % (
            seg, start, i, 100 * acc / max(tot, 1), inst, exc,
            ' '.join('%s=%.1f%%' % (k[6:], 100 * v / max(tot, 1)) for k, v in top), s[:40]))
        seg += 1; acc = 0; inst = 0; st = {s_: 0 for s_ in stalls}; start = i + 1; exc = 0
