#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` dump: tools/ncu_hot.py src.csv [N] [section]
(a dump may hold several kernels; `section` picks one, default 0)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
secs = [rows[s:e] for s, e in zip(starts, starts[1:] + [len(rows)])]
for k, sec in enumerate(secs):
    print(f'section {k}: {sec[0][1][:110]} ({len(sec) - 2} rows)')
sec = secs[want]
hdr = sec[1]
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in sec[2:] if len(r) > max(isamp, iexec)]
tot = sum(int(r[isamp] or 0) for r in body)
print('total samples', tot)
ranked = sorted(enumerate(body), key=lambda kv: -int(kv[1][isamp] or 0))[:n]
for idx, r in sorted(ranked):
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    print(f"{idx:5d} {100 * int(r[isamp]) / tot:5.1f}% x{r[iexec]:>9} {r[isrc].strip()[:70]:70s} {' '.join(f'{h}:{v}' for v, h in st if v)}")
