#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` dump: tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = rows[2:]
tot = sum(int(r[isamp] or 0) for r in body)
print('total samples', tot)
ranked = sorted(enumerate(body), key=lambda kv: -int(kv[1][isamp] or 0))[:n]
for idx, r in sorted(ranked):
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    print(f"{idx:5d} {100 * int(r[isamp]) / tot:5.1f}% x{r[iexec]:>9} {r[isrc].strip()[:70]:70s} {' '.join(f'{h}:{v}' for v, h in st if v)}")
