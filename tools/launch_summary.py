#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: tools/launch_summary.py in.csv [out.md] [title]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('mr::<unnamed>::', '')
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
    a = agg.setdefault((name, row['Grid Size'], row['Block Size']), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
out = [f"# {sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}\n",
       f"{sum(a[0] for a in agg.values())} launches, total {tot / 1e3:.2f} ms (per-launch times are cold-cache and serialised: compare shares).\n",
       "| kernel | grid | block | launches | total ms | avg us | share |", "|---|---|---|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[0][:80]}` | {k[1]} | {k[2]} | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% |")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text)
print(text)
