#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def us(x):
    v = float(x['Metric Value'].replace(',', '')); u = x['Metric Unit']
    return v / 1000 if u.startswith('n') else (v * 1000 if u.startswith('m') else v)
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
gagg = collections.defaultdict(lambda: [0, 0.0])
for x in rows:
    name = x['Kernel Name']
    m = re.search(r'(gemm_tc_kernel<[^>]*>|attn_\w+<[^>]*>)', name)
    key = m.group(1) if m else re.sub(r'[<(].*', '', name).replace('void ', '')
    t = us(x); agg[key][0] += 1; agg[key][1] += t; tot += t
    if 'gemm_tc' in name:
        gagg[(key, x['Grid Size'])][0] += 1; gagg[(key, x['Grid Size'])][1] += t
print(f'total {tot/1000:.2f} ms over {len(rows)} launches')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f'{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d} avg {t/n:7.1f}  {k}')
if '-g' in sys.argv:
    for k, (n, t) in sorted(gagg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f'{t:9.1f} us n={n:4d} avg {t/n:7.1f}  {k}')
