#!/usr/bin/env python
"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ smsp__inst_executed.sum]) per kernel."""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); mi = h.index("Metric Name"); ii = h.index("ID")
t = collections.defaultdict(lambda: [0.0, 0.0, 0]); per = collections.defaultdict(dict)
for r in rows[1:]:
    k = re.sub(r'\(.*', '', r[ki]); v = float(r[vi].replace(',', ''))
    if r[mi].startswith('gpu__time'):
        t[k][0] += v / 1e6; t[k][2] += 1; per[int(r[ii])]['t'] = v / 1e3; per[int(r[ii])]['k'] = k
    else:
        t[k][1] += v; per[int(r[ii])]['i'] = v
tot = sum(v[0] for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} n={v[2]:5d} {v[0]:9.2f} ms {100*v[0]/tot:5.1f}%  {v[1]:.3e} inst")
print(f"total {tot:.2f} ms")
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in sorted(per)[a:b]:
        print(i, per[i]['k'], f"{per[i]['t']:.1f} us", f"{per[i].get('i', 0):.3e}")
