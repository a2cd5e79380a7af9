#!/usr/bin/env python
"""Join an ncu SASS source page (per-instruction executed counts) with nvdisasm --print-line-info
output to get per-CUDA-source-line instruction counts.  usage: ncu_lines.py rep kernel_mangled_substr sassfile units"""
import csv, re, subprocess, sys, collections
rep, kern, sassfile, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]; iI = h.index("Instructions Executed"); iS = h.index("# Samples"); iSrc = h.index("Source")
inst = [(int(r[iI]), int(r[iS]), r[iSrc]) for r in rows[hi + 1:] if len(r) > iI and r[iI].isdigit()]
# parse nvdisasm
lines = open(sassfile).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
cur = None; sass_lines = []
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."): 
        if sass_lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): sass_lines.append(cur)
print("ncu inst", len(inst), "sass inst", len(sass_lines))
agg = collections.defaultdict(lambda: [0, 0])
for (v, s, src), ln in zip(inst, sass_lines):
    agg[ln][0] += v; agg[ln][1] += s
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-inst {tot:.4g} = {tot/units:.1f} per unit")
srcs = {}
for (f, n), (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    if f not in srcs:
        try: srcs[f] = open("/root/repo/skid_b200/csrc/" + f).read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][n - 1].strip()[:90] if n <= len(srcs[f]) else ""
    print(f"{v/units:8.1f}/unit {100*v/tot:5.1f}% stall%={100*s/max(ts,1):5.1f} {f}:{n:4d} {text}")
