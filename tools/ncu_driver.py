#!/usr/bin/env python
"""One pass of the hot path (+ skidgpu_stats) on a synthetic box, for ncu captures:
   ncu --set full --clock-control none --import-source on -k regex:... python tools/ncu_driver.py [log2n] [kind]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skid_b200 import api, synth  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
kind = sys.argv[2] if len(sys.argv) > 2 else "gasdark"
snap = synth.make_box(1 << log2n, seed=7, kind=kind)
res = api.run_skid(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"], want_arrays=False, want_stats=True,
                   **snap["flags"])
print("groups", res["nGroup"], "stage_ms", res["stage_ms"])
