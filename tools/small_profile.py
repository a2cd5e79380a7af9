#!/usr/bin/env python
"""Per-kernel-family device time of the move stage on a small input (the demo), to see what a launch-bound
run spends its time on.  usage: python tools/small_profile.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import DEMO, load_demo_input
from skid_b200 import api
p, ng, nd, ns, _ = load_demo_input()
sk = api.SkidGPU((1.0,) * 3, (0.0,) * 3, bPeriodic=True)
for prof in (False, True, False):
    sk.set_profile(prof)
    t0 = time.perf_counter()
    res = api.run_skid(p, ng, nd, ns, ctx=sk, want_arrays=False, **DEMO)
    wall = time.perf_counter() - t0
    fam = {k: sk.kernel_ms(w) for k, w in dict(tile_step=0, knn=1, builds=2, fallback=3, prune=4).items()}
    print("profile", prof, "wall ms %.2f" % (wall * 1e3), {k: round(v, 2) for k, v in res["stage_ms"].items()}, "launches", res["launches"],
          {k: (round(v[0], 2), v[1]) for k, v in fam.items()})
sk.close()
