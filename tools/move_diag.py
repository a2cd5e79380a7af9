#!/usr/bin/env python
"""Development diagnostic: moved positions of the demo vs the reference's .ray (golden), for the
move-kernel variant selected by SKIDGPU_MOVE_KERNEL / SKIDGPU_LIST_WALK_ALWAYS."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_demo_input, DEMO  # noqa: E402
from skid_b200 import api  # noqa: E402

p, ng, nd, ns, _ = load_demo_input()
g = np.load(os.path.join(ROOT, "tests", "golden", "demo_golden.npz"))
sk = api.SkidGPU((1.0,) * 3, (0.0,) * 3, bPeriodic=True)
sk.set_particles(p, ng, nd, ns)
tau = float(np.float32(DEMO["tau"]))
fCvg = float(np.float32(0.5 * tau))
fStep = float(np.float32(0.5 * fCvg))
sk.smDensityInit(64)
sk.move(fDensMin=170.0, fCvg=fCvg, fStep=fStep, keep_step0=True)
iord0, a0, alive = sk.step0()
ic, rc = sk.moved()
sk.kdFoF(tau)
sk.microstep(5, float(np.float32(0.1 * fStep)))
iord, r = sk.moved()
d = r.astype(np.float64) - p["r"][iord]
d -= np.round(d)
err = np.abs(d - g["ray"][iord]).max(axis=1)
print("variant", os.environ.get("SKIDGPU_MOVE_KERNEL", "list"), "walk_always", os.environ.get("SKIDGPU_LIST_WALK_ALWAYS"))
print("ittr lines", len([1 for l in sk.log if l[0] == 0]), "last", [l for l in sk.log if l[0] == 0][-1])
print("err percentiles 50/90/99/99.9/max:", [float(np.percentile(err, q)) for q in (50, 90, 99, 99.9, 100)])
print("n err > 1e-5:", int((err > 1e-5).sum()), " > 1e-4:", int((err > 1e-4).sum()), "of", len(err))
ref_a = np.zeros((len(p), 3))
ref_a[g["step0_iOrder"]] = g["step0_a"]
da = np.linalg.norm(a0 - ref_a[iord0], axis=1) / np.linalg.norm(ref_a[iord0], axis=1)
print("step0 |da|/|a| percentiles 50/99/max:", [float(np.percentile(da, q)) for q in (50, 99, 100)])
if len(sys.argv) > 1:
    np.savez(sys.argv[1], iord=iord, r=r)
