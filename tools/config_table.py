#!/usr/bin/env python
"""One line per BASELINE.json config that fits one GPU (C1 demo, C2 2^21 dark -nsp, C3 2^24 gas+dark,
C5 2^24 massive halos / tau x 4): stage times, counts, particles/s.  C4 (2^27, 8 GPUs) is `bench.py --gpus 8`."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from skid_b200 import api, synth  # noqa: E402


def run(name, pinit, nGas, nDark, nStar, flags, noprune=False, repeat=2, golden=None):
    best = None
    for _ in range(repeat):
        res = api.run_skid(pinit, nGas, nDark, nStar, bNoPrune=noprune, want_arrays=False, **flags)
        tot = sum(res["stage_ms"].values())
        if best is None or tot < best[0]:
            best = (tot, res)
    tot, res = best
    n = len(pinit)
    vs_ref = None
    if golden is not None:  # full-size golden of the unmodified reference (tests/golden/make_full_size_golden.py)
        import fullsize
        gold = fullsize.load(golden)
        if gold is not None:
            try:
                vs_ref = fullsize.compare(gold, res["grp"], res["nIttr"], res["nGroupBefore"], res["nUnbound"], res["nGroup"],
                                           cat_mass=res["cat"]["fMass"][1:])
                vs_ref["reference_cpu_s"] = float(np.sum(gold["times"]))
            except AssertionError as e:
                vs_ref = {"MISMATCH": str(e)[:400]}
    print(json.dumps(dict(config=name, vs_reference_full_size=vs_ref, n=n, gpu_ms=round(tot, 2), particles_per_s=n / (tot * 1e-3),
                          stage_ms={k: round(v, 2) for k, v in res["stage_ms"].items()}, nMove=int(res["nMove"]),
                          nIttr=int(res["nIttr"]), mover_steps=int(res["mover_steps"]),
                          groups_before_unbind=int(res["nGroupBefore"]), unbound=int(res["nUnbound"]),
                          groups=int(res["nGroup"]))), flush=True)


def main():
    from conftest import DEMO, load_demo_input
    p, nGas, nDark, nStar, _ = load_demo_input()
    run("C1 dark.std demo (32768 dark)", p, nGas, nDark, nStar, DEMO)
    s = synth.make_box(1 << 21, seed=1234, kind="dark")
    run("C2 dark 2^21 -nsp", s["pinit"], s["nGas"], s["nDark"], s["nStar"], s["flags"], noprune=True, golden="C2")
    s = synth.make_box(1 << 24, seed=7, kind="gasdark")
    run("C3 gas+dark 2^24", s["pinit"], s["nGas"], s["nDark"], s["nStar"], s["flags"], golden="C3")
    s = synth.make_box(1 << 24, seed=7, kind="massive")
    run("C5 massive halos 2^24, tau x 4 (every group unbound, no -maxgroup)", s["pinit"], s["nGas"], s["nDark"], s["nStar"], s["flags"])


if __name__ == "__main__":
    main()
