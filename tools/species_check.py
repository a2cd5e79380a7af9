#!/usr/bin/env python
"""GPU check of the species rules (which types scatter, which move: kd.c:555-627) against golden results of the
unmodified reference (tests/golden/species_golden.npz, make_species_golden.py).  Written in round 1 after the GPU
budget was spent: RUN THIS FIRST in round 2 (`gpurun -- python tools/species_check.py`) and move the cases into
tests/ once green.  Exit code 1 on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_species_golden import CASES, make_case  # noqa: E402
from oracle.refdump import canonical_labels  # noqa: E402
from skid_b200 import api  # noqa: E402


def main():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "species_golden.npz"))
    bad = 0
    for name in CASES:
        snap, fl, _ = make_case(name)
        res = api.run_skid(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"], **fl)
        nIttr, nBefore, nUnbound, nGroup, nExtra, nAct0, nScat0 = [int(v) for v in gold[name + "_log"]]
        den = gold[name + "_den"]
        act = den > 0
        moved = np.zeros(len(den), bool)
        d = res["moved_r"] - snap["pinit"]["r"][res["moved_iOrder"]]
        moved[res["moved_iOrder"]] = (d - np.round(d)).any(axis=1)
        same = float(np.mean(canonical_labels(gold[name + "_grp"].astype(np.int64)) == canonical_labels(res["grp"].astype(np.int64))))
        checks = {
            "scatter-active set": bool(np.array_equal(res["rho"] > 0, act)),
            "density 1e-5": bool(act.sum() == 0 or (np.abs(res["rho"][act] - den[act]) / den[act]).max() <= 1e-5),
            "nExtraScat": res["nExtraScat"] == nExtra,
            "Ittr:0 line": res["log"][0][2:] == (nAct0, nScat0) if res["log"] else nAct0 == 0,
            "Ittr lines +-1": abs(res["nIttr"] - nIttr) <= 1,
            "moved set": float(np.mean(moved == gold[name + "_moved"])) >= 0.999,
            "groups before unbind": res["nGroupBefore"] == nBefore,
            "groups +-1": abs(res["nGroup"] - nGroup) <= 1,
            "same group >= 0.999": same >= 0.999,
        }
        ok = all(checks.values())
        bad += not ok
        print(("OK   " if ok else "FAIL ") + name, {k: v for k, v in checks.items() if not v} or "", f"same={same:.6f}",
              f"nMove={res['nMove']}", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
