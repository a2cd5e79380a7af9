#!/usr/bin/env python
"""Development probe: run the GPU pipeline on synthetic boxes of growing size, print stage times,
and (optionally) compare the group catalogue with the reference binary run on the same input."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402
from skid_b200 import api, synth, tipsy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[15, 18])
    ap.add_argument("--kind", default="dark")
    ap.add_argument("--ref-max", type=int, default=16, help="run the reference up to this log2 n")
    ap.add_argument("--noprune", action="store_true")
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    for l2 in a.log2n:
        n = 1 << l2
        t0 = time.time()
        snap = synth.make_box(n, seed=1234, kind=a.kind)
        tg = time.time() - t0
        for rep in range(a.repeat):
            t0 = time.time()
            res = api.run_skid(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"], bNoPrune=a.noprune,
                               want_arrays=False, **snap["flags"])
            wall = time.time() - t0
            ms = res["stage_ms"]
            tot = sum(ms.values())
            print(f"n=2^{l2} kind={a.kind} gen={tg:.1f}s wall={wall:.2f}s gpu_ms={tot:.1f} "
                  + " ".join(f"{k}={v:.1f}" for k, v in ms.items())
                  + f" nMove={res['nMove']} nIttr={res['nIttr']} moverSteps={res['mover_steps']}"
                  + f" groupsBefore={res['nGroupBefore']} unbound={res['nUnbound']} groups={res['nGroup']}"
                  + f" launches={res['launches']} particles/s={n / (tot * 1e-3):.3e}", flush=True)
        if l2 <= a.ref_max and refdump.have_ref():
            with tempfile.TemporaryDirectory() as td:
                f = os.path.join(td, "in.std")
                synth.write_std(snap, f)
                out, dt = refdump.run_ref(f, snap["ref_args"], os.path.join(td, "ref"), noprune=a.noprune)
                log = refdump.parse_log(out)
                grp = tipsy.read_array(os.path.join(td, "ref.grp")).astype(np.int64)
            same = np.mean(refdump.canonical_labels(grp) == refdump.canonical_labels(res["grp"]))
            print(f"   reference: wall={dt:.1f}s times={log['times']} ittr={len(log['ittr'])} "
                  f"groupsBefore={log['nGroupBefore']} unbound={log['nUnbound']} groups={log['nGroup']} "
                  f"same-group fraction={same:.6f}", flush=True)


if __name__ == "__main__":
    main()
