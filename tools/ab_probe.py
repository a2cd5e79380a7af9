#!/usr/bin/env python
"""A/B runs of runtime knobs (environment variables read per pass by the library) on ONE resident snapshot:
   python tools/ab_probe.py [--log2n 24] [--kind gasdark] [--passes 3] "NAME=VAL,NAME2=VAL" "..." ...
Each variant runs `passes` passes of the whole hot path; prints per variant the best total and the stage times of
the best pass plus the result counters (must not change between variants).  An empty string "" = defaults."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skid_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--kind", default="gasdark")
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--host", action="append", default=[],
                    help="also run host/skid -den -ray -stats from a -std file of the same box with this environment "
                         "(NAME=VAL,... ; repeatable; SKID_HOST_TIMING is always set) and print its timing lines")
    ap.add_argument("variants", nargs="*", default=[""])
    a = ap.parse_args()
    snap = synth.make_box(1 << a.log2n, seed=7, kind=a.kind)
    fl = snap["flags"]
    sk = api.SkidGPU((1.0,) * 3, (0.0,) * 3, bPeriodic=True)
    touched = set()
    for var in a.variants:
        for k in touched:
            os.environ.pop(k, None)
        for kv in filter(None, var.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
            touched.add(k)
        best = None
        for _ in range(a.passes):
            res = api.run_skid(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"], ctx=sk, want_arrays=False, **fl)
            tot = sum(res["stage_ms"].values())
            if best is None or tot < best[0]:
                best = (tot, res)
        tot, res = best
        print(json.dumps({"variant": var or "defaults", "total_ms": round(tot, 2),
                          "stage_ms": {k: round(v, 2) for k, v in res["stage_ms"].items()},
                          "move_kernel_ms": round(sk.kernel_ms(0)[0], 2), "nIttr": res["nIttr"],
                          "groups_before": res["nGroupBefore"], "groups": res["nGroup"], "unbound": res["nUnbound"]}),
              flush=True)
    sk.close()
    if a.host:
        import subprocess
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "in.std")
            synth.write_std(snap, f)
            for var in a.host:
                env = dict(os.environ, SKID_HOST_TIMING="1")
                for kv in filter(None, var.split(",")):
                    k, v = kv.split("=")
                    env[k] = v
                args = snap["ref_args"] + ["-den", "-ray", "-stats", "-o", os.path.join(td, "gpu")]
                with open(f, "rb") as fin:
                    r = subprocess.run([os.path.join(ROOT, "host", "skid")] + args, stdin=fin, capture_output=True,
                                       text=True, env=env)
                lines = [json.loads(ln) for ln in r.stderr.splitlines() if ln.startswith("{")]
                gpu = [ln.strip() for ln in r.stdout.splitlines() if ln.startswith("   ")]
                print(json.dumps({"host_variant": var or "defaults", "rc": r.returncode, "timing": lines, "gpu_s": gpu}),
                      flush=True)


if __name__ == "__main__":
    main()
