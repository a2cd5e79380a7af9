#!/usr/bin/env python
"""Measurements for the "next" rows of SURVEY 8f at the size the metric is quoted on (2^24 gas+dark box):
  f1  skidgpu_stats (kdOutStats on the device) vs the oracle's C restatement (qsort per group, 1 core)
  f2  host/skid end to end from a -std file with -den -ray -stats: host wall-clock phases (SKID_HOST_TIMING)
Prints one JSON line; run on the GPU box:  python tools/next_rows_bench.py [log2n]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402  (checker + CPU baseline only)
from skid_b200 import api, synth  # noqa: E402


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << log2n
    snap = synth.make_box(n, seed=7, kind="gasdark")
    fl = snap["flags"]
    p = snap["pinit"]
    out = {"n": n, "workload": f"synthetic gas+dark box 2^{log2n} (bench.py generator, seed 7)"}
    # ---- f1: stats on the device, timed around the synchronous C-ABI call (includes the D2H of the rows)
    sk = api.SkidGPU((1.0,) * 3, (0.0,) * 3, bPeriodic=True)
    res = api.run_skid(p, snap["nGas"], snap["nDark"], snap["nStar"], ctx=sk, want_stats=True, **fl)
    f32 = lambda v: float(np.float32(v))
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        rows = sk.kdOutStats(1.0, f32(fl.get("z", 0.0)), res["fCosmo"], fl["fDensMin"], fl.get("fTempMax", api.FLT_MAX))
        ts.append(time.perf_counter() - t0)
    sk.close()
    t0 = time.perf_counter()
    ref = orc.stats(p["r"], p["v"], p["fMass"], p["fSoft"], p["fTemp"], res["rho"], snap["nGas"], snap["nDark"],
                    res["grp"], len(res["cat"]), res["cat"]["rCenter"], res["cat"]["vcm"], (1.0,) * 3, 1.0,
                    f32(fl.get("z", 0.0)), res["fCosmo"], fl["fDensMin"], fl.get("fTempMax", api.FLT_MAX))
    t_cpu = time.perf_counter() - t0
    exact = np.ones(len(rows), bool)
    for f in rows.dtype.names:
        exact &= rows[f] == ref[f]
    out["stats"] = {"groups": int(len(rows) - 1), "grouped_particles": int((res["grp"] > 0).sum()),
                    "gpu_ms_best_of_3": 1e3 * min(ts), "gpu_ms_all": [1e3 * t for t in ts],
                    "oracle_cpu_s_1_core": t_cpu, "rows_bit_identical_frac": float(exact[1:].mean()),
                    "largest_group": int(rows["nMembers"].max())}
    # ---- f2: the C driver end to end from a file
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "in.std")
        synth.write_std(snap, f)
        args = snap["ref_args"] + ["-den", "-ray", "-stats", "-o", os.path.join(td, "gpu")]
        env = dict(os.environ, SKID_HOST_TIMING="1")
        t0 = time.perf_counter()
        with open(f, "rb") as fin:
            r = subprocess.run([os.path.join(ROOT, "host", "skid")] + args, stdin=fin, capture_output=True, text=True,
                               env=env)
        wall = time.perf_counter() - t0
        timing = [ln for ln in r.stderr.splitlines() if ln.startswith("{\"host_wall_s\"")]
        out["host_skid"] = {"rc": r.returncode, "process_wall_s": wall, "host_cores": os.cpu_count(),
                            "phases": json.loads(timing[-1]) if timing else None,
                            "gpu_time_lines": [ln.strip() for ln in r.stdout.splitlines() if ":" in ln and "   " in ln][-5:],
                            "output_bytes": {e: os.path.getsize(os.path.join(td, "gpu." + e))
                                             for e in ("grp", "den", "ray", "gtp", "stat")
                                             if os.path.exists(os.path.join(td, "gpu." + e))}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
