#!/usr/bin/env python
"""A/B of library builds on one GPU box: every .so given on the command line runs the same synthetic box
(1 warm-up pass + 2 timed passes, kernel-family spans on) in its own process; one JSON line per build.

   python tools/ab_variants.py 24 gasdark skid_b200/libskidgpu.so build_ab/libskidgpu_x.so ...

The snapshot is generated once and cached under /tmp.  Measurement scaffolding, not part of the product."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(log2n, kind, lib):
    import numpy as np
    from skid_b200 import api, synth
    api.LIB_PATH = os.path.abspath(lib)
    cache = f"/tmp/ab_{kind}_{log2n}.npy"
    snap = synth.make_box(1 << 12, seed=7, kind=kind)  # flags only
    if os.path.exists(cache):
        pin = np.load(cache)
    else:
        pin = synth.make_box(1 << log2n, seed=7, kind=kind)["pinit"]
        np.save(cache, pin)
    n = len(pin)
    fl = dict(snap["flags"])
    tau = float(np.float32(0.0288 * n ** (-1.0 / 3.0)))
    fl["tau"] = float(np.float32(4.0 * tau)) if kind == "massive" else tau
    nGas = n // 4 if kind == "gasdark" else 0
    per = (fl.pop("period"),) * 3
    sk = api.SkidGPU(per, (0.0, 0.0, 0.0), bPeriodic=True)
    sk.set_profile(True)
    rows = []
    for it in range(3):
        res = api.run_skid(pin, nGas, n - nGas, 0, want_arrays=False, ctx=sk, period=per[0], **fl)
        km = [round(sk.kernel_ms(k)[0], 2) for k in range(5)]
        rows.append(dict(stage_ms=res["stage_ms"], kernel_ms=km, groups=res["nGroup"], unbound=res["nUnbound"],
                         ittr=res["nIttr"], launches=res["launches"]))
    last = rows[-1]
    st = last["stage_ms"]
    tot = sum(st.values()) if isinstance(st, dict) else float(sum(st))
    print(json.dumps(dict(lib=lib, total_ms=tot, stage_ms=st, kernel_ms=last["kernel_ms"], groups=last["groups"],
                          unbound=last["unbound"], ittr=last["ittr"], launches=last["launches"],
                          prev_stage_ms=[{k: round(v, 1) for k, v in r["stage_ms"].items()} for r in rows[:-1]])), flush=True)
    sk.close()


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3], sys.argv[4])
    else:
        log2n, kind = sys.argv[1], sys.argv[2]
        for lib in sys.argv[3:]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", log2n, kind, lib], check=False)
