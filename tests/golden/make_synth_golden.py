#!/usr/bin/env python
"""Generate tests/golden/synth_golden.npz: results of the UNMODIFIED reference (oracle/_ref/skid_ref, built by
oracle/build_ref.sh from /root/reference) on small synthetic boxes of skid_b200.synth (the generator of BASELINE
configs 2-5).  Runs only in the build container.  Keys per case <name>: <name>_grp (final .grp), <name>_den,
<name>_log = [nIttr lines, Groups before Unbind, particles Unbound, Number of Groups, nExtraScat], <name>_gtp_mass
(bound mass per final group, numbering of <name>_grp).

Usage:  python tests/golden/make_synth_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402
from skid_b200 import synth, tipsy  # noqa: E402

CASES = {"dark13": ("dark", 1 << 13, 3), "gasdark13": ("gasdark", 1 << 13, 11), "dark14": ("dark", 1 << 14, 5),
         "massive14": ("massive", 1 << 14, 9),
         # BASELINE configs[1] semantics: -nsp, exact scatterer handling = the reference with pruning disabled
         # (skid_ref_dump with SKID_NOPRUNE=1, the one-line patch described in oracle/build_ref.sh)
         "dark13_nsp": ("dark", 1 << 13, 3)}


def main():
    out = {}
    for name, (kind, n, seed) in CASES.items():
        snap = synth.make_box(n, seed=seed, kind=kind)
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "in.std")
            synth.write_std(snap, f)
            text, _ = refdump.run_ref(f, snap["ref_args"] + ["-den"], os.path.join(td, "ref"),
                                      noprune=name.endswith("_nsp"))
            log = refdump.parse_log(text)
            out[name + "_grp"] = tipsy.read_array(os.path.join(td, "ref.grp")).astype(np.int32)
            out[name + "_den"] = tipsy.read_array(os.path.join(td, "ref.den")).astype(np.float32)
            # bound masses of the final groups (the .gtp star records, kd.c:1665-1682), in the .grp's numbering
            out[name + "_gtp_mass"] = tipsy.read_gtp(os.path.join(td, "ref.gtp"), standard=True)["mass"]
        out[name + "_log"] = np.array([len(log["ittr"]), log["nGroupBefore"], log["nUnbound"], log["nGroup"],
                                       log.get("nExtraScat", 0)], np.int64)
        print(name, out[name + "_log"])
    np.savez_compressed(os.path.join(HERE, "synth_golden.npz"), **out)


if __name__ == "__main__":
    main()
