#!/usr/bin/env python
"""Golden results of the UNMODIFIED reference for the species rules (ScatterCriterion kd.c:600-627, CutCriterion
kd.c:555-597): which particle types scatter density and which ones move, for every input type the reference
distinguishes.  Small boxes of the bench generator with the species boundaries moved; runs in the build
container (needs oracle/_ref).  Output tests/golden/species_golden.npz: per case <name>_grp, <name>_den,
<name>_moved (bool: non-zero .ray displacement) and <name>_log.

Usage:  python tests/golden/make_species_golden.py
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

N = 1 << 13
# name -> (nGas, nStar, extra reference flags replacing -gd, python flags replacing bGasAndDark)
CASES = {
    "gas_dark":        (N // 4, 0,      [],             dict(bGasAndDark=False)),              # gas scatters and moves
    "gas_dark_star_gd": (N // 4, N // 8, ["-gd"],        dict(bGasAndDark=True)),               # everything
    "gas_dark_star":   (N // 4, N // 8, [],             dict(bGasAndDark=False)),              # gas + stars
    "gas_dark_star_go": (N // 4, N // 8, ["-go"],        dict(bGasAndDark=False, bGasOnly=True)),  # gas only
    "dark_star":       (0,      N // 4, [],             dict(bGasAndDark=False)),              # stars only
    "gas_only":        (N,      0,      [],             dict(bGasAndDark=False)),              # inType GAS
    # not a species case but the same kind of golden: the dark box WITHOUT -p (no replicas, no wrapping)
    "dark_nonperiodic": (0,     0,      [],             dict(bGasAndDark=False, period=None)),
}


def make_case(name):
    nGas, nStar, ref_extra, py = CASES[name]
    snap = synth.make_box(N, seed=17, kind="gasdark")
    p = snap["pinit"]
    p["fTemp"][:] = 0.0
    p["fTemp"][:nGas] = 1.0e4
    p["fTemp"][:nGas:7] = 5.0e4          # some gas above the -t cut: scatters but does not move
    snap["nGas"], snap["nStar"], snap["nDark"] = nGas, nStar, N - nGas - nStar
    fl = dict(snap["flags"])
    fl.update(py)
    args = [a for a in snap["ref_args"] if a != "-gd"] + ref_extra
    if "period" in py and py["period"] is None:
        i = args.index("-p")
        del args[i:i + 2]
        fl.pop("period")
        # dark-only input: drop the gas flags of the generator too
        for k in ("Omega0", "Lambda", "z", "fTempMax"):
            fl.pop(k, None)
        args = [a for a in args if a not in ("-O", "0.3", "-Lambda", "0.7", "-z", "0.5", "-t", "30000")]
    return snap, fl, args


def main():
    out = {}
    for name in CASES:
        snap, fl, args = make_case(name)
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "in.std")
            synth.write_std(snap, f)
            # Two runs.  With -den the reference re-sorts pInit by iOrder (kdOutDensity -> Order, kd.c:1541) after
            # kdScatterActive put the scatter-active species first: when those are not an iOrder prefix (stars)
            # the move stage then scatters from the WRONG particles - a latent bug of the reference.
            # So the densities come from a -den run (they are computed before the bug bites) and everything else
            # from a run without -den; a third run without -ray checks that -ray does not change the groups.
            refdump.run_ref(f, args + ["-den"], os.path.join(td, "den"))
            out[name + "_den"] = tipsy.read_array(os.path.join(td, "den.den")).astype(np.float32)
            text, _ = refdump.run_ref(f, args + ["-ray"], os.path.join(td, "ref"))
            log = refdump.parse_log(text)
            out[name + "_grp"] = tipsy.read_array(os.path.join(td, "ref.grp")).astype(np.int32)
            out[name + "_moved"] = tipsy.read_vector(os.path.join(td, "ref.ray")).any(axis=1)
            refdump.run_ref(f, args, os.path.join(td, "plain"))
            plain = tipsy.read_array(os.path.join(td, "plain.grp")).astype(np.int32)
            assert np.array_equal(plain, out[name + "_grp"]), name
        out[name + "_log"] = np.array([len(log["ittr"]), log["nGroupBefore"], log["nUnbound"], log["nGroup"],
                                       log.get("nExtraScat", 0), log["ittr"][0][1], log["ittr"][0][2]], np.int64)
        print(name, out[name + "_log"], "scatterers", int((out[name + "_den"] > 0).sum()), "movers",
              int(out[name + "_moved"].sum()), flush=True)
    np.savez_compressed(os.path.join(HERE, "species_golden.npz"), **out)


if __name__ == "__main__":
    main()
