#!/usr/bin/env python
"""Golden results of the UNMODIFIED reference at BASELINE.json's FULL sizes (CPU minutes to an hour each; run in
the build container, needs oracle/_ref):

  C2  synthetic dark box 2^21, seed 1234, demo flags, pruning disabled (= -nsp; skid_ref_dump + SKID_NOPRUNE)
  C3  synthetic gas+dark box 2^24, seed 7 (the bench.py workload), -gd -O 0.3 -Lambda 0.7 -z 0.5 -t 30000
  C5  synthetic massive-halo box 2^24, seed 7, tau x 4, with -maxgroup 20000 (see EXTRA_ARGS)

Stored per case in tests/golden/full_<name>.npz: the reference's log numbers (Ittr lines, groups before
unbinding, unbound, groups), its stage times, the sorted group sizes, and for every STRIDE-th particle the
canonical id of its group (= smallest member iOrder, 0-based, -1 for no group), which lets a test compute the
same-group fraction on that sample without the reference's group numbering; and per final group its canonical id,
member count and bound mass (the .gtp star record) for the bound-mass check.

Usage:  python tests/golden/make_full_size_golden.py C2 [C3] [C5]
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

CASES = {"C2": ("dark", 1 << 21, 1234, True, 4), "C3": ("gasdark", 1 << 24, 7, False, 16),
         # the C2 box with the reference's default scatterer pruning (initial cut + ScatterCut: dark-only input)
         "C2p": ("dark", 1 << 21, 1234, False, 4),
         "C5": ("massive", 1 << 24, 7, False, 16)}
# C5: the serial reference needs O(n^2) pair evaluations per group for the potentials - half a day for the 2 M
# member halo of this box - so its golden is taken with -maxgroup 20000 (groups of >= 20000 members are left
# untouched by kdUnbind, kd.c:1330): density, move, FoF and the unbinding of all other groups are the full run
EXTRA_ARGS = {"C5": ["-maxgroup", "20000"]}


def canonical_min_member(grp):
    """group id -> smallest member index; particles of group 0 get -1."""
    grp = np.asarray(grp, np.int64)
    first = np.full(int(grp.max()) + 1, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(first, grp, np.arange(len(grp), dtype=np.int64))
    out = first[grp]
    out[grp == 0] = -1
    return out.astype(np.int32)


def main():
    for name in sys.argv[1:]:
        kind, n, seed, noprune, stride = CASES[name]
        snap = synth.make_box(n, seed=seed, kind=kind)
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "in.std")
            synth.write_std(snap, f)
            text, wall = refdump.run_ref(f, snap["ref_args"] + EXTRA_ARGS.get(name, []), os.path.join(td, "ref"),
                                         noprune=noprune, timeout=6 * 3600)
            log = refdump.parse_log(text)
            grp = tipsy.read_array(os.path.join(td, "ref.grp")).astype(np.int32)
            gtp_mass = tipsy.read_gtp(os.path.join(td, "ref.gtp"), standard=True)["mass"]
        sizes = np.sort(np.bincount(grp)[1:])[::-1].astype(np.int32)
        canon = canonical_min_member(grp)
        # per final group: canonical id (smallest member index), member count, bound mass from the .gtp
        first = np.full(int(grp.max()) + 1, np.iinfo(np.int64).max, np.int64)
        np.minimum.at(first, grp, np.arange(len(grp), dtype=np.int64))
        group_table = dict(group_canon=first[1:].astype(np.int32), group_count=np.bincount(grp)[1:].astype(np.int32),
                           group_mass=gtp_mass.astype(np.float32))
        np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"),
                            log=np.array([len(log["ittr"]), log["nGroupBefore"], log["nUnbound"], log["nGroup"],
                                          log.get("nExtraScat", 0)], np.int64),
                            ittr=np.array(log["ittr"], np.int32), times=np.array(list(log["times"].values())),
                            time_names=np.array(list(log["times"].keys())), wall_s=wall, sizes=sizes,
                            stride=stride, sample_canon=canon[::stride], n=n, seed=seed, kind=kind, **group_table,
                            ref_args=" ".join(snap["ref_args"] + EXTRA_ARGS.get(name, []))
                            + (" [SKID_NOPRUNE=1]" if noprune else ""))
        print(name, "wall %.0f s" % wall, "log", len(log["ittr"]), log["nGroupBefore"], log["nUnbound"], log["nGroup"],
              flush=True)


if __name__ == "__main__":
    main()
