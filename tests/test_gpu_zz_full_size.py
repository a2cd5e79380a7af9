"""CUDA path at BASELINE.json's FULL sizes against golden results of the unmodified reference (CPU hours,
generated once in the build container: tests/golden/make_full_size_golden.py).  C2 = synthetic dark box 2^21
with -nsp (reference with pruning disabled); C2p = the same box with the default scatterer pruning; C3 = the bench workload, gas+dark 2^24; C5 = massive halos 2^24 with
a 4x linking length and -maxgroup 20000 (the serial reference cannot unbind the 2 M member halo in useful time).  Sorted last on purpose:
these are the slowest tests (the 2^24 box takes ~40 s to generate on the host)."""
import numpy as np
import pytest

import fullsize
from skid_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["C2", "C2p", "C3", "C5"])
def test_full_size_box_matches_reference(name):
    gold = fullsize.load(name)
    if gold is None:
        pytest.skip(f"tests/golden/full_{name}.npz not generated")
    snap = synth.make_box(int(gold["n"]), seed=int(gold["seed"]), kind=str(gold["kind"]))
    fl = dict(snap["flags"])
    if name == "C2":
        fl["bNoPrune"] = True
    if name == "C5":
        fl["nMaxMembers"] = 20000   # the golden's -maxgroup (make_full_size_golden.py: EXTRA_ARGS)
    res = api.run_skid(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"], want_arrays=False, **fl)
    rep = fullsize.compare(gold, res["grp"], res["nIttr"], res["nGroupBefore"], res["nUnbound"], res["nGroup"],
                           cat_mass=res["cat"]["fMass"][1:])
    print(name, rep)
