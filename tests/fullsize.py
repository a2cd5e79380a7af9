"""Helpers for the full-size parity tests (tests/golden/full_<case>.npz, made by make_full_size_golden.py from
the unmodified reference at BASELINE.json's full sizes)."""
import os

import numpy as np

from conftest import GOLDEN


def load(name):
    path = os.path.join(GOLDEN, f"full_{name}.npz")
    return np.load(path) if os.path.exists(path) else None


def canonical_min_member(grp):
    """group id -> smallest member index (0-based); particles of group 0 get -1.  Independent of how the groups
    are numbered, so two catalogues agree on a particle iff these values agree."""
    grp = np.asarray(grp, np.int64)
    first = np.full(int(grp.max()) + 1, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(first, grp, np.arange(len(grp), dtype=np.int64))
    out = first[grp]
    out[grp == 0] = -1
    return out.astype(np.int32)


def compare(gold, grp, nIttr, nGroupBefore, nUnbound, nGroup):
    """north_star tolerances for the synthetic boxes: >= 99.9 % of the particles in the same group; the counters
    of the run agree with the reference's to a fraction of a per cent (they are identical on every box up to
    2^18, but a single mover converging one block later may shift them)."""
    gI, gB, gU, gG, _ = [int(v) for v in gold["log"]]
    stride = int(gold["stride"])
    same = float(np.mean(canonical_min_member(grp)[::stride] == gold["sample_canon"]))
    sizes = np.sort(np.bincount(np.asarray(grp, np.int64))[1:])[::-1]
    k = min(len(sizes), len(gold["sizes"]), 100)
    report = dict(same_group=same, nIttr=(nIttr, gI), groups_before=(nGroupBefore, gB), unbound=(nUnbound, gU),
                  groups=(nGroup, gG), largest=(int(sizes[0]), int(gold["sizes"][0])))
    assert same >= 0.999, report
    assert abs(nIttr - gI) <= 2, report
    assert abs(nGroupBefore - gB) <= max(1, gB // 1000), report
    assert abs(nGroup - gG) <= max(1, gG // 1000), report
    assert abs(nUnbound - gU) <= max(2, gU // 100), report
    assert np.all(np.abs(sizes[:k].astype(np.int64) - gold["sizes"][:k]) <= np.maximum(2, gold["sizes"][:k] // 100)), report
    return report
