"""TEST INFRASTRUCTURE: compare one GPU run (skid_b200.api.run_skid result) with the oracle.
Used by __graft_entry__.smoke() and tests; never by the product path."""
import os
import tempfile

import numpy as np

from . import orc, refdump


def check_against_oracle(snap, res, use_reference=True):
    """snap: skid_b200.synth.make_box result; res: run_skid result (want_arrays=True).
    Asserts parity and returns a one-line summary."""
    from skid_b200 import synth, tipsy
    p = snap["pinit"]
    n = len(p)
    msgs = []
    # (1) C restatement: kNN radii bitwise + density 1e-5 on the scatter-active set (dark boxes: everything)
    if snap["nGas"] == 0 and n <= 20000:
        ball2, rho = orc.knn_density(p["r"], p["fMass"], snap["flags"]["nSmooth"], snap["flags"]["period"])
        assert np.array_equal(ball2.view(np.uint32), res["ball2"].view(np.uint32)), "fBall2 not bitwise equal to oracle"
        rel = np.abs(rho.astype(np.float64) - res["rho"]) / rho
        assert rel.max() <= 1e-5, f"density rel err {rel.max()}"
        msgs.append(f"fBall2 bitwise ok ({n}), density max rel {rel.max():.1e}")
    # (2) the unmodified reference binary, when it travelled with the repo
    if use_reference and refdump.have_ref():
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "in.std")
            synth.write_std(snap, f)
            out, dt = refdump.run_ref(f, snap["ref_args"], os.path.join(td, "ref"))
            log = refdump.parse_log(out)
            grp = tipsy.read_array(os.path.join(td, "ref.grp")).astype(np.int64)
        same = float(np.mean(refdump.canonical_labels(grp) == refdump.canonical_labels(res["grp"])))
        assert log["nGroupBefore"] == res["nGroupBefore"], (log["nGroupBefore"], res["nGroupBefore"])
        assert same >= 0.999, f"same-group fraction {same}"
        assert abs(log["nGroup"] - res["nGroup"]) <= max(1, log["nGroup"] // 100)
        msgs.append(f"reference: groups {log['nGroup']} vs {res['nGroup']}, same-group {same:.6f}, ref {dt:.1f}s")
    else:
        msgs.append("reference binary not present")
    return "; ".join(msgs)
