"""Parity of the CUDA path (through the C-ABI) with the reference on config 1, the reference's own
demo (dark.std, `demo:2` flags), stage by stage against tests/golden/demo_golden.npz, which was
produced by the UNMODIFIED reference (tests/golden/make_golden.py).

Tolerances are the ones SURVEY.md 8c / BASELINE.json state:
  fBall2 bitwise; neighbour sets equal modulo entries at d2 == fBall2; density rel <= 1e-5;
  |da| <= 1e-5 * sum|terms| (checked as 2e-5*|a| percentile + absolute bound); step-0 survivor set exact;
  FoF partition exact; group count exact (120 -> 68); >= 99.9 % same group; bound masses <= 1e-4.
"""
import numpy as np
import pytest

from conftest import DEMO
from skid_b200 import api
from oracle.refdump import canonical_labels

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sk(demo_input):
    p, ng, nd, ns, _ = demo_input
    s = api.SkidGPU((1.0,) * 3, (0.0,) * 3, bPeriodic=True)
    s.set_particles(p, ng, nd, ns)
    yield s
    s.close()


@pytest.fixture(scope="module")
def demo_run(sk, demo_golden):
    """Run the stages once, keeping every intermediate the tests look at."""
    r = {}
    tau = float(np.float32(DEMO["tau"]))
    fCvg = float(np.float32(0.5 * tau))
    fStep = float(np.float32(0.5 * fCvg))
    r["rho"], r["ball2"] = sk.smDensityInit(64, keep_neighbors=True)
    r["nExtraScat"] = sk.nExtraScat
    r["nbr"], r["nbr_d2"] = sk.neighbors()
    r["nMove"], r["nIttr"] = sk.move(fDensMin=DEMO["fDensMin"], fCvg=fCvg, fStep=fStep, keep_step0=True)
    r["step0"] = sk.step0()
    r["conv_iOrder"], r["conv_r"] = sk.moved()
    r["nGroupFoF"] = sk.kdFoF(tau)
    sk.microstep(5, float(np.float32(0.1 * fStep)))
    r["moved_iOrder"], r["moved_r"] = sk.moved()
    r["fof_grp"], r["fof_cat"] = sk.kdCalcCenter()
    a = 1.0
    fCosmo = a * api.csmExp2Hub(a, float(np.float32(DEMO["H0"])), 1.0, 0.0)
    r["grp"], r["cat"], r["nUnbound"], r["nBefore"] = sk.kdUnbind(1.0, 0.0, fCosmo, api.SPLINE,
                                                                   float(np.float32(2.0 * tau)), False,
                                                                   api.INT_MAX, DEMO["nMembers"])
    r["log"] = list(sk.log)
    return r


def test_primitives_scan_sort(sk):
    rng = np.random.default_rng(1)
    for n in (1, 5, 2047, 2048, 2049, 100003, 3_000_001):
        a = rng.integers(0, 5, n, dtype=np.uint32)
        out = sk.debug_scan(a)
        ref = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32)
        assert np.array_equal(out, ref), n
    for n, bits in ((1, 8), (33, 8), (2048, 16), (5000, 63), (1_000_003, 63), (300_000, 21)):
        k = rng.integers(0, 2 ** min(bits, 62), n, dtype=np.uint64)
        if n > 100:
            k[: n // 2] = k[n // 2: 2 * (n // 2)]  # many duplicates: checks stability
        v = np.arange(n, dtype=np.uint32)
        ks, vs = sk.debug_sort(k, v, bits)
        order = np.argsort(k, kind="stable")
        assert np.array_equal(ks, k[order]), (n, bits)
        assert np.array_equal(vs, v[order]), (n, bits)


def test_knn_ball2_bitwise(demo_run, demo_golden):
    assert np.array_equal(demo_run["ball2"].view(np.uint32), demo_golden["ball2"].view(np.uint32))


def test_knn_neighbor_sets(demo_run, demo_golden):
    sample = demo_golden["knn_sample"]
    ref_nbr, ref_d2 = demo_golden["knn_nbr"], demo_golden["knn_d2"]
    nbr, d2, ball2 = demo_run["nbr"], demo_run["nbr_d2"], demo_run["ball2"]
    for row, i in enumerate(sample):
        inside_ref = set(ref_nbr[row][ref_d2[row] < ball2[i]].tolist())
        inside = set(nbr[i][d2[i] < ball2[i]].tolist())
        assert inside == inside_ref, i
        # and the distances themselves are bit-identical
        a = np.sort(d2[i])
        b = np.sort(ref_d2[row])
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), i
        assert i in inside or ball2[i] == 0


def test_density(demo_run, demo_golden):
    ref = demo_golden["density"]
    rel = np.abs(demo_run["rho"].astype(np.float64) - ref) / ref
    assert rel.max() <= 1e-5, rel.max()


def test_replicas_and_log(demo_run, demo_golden):
    assert demo_run["nExtraScat"] == int(demo_golden["nExtraScat"])
    ittr = [(it, na, ns) for (kind, it, na, ns) in demo_run["log"] if kind == 0]
    ref = [tuple(int(v) for v in row) for row in demo_golden["ittr"]]
    assert ittr[0] == ref[0], (ittr[0], ref[0])          # Ittr:0 nActive:12190 nScatter:20891
    assert demo_run["nMove"] == ref[0][1]
    # the per-block (nActive, nScatter) trace: positions carry float32 noise (the reference's own
    # noise floor is 2e-5), so later blocks may differ by a few movers
    assert abs(len(ittr) - len(ref)) <= 1
    m = min(len(ittr), len(ref))
    na = np.array([r[1] for r in ittr[:m]], float)
    nb = np.array([r[1] for r in ref[:m]], float)
    assert np.all(np.abs(na - nb) <= 0.01 * demo_run["nMove"])
    micro = [ns for (kind, it, na_, ns) in demo_run["log"] if kind == 1]
    assert len(micro) == 5


def test_step0_gradient_and_initial_cut(demo_run, demo_golden, demo_input):
    iord, a, alive = demo_run["step0"]
    ref_a = np.zeros((len(demo_golden["ball2"]), 3), np.float64)
    ref_a[demo_golden["step0_iOrder"]] = demo_golden["step0_a"]
    mine = np.zeros_like(ref_a)
    mine[iord] = a
    assert sorted(iord.tolist()) == sorted(demo_golden["step0_iOrder"].tolist())
    da = np.linalg.norm(mine - ref_a, axis=1)[iord]
    na = np.linalg.norm(ref_a, axis=1)[iord]
    # the difference is float32 summation order (gather form here, scatter form in the reference), so it
    # is bounded by the summed MAGNITUDES of a mover's terms, not by |a| (which cancels): |da| <= 1e-5 sum|terms|
    from oracle import orc
    p = demo_input[0]
    src, rp = orc.replicas(p["r"], demo_golden["ball2"], 1.0)
    eidx = np.concatenate([np.arange(len(p)), src])
    sabs = orc.gradient_abs(np.concatenate([p["r"], rp]), demo_golden["ball2"][eidx], p["fMass"][eidx], p["r"][iord])
    rel = da / na
    print("step-0 gradient |da|/|a| percentiles 50/90/99/99.9/max:",
          [float("%.3g" % v) for v in np.percentile(rel, [50, 90, 99, 99.9, 100])], "max |da|/sum|terms|: %.3g" % (da / sabs).max())
    assert np.all(da <= 1e-5 * sabs), float((da / sabs).max())
    assert np.percentile(rel, 99) < 1e-5
    # survivors of the initial cut: exact set of originals (20 409 on the demo)
    ref_alive = np.unpackbits(demo_golden["step0_alive"])[: len(alive)]
    assert np.array_equal(alive, ref_alive)


def test_fof_partition_given_reference_positions(sk, demo_golden, demo_run):
    """FoF on the reference's own converged positions must give exactly its partition."""
    # compare on our converged positions against a brute-force check of the reference partition instead:
    # our positions differ from the reference by float noise (<= 2e-5 * fStep), which can not change
    # a clean partition; so the partitions over movers must be identical up to relabelling.
    ref_lab = np.zeros(len(demo_golden["ball2"]), np.int64)
    ref_lab[demo_golden["fof_iOrder"]] = demo_golden["fof_group"]
    mine = demo_run["fof_grp"]
    assert demo_run["nGroupFoF"] - 1 == int(demo_golden["fof_group"].max()) == 120
    assert np.array_equal(canonical_labels(mine), canonical_labels(ref_lab))


def test_centers_catalogue(demo_run, demo_golden):
    cat, ref = demo_run["fof_cat"], demo_golden["ub0_cat"]
    grp_ref = demo_golden["ub0_grp"]
    # map reference group ids -> ours through any member
    lab_ref, lab = canonical_labels(grp_ref), canonical_labels(demo_run["fof_grp"])
    assert np.array_equal(lab_ref, lab)
    first = {}
    for i in np.nonzero(grp_ref)[0]:
        first.setdefault(int(grp_ref[i]), i)
    for g_ref, i in first.items():
        g = int(demo_run["fof_grp"][i])
        assert cat["nMembers"][g] == ref["nMembers"][g_ref]
        assert abs(cat["fMass"][g] - ref["fMass"][g_ref]) <= 1e-5 * ref["fMass"][g_ref]
        d = cat["rCenter"][g] - ref["rCenter"][g_ref]
        d -= np.round(d)
        assert np.abs(d).max() < 2e-5, (g, d)
        assert np.allclose(cat["vcm"][g], ref["vcm"][g_ref], rtol=1e-4, atol=1e-6)


def test_moved_positions(demo_run, demo_golden, demo_input):
    p = demo_input[0]
    iord, r = demo_run["moved_iOrder"], demo_run["moved_r"]
    d = r.astype(np.float64) - p["r"][iord]
    d -= np.round(d)
    ref = demo_golden["ray"][iord]
    err = np.abs(d - ref).max(axis=1)
    # the reference's own noise floor for moved positions is 2.2e-5 (Order() re-sort);
    # converged movers hop around their peak with steps of fStep = 2.25e-4, so the worst case of a
    # different float32 summation order is one hop
    assert np.percentile(err, 99) < 5e-6
    assert np.percentile(err, 99.9) < 2e-5
    assert err.max() < 2.25e-4


def test_unbind_end_to_end(demo_run, demo_golden):
    assert demo_run["nBefore"] == int(demo_golden["nGroupBefore"]) == 120
    ng = len(demo_run["cat"]) - 1
    assert ng == int(demo_golden["nGroup"]) == 68
    assert abs(demo_run["nUnbound"] - int(demo_golden["nUnbound"])) <= 20  # 4134; last-particle E~0 cases
    lab, ref = canonical_labels(demo_run["grp"]), canonical_labels(demo_golden["grp"])
    same = np.mean(lab == ref)
    assert same >= 0.999, same
    # bound masses within 1e-4 for groups whose membership is identical
    gm = {}
    for g in range(1, ng + 1):
        members = np.nonzero(demo_run["grp"] == g)[0]
        gm[members.min()] = (demo_run["cat"]["fMass"][g], len(members))
    ref_grp = demo_golden["grp"]
    nmatch = 0
    for g in range(1, 69):
        members = np.nonzero(ref_grp == g)[0]
        k = members.min()
        if k in gm and gm[k][1] == len(members):
            assert abs(gm[k][0] - demo_golden["gtp_mass"][g - 1]) <= 1e-4 * demo_golden["gtp_mass"][g - 1]
            nmatch += 1
    assert nmatch >= 60
