"""Pins the oracle (oracle/skid_oracle.c, the CPU restatement) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py), stage by stage on config 1 (the demo).
CPU only; the whole file runs in about a minute."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import DEMO, GOLDEN, ROOT

sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def orc():
    subprocess.run(["make", "-C", ROOT, "oracle/liboracle.so"], check=True, capture_output=True)
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="module")
def knn(orc, demo_input):
    p = demo_input[0]
    return orc.knn_density(p["r"], p["fMass"], 64, 1.0, want_nbr=True)


def test_oracle_knn_ball2_bitwise(knn, demo_golden):
    ball2 = knn[0]
    assert np.array_equal(ball2.view(np.uint32), demo_golden["ball2"].view(np.uint32))


def test_oracle_knn_neighbor_sets(knn, demo_golden):
    ball2, _, nbr, d2 = knn
    for row, i in enumerate(demo_golden["knn_sample"]):
        ref = set(demo_golden["knn_nbr"][row][demo_golden["knn_d2"][row] < ball2[i]].tolist())
        mine = set(nbr[i][d2[i] < ball2[i]].tolist())
        assert mine == ref, i


def test_oracle_density(knn, demo_golden):
    rho = knn[1].astype(np.float64)
    ref = demo_golden["density"].astype(np.float64)
    assert (np.abs(rho - ref) / ref).max() <= 1e-5


def test_oracle_replicas(orc, demo_input, demo_golden):
    p = demo_input[0]
    src, rp = orc.replicas(p["r"], demo_golden["ball2"], 1.0)
    assert len(src) == int(demo_golden["nExtraScat"]) == 8300


@pytest.fixture(scope="module")
def entities(orc, demo_input, demo_golden):
    p = demo_input[0]
    src, rp = orc.replicas(p["r"], demo_golden["ball2"], 1.0)
    epos = np.concatenate([p["r"], rp])
    idx = np.concatenate([np.arange(len(p)), src])
    return dict(pos=epos, ball2=demo_golden["ball2"][idx], mass=p["fMass"][idx], rho=demo_golden["density"][idx],
                src=idx, nOrig=len(p))


def test_oracle_step0_gradient_and_cut(orc, demo_input, demo_golden, entities):
    p = demo_input[0]
    movers = np.nonzero(demo_golden["density"] >= np.float32(DEMO["fDensMin"]))[0]
    assert len(movers) == len(demo_golden["step0_iOrder"]) == 12190
    acc, touched, fsd = orc.gradient(entities["pos"], entities["ball2"], entities["mass"], entities["rho"],
                                     p["r"][movers])
    ref = np.zeros((len(p), 3))
    ref[demo_golden["step0_iOrder"]] = demo_golden["step0_a"]
    da = np.linalg.norm(acc - ref[movers], axis=1)
    na = np.linalg.norm(ref[movers], axis=1)
    assert np.percentile(da / na, 99) < 2e-5 and (da / na).max() < 5e-4
    # the noise is float32 summation order, bounded by 1e-5 of the summed term magnitudes
    sabs = orc.gradient_abs(entities["pos"], entities["ball2"], entities["mass"], p["r"][movers])
    assert np.all(da <= 1e-5 * sabs), float((da / sabs).max())
    # initial cut + ScatterCut: touched and rho >= fScatDens
    alive = (touched != 0) & (entities["rho"] >= np.float32(fsd))
    n = entities["nOrig"]
    ref_alive = np.unpackbits(demo_golden["step0_alive"])[:n]
    assert np.array_equal(alive[:n].astype(np.uint8), ref_alive)
    assert int(alive[n:].sum()) == int(demo_golden["step0_nReplicaAlive"])
    assert int(alive.sum()) == int(demo_golden["ittr"][0][2])  # nScatter of "Ittr:0" = 20891


def test_oracle_fof_on_reference_positions(orc, demo_golden):
    lab, g = orc.fof(demo_golden["fof_r"], float(np.float32(DEMO["tau"])), 1.0)
    assert g == 120
    from oracle.refdump import canonical_labels
    assert np.array_equal(canonical_labels(lab), canonical_labels(demo_golden["fof_group"]))


def test_oracle_move_loop_then_fof(orc, demo_input, demo_golden, entities):
    """The whole flow loop of the restatement reproduces the reference's Ittr trace and FoF partition."""
    p = demo_input[0]
    movers = np.nonzero(demo_golden["density"] >= np.float32(DEMO["fDensMin"]))[0]
    tau = float(np.float32(DEMO["tau"]))
    fCvg = float(np.float32(0.5 * tau))
    fStep = float(np.float32(0.5 * fCvg))
    r = orc.move_loop(entities["pos"], entities["ball2"], entities["mass"], entities["rho"], p["r"][movers], 1.0,
                      (0, 0, 0), fCvg, fStep, bInitial=True)
    ref = demo_golden["ittr"]
    assert (r["nActive"][0], r["nScatter"][0]) == (ref[0][1], ref[0][2])
    assert abs(r["nIttr"] - len(ref)) <= 1
    m = min(r["nIttr"], len(ref))
    assert np.all(np.abs(r["nActive"][:m] - ref[:m, 1]) <= 0.01 * len(movers))
    lab, g = orc.fof(r["converged"], tau, 1.0)
    assert g == 120
    from oracle.refdump import canonical_labels
    full = np.zeros(len(p), np.int64)
    full[movers] = lab
    ref_full = np.zeros(len(p), np.int64)
    ref_full[demo_golden["fof_iOrder"]] = demo_golden["fof_group"]
    assert np.array_equal(canonical_labels(full), canonical_labels(ref_full))
    # moved positions after the micro steps vs the reference's .ray
    d = r["final"].astype(np.float64) - p["r"][movers]
    d -= np.round(d)
    err = np.abs(d - demo_golden["ray"][movers]).max(axis=1)
    assert np.percentile(err, 99) < 5e-6


def test_oracle_unbind(orc, demo_input, demo_golden):
    """kdUnbind restated: run every pre-unbind group of the reference (dump 'ub0') through the oracle."""
    from skid_b200 import api
    p = demo_input[0]
    grp0, cat0 = demo_golden["ub0_grp"], demo_golden["ub0_cat"]
    grp1, cat1 = demo_golden["ub1_grp"], demo_golden["ub1_cat"]
    tau = float(np.float32(DEMO["tau"]))
    fScoop2 = np.float32(np.float32(2.0 * tau) ** 2)
    fCosmo = float(np.float32(1.0 * api.csmExp2Hub(1.0, float(np.float32(DEMO["H0"])), 1.0, 0.0)))
    loose = np.nonzero(grp0 == 0)[0]
    total_removed, exact_groups = 0, 0
    for g in range(1, len(cat0)):
        mem = np.nonzero(grp0 == g)[0]
        rel = cat0["rel"][g]
        dr = (p["r"][mem] - rel).astype(np.float32)
        dr = np.where(dr > 0.5, dr - 1, dr)
        dr = np.where(dr <= -0.5, dr + 1, dr).astype(np.float32)
        # scoop sources: ungrouped particles within fScoop of the density centre (min image)
        dc = p["r"][loose] - cat0["rCenter"][g]
        dc -= np.round(dc)
        sc = loose[(dc.astype(np.float32) ** 2).sum(axis=1) < fScoop2]
        sr = (p["r"][sc] - rel).astype(np.float32)
        sr = np.where(sr > 0.5, sr - 1, sr)
        sr = np.where(sr <= -0.5, sr + 1, sr).astype(np.float32)
        k, removed, bm, vcm = orc.unbind_group(dr, p["v"][mem], p["fMass"][mem], p["fSoft"][mem], sr, p["fMass"][sc],
                                               p["fSoft"][sc], 1.0, 0.0, fCosmo)
        total_removed += k
        ref_removed = grp1[mem] == 0
        if np.array_equal(removed.astype(bool), ref_removed):
            exact_groups += 1
            if bm > 0:
                assert abs(bm - cat1["fMass"][g]) <= 1e-4 * cat1["fMass"][g]
    assert abs(total_removed - int(demo_golden["nUnbound"])) <= 20     # 4134
    assert exact_groups >= 100                                           # 105/120 identical


def test_stats_match_reference_stat_file(orc, demo_input, demo_golden):
    """orc_stats (kdOutStats kd.c:1703-1839) on the reference's own final groups, centres and velocities
    reproduces the reference's dark.stat line by line (text-identical "%g" columns)."""
    from skid_b200.api import csmExp2Hub
    p, ng, nd, ns, _ = demo_input
    ref_lines = open(os.path.join(GOLDEN, "demo.stat")).read().strip().splitlines()
    ref = np.array([[float(t) for t in ln.split()] for ln in ref_lines])
    nGroup = len(ref_lines) + 1
    rc = np.zeros((nGroup, 3), np.float32)
    vc = np.zeros((nGroup, 3), np.float32)
    rb = np.zeros((nGroup, 3), np.float32)
    rc[1:], vc[1:] = demo_golden["gtp_pos"], demo_golden["gtp_vel"]
    rb[1:] = ref[:, 18:21]
    f32 = lambda v: float(np.float32(v))
    fExp = f32(1.0)
    dExpHub = fExp * csmExp2Hub(fExp, f32(DEMO["H0"]), 1.0, 0.0)
    rows = orc.stats(p["r"], p["v"], p["fMass"], p["fSoft"], p["fTemp"], demo_golden["density"], ng, nd,
                     demo_golden["grp"], nGroup, rc, vc, (1.0, 1.0, 1.0), 1.0, 0.0, dExpHub, DEMO["fDensMin"],
                     3.4028234663852886e38)
    lines = orc.stat_lines(rows, rc, vc, rb)
    assert len(lines) == len(ref_lines) == 68
    same = sum(a.split()[:18] == b.split()[:18] for a, b in zip(lines, ref_lines))
    assert same == 68, [(a, b) for a, b in zip(lines, ref_lines) if a.split()[:18] != b.split()[:18]][:3]


def test_stats_kernel_scheme_model(orc, demo_input, demo_golden):
    """tools/stats_model.py restates the device kernel's scheme (radix-sort keys, chunks of 32 members, carried
    float chains, ballot-style selections) in numpy; on the demo's 68 groups it is bit-identical to orc_stats."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stats_model
    from skid_b200.api import csmExp2Hub
    p, ng, nd, ns, _ = demo_input
    nGroup = 69
    rc = np.zeros((nGroup, 3), np.float32)
    vc = np.zeros((nGroup, 3), np.float32)
    rc[1:], vc[1:] = demo_golden["gtp_pos"], demo_golden["gtp_vel"]
    f32 = np.float32
    dExpHub = csmExp2Hub(1.0, float(f32(DEMO["H0"])), 1.0, 0.0)
    ref = orc.stats(p["r"], p["v"], p["fMass"], p["fSoft"], p["fTemp"], demo_golden["density"], ng, nd,
                    demo_golden["grp"], nGroup, rc, vc, (1.0, 1.0, 1.0), 1.0, 0.0, dExpHub, 170.0, 3.4e38)
    with np.errstate(all="ignore"):
        mod = stats_model.model(p, demo_golden["density"], ng, nd, demo_golden["grp"], nGroup, rc, vc,
                                np.array([0.5, 0.5, 0.5], f32), 1.0, f32(1.0), f32(dExpHub), f32(170.0), f32(3.4e38))
    for f in ref.dtype.names:
        assert np.array_equal(ref[f], mod[f]), f


SYNTH_CASES = {"dark13": ("dark", 1 << 13, 3), "gasdark13": ("gasdark", 1 << 13, 11), "massive14": ("massive", 1 << 14, 9),
               "dark13_nsp": ("dark", 1 << 13, 3)}   # -nsp: golden = the reference with pruning disabled (SKID_NOPRUNE)


@pytest.mark.parametrize("name", sorted(SYNTH_CASES))
def test_oracle_pipeline_matches_reference_on_synthetic_boxes(orc, name):
    """The whole stage script over the restatement (oracle/pipeline.py) against golden results of the unmodified
    reference on synthetic boxes (tests/golden/synth_golden.npz, made by make_synth_golden.py): dark, gas+dark
    (-gd, Lambda cosmology, no potential update after removals) and massive halos with a 4x linking length."""
    from oracle import pipeline
    from oracle.refdump import canonical_labels
    from skid_b200 import synth
    from skid_b200.api import csmExp2Hub
    gold = np.load(os.path.join(GOLDEN, "synth_golden.npz"))
    kind, n, seed = SYNTH_CASES[name]
    snap = synth.make_box(n, seed=seed, kind=kind)
    if name.endswith("_nsp"):
        snap["flags"]["bNoPrune"] = True
    res = pipeline.run_port(snap, csmExp2Hub)
    nIttr, nBefore, nUnbound, nGroup, _ = gold[name + "_log"]
    assert abs(res["nIttr"] - nIttr) <= 1
    assert res["nGroupBefore"] == nBefore
    assert abs(res["nUnbound"] - nUnbound) <= max(2, nUnbound // 50)
    assert abs(res["nGroup"] - nGroup) <= 1
    same = np.mean(canonical_labels(gold[name + "_grp"].astype(np.int64)) == canonical_labels(res["grp"]))
    assert same >= 0.999, same


def test_fullsize_compare_helper():
    """tests/fullsize.py (used by the full-size GPU parity test): the sample of canonical group ids reproduces the
    same-group fraction; renumbering the groups changes nothing; moving 1 % of the particles to another group
    is detected."""
    import fullsize
    g = np.load(os.path.join(GOLDEN, "synth_golden.npz"))
    grp = g["dark14_grp"].astype(np.int32)
    nI, nB, nU, nG, _ = [int(v) for v in g["dark14_log"]]
    sizes = np.sort(np.bincount(grp)[1:])[::-1].astype(np.int32)
    gold = dict(log=g["dark14_log"], stride=4, sample_canon=fullsize.canonical_min_member(grp)[::4], sizes=sizes)
    rep = fullsize.compare(gold, grp, nI, nB, nU, nG)
    assert rep["same_group"] == 1.0
    perm = np.concatenate([[0], 1 + np.random.default_rng(1).permutation(nG)]).astype(np.int32)
    assert fullsize.compare(gold, perm[grp], nI, nB, nU, nG)["same_group"] == 1.0     # numbering is irrelevant
    bad = grp.copy()
    members = np.nonzero(grp > 0)[0]
    bad[members[:: max(1, len(members) // (len(grp) // 100))]] = 0                       # ~1 % of all particles
    with pytest.raises(AssertionError):
        fullsize.compare(gold, bad, nI, nB, nU, nG)


@pytest.mark.parametrize("stars", [False, True])
def test_stats_match_live_reference_with_species(orc, tmp_path, stars):
    """orc_stats against the .stat file of the live reference on a gas+dark(+star) box with -gd, Lambda cosmology,
    z = 0.5 and a temperature cut: gas-mass and star-mass columns, the Hubble term of the velocity dispersion.
    Inputs of orc_stats are the reference's own .grp / .gtp / .den, so every computed column must print the same."""
    from oracle import refdump
    from skid_b200 import synth, tipsy
    from skid_b200.api import csmExp2Hub
    if not refdump.have_ref():
        pytest.skip("oracle/_ref not present")
    n = 1 << 13
    snap = synth.make_box(n, seed=11, kind="gasdark")
    if stars:
        snap["nStar"] = n // 8
        snap["nDark"] -= snap["nStar"]
    f = str(tmp_path / "in.std")
    synth.write_std(snap, f)
    refdump.run_ref(f, snap["ref_args"] + ["-den", "-stats"], str(tmp_path / "ref"))
    ref_lines = open(str(tmp_path / "ref.stat")).read().strip().splitlines()
    assert len(ref_lines) > 5
    ref = np.array([[float(t) for t in ln.split()] for ln in ref_lines])
    grp = tipsy.read_array(str(tmp_path / "ref.grp")).astype(np.int32)
    den = tipsy.read_array(str(tmp_path / "ref.den")).astype(np.float32)
    gtp = tipsy.read_gtp(str(tmp_path / "ref.gtp"), standard=True)
    nGroup = len(ref_lines) + 1
    rc = np.zeros((nGroup, 3), np.float32)
    vc = np.zeros((nGroup, 3), np.float32)
    rb = np.zeros((nGroup, 3), np.float32)
    rc[1:], vc[1:], rb[1:] = gtp["pos"], gtp["vel"], ref[:, 18:21]
    p, fl = snap["pinit"], snap["flags"]
    f32 = lambda v: float(np.float32(v))
    z = f32(fl["z"])
    a = f32(1.0 / (1.0 + z))
    dExpHub = a * csmExp2Hub(a, f32(fl["H0"]), f32(fl["Omega0"]), f32(fl["Lambda"]))
    rows = orc.stats(p["r"], p["v"], p["fMass"], p["fSoft"], p["fTemp"], den, snap["nGas"], snap["nDark"], grp, nGroup,
                     rc, vc, (1.0, 1.0, 1.0), 1.0, z, dExpHub, fl["fDensMin"], fl["fTempMax"])
    lines = orc.stat_lines(rows, rc, vc, rb)
    assert len(lines) == len(ref_lines)
    same = sum(x.split()[:18] == y.split()[:18] for x, y in zip(lines, ref_lines))
    assert same >= len(ref_lines) - 1, [(x, y) for x, y in zip(lines, ref_lines) if x.split()[:18] != y.split()[:18]][:2]
    assert rows["fGasMass"].sum() > 0
    if stars:
        assert rows["fStarMass"].sum() > 0


@pytest.mark.parametrize("name", ["gas_only", "gas_dark", "gas_dark_star_go", "gas_dark_star_gd", "gas_dark_star",
                                  "dark_star", "dark_nonperiodic"])
def test_oracle_pipeline_species_rules(orc, name):
    """Species rules (ScatterCriterion kd.c:600-627, CutCriterion kd.c:555-597) in the restatement's stage script
    against goldens of the unmodified reference (tests/golden/species_golden.npz) for every input type the
    reference distinguishes: gas only; gas + dark without -gd (only gas scatters and moves); gas + dark + stars
    with -go, with -gd (every star moves regardless of its density, 223 iterations) and plain (gas + stars);
    dark + stars (stars only); plus the dark box without -p (not periodic: no replicas, no wrapping).  The goldens
    come from runs WITHOUT -den: with it the reference scatters from the
    wrong particles whenever stars are scatter-active (latent bug, kd.c:1541 after kd.c:669-693)."""
    sys.path.insert(0, GOLDEN)
    from make_species_golden import make_case
    from oracle import pipeline
    from oracle.refdump import canonical_labels
    from skid_b200.api import csmExp2Hub
    gold = np.load(os.path.join(GOLDEN, "species_golden.npz"))
    snap, fl, _ = make_case(name)
    snap["flags"] = fl
    res = pipeline.run_port(snap, csmExp2Hub)
    nIttr, nBefore, nUnbound, nGroup, _, nAct0, _ = [int(v) for v in gold[name + "_log"]]
    assert res["nMove"] == nAct0
    assert abs(res["nIttr"] - nIttr) <= 1 and res["nGroupBefore"] == nBefore and abs(res["nGroup"] - nGroup) <= 1
    assert abs(res["nUnbound"] - nUnbound) <= max(2, nUnbound // 50)
    same = np.mean(canonical_labels(gold[name + "_grp"].astype(np.int64)) == canonical_labels(res["grp"]))
    assert same >= 0.999, same
