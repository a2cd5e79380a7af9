"""TEST INFRASTRUCTURE: the reference's stage script (main.c:343-471) over the oracle's C restatement
(oracle/skid_oracle.c) - density, move loop, FoF, micro steps, centres, unbinding, too-small removal.

Used by tests (end-to-end pin of the restatement against the live reference on synthetic boxes) and by
bench.py as the "port" CPU baseline when the compiled reference (oracle/_ref) did not travel.  Never by
the product path.  Restriction of the restatement: the period, if any, is cubic.
"""
import time

import numpy as np

from . import orc


def _wrap(d, L):
    """min-image of float32 offsets the way the reference does it: > L/2 -> -L, <= -L/2 -> +L (kd.c:1341-1355)."""
    d = d.astype(np.float32)
    if not L or L <= 0:          # not periodic (fPeriod = FLT_MAX, main.c:125-128): no image is ever closer
        return d
    h = np.float32(0.5 * L)
    d = np.where(d > h, d - np.float32(L), d).astype(np.float32)
    d = np.where(d <= -h, d + np.float32(L), d).astype(np.float32)
    return d


def run_port(snap, csm_exp2hub):
    """snap: skid_b200.synth.make_box result.  csm_exp2hub: H(a) callable (cosmo.c:46-58, host-side scalar).
    Returns dict(grp, nGroupBefore, nGroup, nUnbound, nIttr, times={stage: seconds})."""
    p = snap["pinit"]
    fl = snap["flags"]
    n = len(p)
    L = float(fl.get("period") or 0.0)      # 0: not periodic
    f32 = lambda v: float(np.float32(v))
    tau = f32(fl["tau"])
    fCvg, fScoop = f32(0.5 * tau), f32(2.0 * tau)          # main.c:343-344
    fStep = f32(0.5 * fCvg)                                # main.c:345
    nGas, nDark, nStar = snap["nGas"], snap["nDark"], snap["nStar"]
    bGD, bGO = fl.get("bGasAndDark", False), fl.get("bGasOnly", False)
    idx_all = np.arange(n)
    is_gas, is_star = idx_all < nGas, idx_all >= nGas + nDark
    is_dark = ~is_gas & ~is_star
    # ScatterCriterion (kd.c:600-627): which particles carry the density field
    if nGas == 0 and nStar == 0:
        active = np.ones(n, bool)                                  # DARK
    elif nStar == 0:
        active = np.ones(n, bool) if bGD else is_gas               # GAS, DARK|GAS
    elif nGas == 0:
        active = is_star                                           # STAR, DARK|STAR
    else:
        active = np.ones(n, bool) if bGD else (is_gas | (is_star & (not bGO)))   # GAS|STAR(|DARK)
    act = np.nonzero(active)[0]
    times = {}
    # ---- stage 1/2: kNN + density over the scatter-active particles (smDensityInit)
    t0 = time.perf_counter()
    ball2 = np.zeros(n, np.float32)
    rho = np.zeros(n, np.float32)
    ball2[act], rho[act] = orc.knn_density(p["r"][act], p["fMass"][act], fl["nSmooth"], L)
    src, rp = orc.replicas(p["r"][act], ball2[act], L)
    src = act[src]
    times["Initial Density"] = time.perf_counter() - t0
    # ---- stage 3: movers (CutCriterion kd.c:555-597) + flow loop + micro steps
    t0 = time.perf_counter()
    dens_ok = rho >= np.float32(fl["fDensMin"])
    gas_rule = is_gas & dens_ok & (p["fTemp"] <= np.float32(fl.get("fTempMax", 3.4e38)))
    if nGas == 0 and nStar == 0:
        move = dens_ok
    elif nStar == 0:
        move = gas_rule | (is_dark & dens_ok & bGD)
    elif nGas == 0:
        move = is_star
    else:
        move = gas_rule | (is_dark & dens_ok & bGD) | (is_star & (not bGO))
    move = move & (p["fMass"] <= np.float32(fl.get("fMassMax", 3.4e38)))
    movers = np.nonzero(move)[0]
    idx = np.concatenate([act, src])
    epos = np.concatenate([p["r"][act], rp]).astype(np.float32)
    bInitial = (nGas == 0 and nStar == 0) or fl.get("bForceInitialCut", False)   # main.c:396
    mv = orc.move_loop(epos, ball2[idx], p["fMass"][idx], rho[idx], p["r"][movers], L, (0, 0, 0), fCvg, fStep,
                       bInitial=bInitial, bNoPrune=fl.get("bNoPrune", False))
    times["Moving Particles"] = time.perf_counter() - t0
    # ---- stage 4: FoF on the converged positions (before the micro steps)
    t0 = time.perf_counter()
    lab, G = orc.fof(mv["converged"], tau, L)
    times["Friends of Friends"] = time.perf_counter() - t0
    grp = np.zeros(n, np.int32)
    grp[movers] = lab
    nGroupBefore = G
    # ---- centres (kdCalcCenter kd.c:1013-1075): rel = a member, centre = rel + mean min-image offset of the
    # moved (post micro step) positions
    t0 = time.perf_counter()
    final = mv["final"]
    order = np.argsort(lab, kind="stable")
    bounds = np.searchsorted(lab[order], np.arange(1, G + 2))
    z = f32(fl.get("z", 0.0))
    a = f32(1.0 / (1.0 + z))
    fCosmo = f32(a * csm_exp2hub(a, f32(fl["H0"]), f32(fl.get("Omega0", 1.0)), f32(fl.get("Lambda", 0.0))))
    # kd.c:1441: potentials are updated after a removal only for pure-dark or pure-star inputs
    bSubPot = (nGas == 0 and nStar == 0) or (nGas == 0 and nDark == 0)
    loose = np.nonzero(grp == 0)[0]
    fScoop2 = np.float32(np.float32(fScoop) ** 2)
    nUnbound = 0
    keep = np.zeros(G + 1, bool)
    for g in range(1, G + 1):
        mm = order[bounds[g - 1]:bounds[g]]          # mover slots of group g
        mem = movers[mm]
        rel = p["r"][mem[-1]]
        off = _wrap(final[mm] - rel, L)
        cen = (rel + off.astype(np.float64).mean(axis=0)).astype(np.float32)
        if L > 0:
            cen = np.where(cen > 0.5 * L, cen - L, cen)
            cen = np.where(cen <= -0.5 * L, cen + L, cen).astype(np.float32)
        dr = _wrap(p["r"][mem] - rel, L)
        dc = _wrap(p["r"][loose] - cen, L)
        sc = loose[(dc ** 2).sum(axis=1) < fScoop2]
        sr = _wrap(p["r"][sc] - rel, L)
        k, removed, bm, vcm = orc.unbind_group(dr, p["v"][mem], p["fMass"][mem], p["fSoft"][mem], sr, p["fMass"][sc],
                                               p["fSoft"][sc], f32(fl.get("G", 1.0)), z, fCosmo,
                                               bNoUnbind=fl.get("bNoUnbind", False), bSubPot=bSubPot)
        nUnbound += k
        grp[mem[removed.astype(bool)]] = 0
        keep[g] = (len(mem) - k) >= fl["nMembers"]   # kdTooSmall kd.c:1251-1293
    times["Unbinding"] = time.perf_counter() - t0
    newid = np.zeros(G + 1, np.int32)
    newid[keep] = np.arange(1, int(keep.sum()) + 1)
    grp = newid[grp]
    return dict(grp=grp, nGroupBefore=nGroupBefore, nGroup=int(keep.sum()), nUnbound=int(nUnbound),
                nIttr=int(mv["nIttr"]), nMove=len(movers), times=times)
