"""ctypes loader for oracle/liboracle.so (the C restatement of the reference algorithm).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg — never from skid_b200/ (the product path has no CPU fallback).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} not built: run `make oracle`")
        L = C.CDLL(LIB)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.orc_knn_density.argtypes = [i, vp, vp, i, f, vp, vp, vp, vp]
        L.orc_knn_density.restype = None
        L.orc_replicas.argtypes = [i, vp, vp, f, vp, i, vp, vp]
        L.orc_replicas.restype = i
        L.orc_gradient.argtypes = [i, vp, vp, vp, vp, vp, i, vp, vp, vp]
        L.orc_gradient.restype = f
        L.orc_gradient_abs.argtypes = [i, vp, vp, vp, i, vp, vp]
        L.orc_gradient_abs.restype = None
        L.orc_move_loop.argtypes = [i, vp, vp, vp, vp, i, vp, f, vp, f, f, i, i, i, vp, vp, i, f, vp]
        L.orc_move_loop.restype = i
        L.orc_fof.argtypes = [i, vp, f, f, vp]
        L.orc_fof.restype = i
        L.orc_unbind_group.argtypes = [i, vp, vp, vp, vp, i, vp, vp, vp, f, f, f, i, i, i, vp, vp, vp]
        L.orc_unbind_group.restype = i
        L.orc_stats.argtypes = [i, vp, vp, vp, vp, vp, vp, i, i, vp, i, vp, vp, vp, f, f, C.c_double, f, f, vp]
        L.orc_stats.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def knn_density(pos, mass, k, period, want_nbr=False):
    pos, mass = _f32(pos), _f32(mass)
    n = len(pos)
    ball2, rho = np.empty(n, np.float32), np.empty(n, np.float32)
    nbr = np.empty((n, k), np.int32) if want_nbr else None
    nd2 = np.empty((n, k), np.float32) if want_nbr else None
    lib().orc_knn_density(n, _p(pos), _p(mass), k, period or 0.0, _p(ball2), _p(rho), _p(nbr), _p(nd2))
    return (ball2, rho, nbr, nd2) if want_nbr else (ball2, rho)


def replicas(pos, ball2, period, center=(0.0, 0.0, 0.0)):
    pos, ball2 = _f32(pos), _f32(ball2)
    cen = _f32(center)
    n = len(pos)
    cnt = lib().orc_replicas(n, _p(pos), _p(ball2), period, _p(cen), 0, None, None)
    src = np.empty(cnt, np.int32)
    rp = np.empty((cnt, 3), np.float32)
    lib().orc_replicas(n, _p(pos), _p(ball2), period, _p(cen), cnt, _p(src), _p(rp))
    return src, rp


def gradient(epos, eball2, emass, erho, mpos, alive=None):
    epos, eball2, emass, erho, mpos = map(_f32, (epos, eball2, emass, erho, mpos))
    ne, nm = len(epos), len(mpos)
    acc = np.empty((nm, 3), np.float32)
    touched = np.empty(ne, np.uint8)
    al = None if alive is None else np.ascontiguousarray(alive, np.uint8)
    fsd = lib().orc_gradient(ne, _p(epos), _p(eball2), _p(emass), _p(erho), _p(al), nm, _p(mpos), _p(acc), _p(touched))
    return acc, touched, float(fsd)


def gradient_abs(epos, eball2, emass, mpos):
    """Per mover, the sum of the magnitudes of its gradient terms (the scale of the float32 summation noise)."""
    epos, eball2, emass, mpos = map(_f32, (epos, eball2, emass, mpos))
    sabs = np.empty(len(mpos), np.float64)
    lib().orc_gradient_abs(len(epos), _p(epos), _p(eball2), _p(emass), len(mpos), _p(mpos), _p(sabs))
    return sabs


def move_loop(epos, eball2, emass, erho, mpos, period, center, fCvg, fStep, bInitial, bNoPrune=False, nMicro=5,
              fMicroStep=None):
    epos, eball2, emass = map(_f32, (epos, eball2, emass))
    erho = _f32(erho).copy()
    mpos = _f32(mpos).copy()
    cen = _f32(center)
    maxlog = 4096
    la, ls = np.zeros(maxlog, np.int32), np.zeros(maxlog, np.int32)
    at_fof = np.empty_like(mpos)
    if fMicroStep is None:
        fMicroStep = float(np.float32(0.1 * fStep))
    n = lib().orc_move_loop(len(epos), _p(epos), _p(eball2), _p(emass), _p(erho), len(mpos), _p(mpos), period or 0.0,
                            _p(cen), fCvg, fStep, int(bInitial), int(bNoPrune), maxlog, _p(la), _p(ls), nMicro,
                            fMicroStep, _p(at_fof))
    return dict(nIttr=n, nActive=la[:n].copy(), nScatter=ls[:n].copy(), converged=at_fof, final=mpos)


def fof(pos, tau, period):
    pos = _f32(pos)
    lab = np.empty(len(pos), np.int32)
    g = lib().orc_fof(len(pos), _p(pos), tau, period or 0.0, _p(lab))
    return lab, g


def unbind_group(r, v, mass, soft, sr, smass, ssoft, G, z, fCosmo, iSoftType=2, bNoUnbind=False, bSubPot=True):
    r, v, mass, soft = map(_f32, (r, v, mass, soft))
    sr, smass, ssoft = _f32(sr).reshape(-1, 3), _f32(smass), _f32(ssoft)
    n = len(r)
    removed = np.zeros(n, np.uint8)
    bm = C.c_double(0)
    vcm = np.zeros(3, np.float64)
    k = lib().orc_unbind_group(n, _p(r), _p(v), _p(mass), _p(soft), len(sr), _p(sr), _p(smass), _p(ssoft), G, z,
                               fCosmo, iSoftType, int(bNoUnbind), int(bSubPot), _p(removed), C.byref(bm), _p(vcm))
    return k, removed, bm.value, vcm


STAT_ROW_DTYPE = np.dtype([("nMembers", "<i4"), ("fTotMass", "<f4"), ("fGasMass", "<f4"), ("fStarMass", "<f4"),
                           ("fVcirc", "<f4"), ("fmVcirc", "<f4"), ("flVcirc", "<f4"), ("fRVmax", "<f4"),
                           ("fRhmass", "<f4"), ("fRouter2", "<f4"), ("fVdispSum", "<f4")])


def stats(pos, vel, mass, soft, temp, rho, nGas, nDark, grp, nGroup, rCenter, vcm, period, G, z, dExpHub,
          fDensMin, fTempMax):
    """kdOutStats accumulators per group (row 0 unused); rCenter/vcm: (nGroup, 3)."""
    pos, vel, mass, soft, temp, rho = map(_f32, (pos, vel, mass, soft, temp, rho))
    rc, vc = _f32(rCenter), _f32(vcm)
    grp = np.ascontiguousarray(grp, np.int32)
    per = _f32(period)
    rows = np.zeros(nGroup, STAT_ROW_DTYPE)
    lib().orc_stats(len(pos), _p(pos), _p(vel), _p(mass), _p(soft), _p(temp), _p(rho), nGas, nDark, _p(grp), nGroup,
                    _p(rc), _p(vc), _p(per), G, z, dExpHub, fDensMin, fTempMax, _p(rows))
    return rows


def stat_lines(rows, rCenter, vcm, rBound):
    """The text of a .stat file (kd.c:1811-1833) from accumulator rows; C's %g == Python's %g."""
    out = []
    for ig in range(1, len(rows)):
        r = rows[ig]
        n = int(r["nMembers"])
        if n <= 0:
            continue
        f32 = np.float32
        vals = [float(r["fTotMass"]), float(r["fGasMass"]), float(r["fStarMass"]),
                float(np.sqrt(np.float64(r["fVcirc"]))), float(np.sqrt(np.float64(r["fmVcirc"]))),
                float(np.sqrt(np.float64(r["flVcirc"]))), float(r["fRVmax"]), float(r["fRhmass"]),
                float(np.sqrt(np.float64(r["fRouter2"]))),
                float(f32(np.sqrt(np.float64(r["fVdispSum"]) / (3.0 * n))))]
        vals += [float(v) for v in rCenter[ig]] + [float(v) for v in vcm[ig]] + [float(v) for v in rBound[ig]]
        out.append("%d %d " % (ig, n) + " ".join("%g" % v for v in vals))
    return out
