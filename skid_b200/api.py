"""ctypes binding of libskidgpu.so (include/skidgpu.h) + the main.c stage script in Python.

This is plumbing for tests and bench.py: the product host driver is host/skid_main.c (C, same
flags as the reference's main.c).  `SkidGPU` methods are named after the reference stage calls
they replace (main.c:347-495) so the parity tests read like the reference's own main().
There is no CPU fallback: loading fails loudly if the CUDA library is missing.
"""
import ctypes as C
import math
import os

import numpy as np

from .tipsy import PINIT_DTYPE, PGROUP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libskidgpu.so")

FLT_MAX = float(np.finfo(np.float32).max)
INT_MAX = 2 ** 31 - 1
DARK, GAS, STAR = 1, 2, 4
PLUMMER, SPLINE = 1, 2

# skidgpu_stat_row (include/skidgpu.h)
STAT_ROW_DTYPE = np.dtype([("nMembers", "<i4"), ("fTotMass", "<f4"), ("fGasMass", "<f4"), ("fStarMass", "<f4"),
                           ("fVcirc", "<f4"), ("fmVcirc", "<f4"), ("flVcirc", "<f4"), ("fRVmax", "<f4"),
                           ("fRhmass", "<f4"), ("fRouter2", "<f4"), ("fVdispSum", "<f4")])

LOG_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)

_lib = None

# every symbol include/skidgpu.h declares
EXPORTS = [
    "skidgpu_create", "skidgpu_destroy", "skidgpu_last_error", "skidgpu_reserve", "skidgpu_set_shard", "skidgpu_set_particles",
    "skidgpu_set_particles_dev", "skidgpu_set_soft", "skidgpu_density", "skidgpu_keep_neighbors",
    "skidgpu_get_neighbors", "skidgpu_move", "skidgpu_keep_step0", "skidgpu_get_step0", "skidgpu_fof",
    "skidgpu_microstep", "skidgpu_get_moved", "skidgpu_centers", "skidgpu_set_groups",
    "skidgpu_unbind", "skidgpu_stats", "skidgpu_stage_ms", "skidgpu_counter", "skidgpu_debug_sort", "skidgpu_debug_scan",
    "skidgpu_kernel_ms", "skidgpu_stream", "skidgpu_set_reduce_cb", "skidgpu_comm_unique_id", "skidgpu_comm_init",
    "skidgpu_comm_bytes", "skidgpu_set_profile", "skidgpu_debug_move_kernel",
]
UNIQUE_ID_BYTES = 128


def load_library():
    """dlopen libskidgpu.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `make lib` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    P = C.POINTER
    lib.skidgpu_create.argtypes = [P(vp), i, P(f), P(f), i, i]
    lib.skidgpu_destroy.argtypes = [vp]
    lib.skidgpu_destroy.restype = None
    lib.skidgpu_last_error.argtypes = [vp]
    lib.skidgpu_last_error.restype = C.c_char_p
    lib.skidgpu_reserve.argtypes = [vp, C.c_ulonglong]
    lib.skidgpu_set_shard.argtypes = [vp, i, i]
    lib.skidgpu_set_particles.argtypes = [vp, vp, i, i, i, i]
    lib.skidgpu_set_particles_dev.argtypes = [vp] + [vp] * 9 + [i, i, i, i]
    lib.skidgpu_set_soft.argtypes = [vp, f]
    lib.skidgpu_density.argtypes = [vp, i, i, i, vp, vp, P(i)]
    lib.skidgpu_keep_neighbors.argtypes = [vp, i]
    lib.skidgpu_get_neighbors.argtypes = [vp, vp, vp]
    lib.skidgpu_move.argtypes = [vp, f, f, f, f, f, i, i, LOG_CB, vp, P(i), P(i)]
    lib.skidgpu_keep_step0.argtypes = [vp, i]
    lib.skidgpu_get_step0.argtypes = [vp, vp, vp, vp]
    lib.skidgpu_fof.argtypes = [vp, f, P(i)]
    lib.skidgpu_microstep.argtypes = [vp, i, f, LOG_CB, vp]
    lib.skidgpu_get_moved.argtypes = [vp, vp, vp]
    lib.skidgpu_centers.argtypes = [vp, vp, vp]
    lib.skidgpu_set_groups.argtypes = [vp, vp, i, vp]
    lib.skidgpu_unbind.argtypes = [vp, f, f, d, i, f, i, i, i, vp, vp, P(i), P(i), P(i)]
    lib.skidgpu_stats.argtypes = [vp, f, f, d, f, f, vp]
    lib.skidgpu_stage_ms.argtypes = [vp, i]
    lib.skidgpu_stage_ms.restype = d
    lib.skidgpu_counter.argtypes = [vp, i]
    lib.skidgpu_counter.restype = C.c_longlong
    lib.skidgpu_kernel_ms.argtypes = [vp, i, P(i)]
    lib.skidgpu_kernel_ms.restype = d
    lib.skidgpu_stream.argtypes = [vp]
    lib.skidgpu_stream.restype = vp
    lib.skidgpu_set_reduce_cb.argtypes = [vp, vp, vp]
    lib.skidgpu_comm_unique_id.argtypes = [vp]
    lib.skidgpu_comm_init.argtypes = [vp, vp, i, i]
    lib.skidgpu_comm_bytes.argtypes = [vp, P(C.c_longlong)]
    lib.skidgpu_comm_bytes.restype = C.c_longlong
    lib.skidgpu_set_profile.argtypes = [vp, i]
    lib.skidgpu_debug_move_kernel.argtypes = [vp, i]
    lib.skidgpu_debug_sort.argtypes = [vp, vp, vp, C.c_longlong, i]
    lib.skidgpu_debug_scan.argtypes = [vp, vp, vp, C.c_longlong]
    _lib = lib
    return lib


def csmExp2Hub(dExp, H0, Omega0, Lambda, OmegaRad=0.0, Quintess=0.0):
    """H(a) — the one cosmology scalar the hot path needs (reference cosmo.c:46-58); stays on the host."""
    curve = 1.0 - Omega0 - Lambda - OmegaRad - Quintess
    return H0 * math.sqrt(Omega0 * dExp + curve * dExp * dExp + OmegaRad + Quintess * dExp * dExp * math.sqrt(dExp)
                          + Lambda * dExp ** 4) / (dExp * dExp)


class SkidError(RuntimeError):
    pass


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 calls this, every rank passes it to SkidGPU.comm_init)."""
    lib = load_library()
    buf = (C.c_char * UNIQUE_ID_BYTES)()
    if lib.skidgpu_comm_unique_id(C.cast(buf, C.c_void_p)) != 0:
        raise SkidError(lib.skidgpu_last_error(None).decode())
    return bytes(buf.raw)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SkidGPU:
    """One context = one GPU.  Mirrors kdInit ... kdFinish."""

    def __init__(self, fPeriod=(FLT_MAX,) * 3, fCenter=(0.0, 0.0, 0.0), bPeriodic=False, device=0, bDiag=False):
        self.lib = load_library()
        self.h = C.c_void_p()
        per = (C.c_float * 3)(*[float(np.float32(v)) for v in fPeriod])
        cen = (C.c_float * 3)(*[float(np.float32(v)) for v in fCenter])
        if self.lib.skidgpu_create(C.byref(self.h), device, per, cen, int(bPeriodic), int(bDiag)) != 0:
            raise SkidError(self.lib.skidgpu_last_error(None).decode())
        self.n = 0
        self.nSmooth = 0
        self.nMove = 0
        self.nGroup = 0
        self.log = []
        self._cb = LOG_CB(self._on_log)

    def _on_log(self, user, kind, it, nActive, nScatter):
        self.log.append((kind, it, nActive, nScatter))

    def _ck(self, rc):
        if rc != 0:
            raise SkidError(self.lib.skidgpu_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.skidgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- kdReadTipsy result
    def set_particles(self, pinit, nGas, nDark, nStar):
        pinit = np.ascontiguousarray(pinit, dtype=PINIT_DTYPE)
        self.n = len(pinit)
        self._ck(self.lib.skidgpu_set_particles(self.h, _ptr(pinit), self.n, nGas, nDark, nStar))

    def set_particles_dev(self, ptrs, n, nGas, nDark, nStar):
        """ptrs: 9 raw device pointers (x y z vx vy vz mass soft temp)."""
        self.n = n
        self._ck(self.lib.skidgpu_set_particles_dev(self.h, *[C.c_void_p(p) for p in ptrs], n, nGas, nDark, nStar))

    def set_shard(self, rank, nranks):
        self._ck(self.lib.skidgpu_set_shard(self.h, rank, nranks))

    def set_reduce_cb(self, cfunc):
        """cfunc: a parallel.REDUCE_CB instance (kept alive by the caller) or None."""
        self._reduce_cb = cfunc
        self._ck(self.lib.skidgpu_set_reduce_cb(self.h, C.cast(cfunc, C.c_void_p) if cfunc else None, None))

    def comm_init(self, unique_id, rank, nranks):
        """NCCL communicator of this context (collective over all ranks); unique_id: 128 bytes from comm_unique_id()
        of rank 0, handed over by any transport (torch.distributed broadcast, a file, MPI ...)."""
        buf = (C.c_char * UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.skidgpu_comm_init(self.h, C.cast(buf, C.c_void_p), rank, nranks))

    def comm_bytes(self):
        nc = C.c_longlong(0)
        b = self.lib.skidgpu_comm_bytes(self.h, C.byref(nc))
        return int(b), int(nc.value)

    def set_profile(self, on=True):
        self._ck(self.lib.skidgpu_set_profile(self.h, int(on)))

    def debug_move_kernel(self, which):
        self._ck(self.lib.skidgpu_debug_move_kernel(self.h, which))

    def kdSetSoft(self, fEps):
        self._ck(self.lib.skidgpu_set_soft(self.h, fEps))

    # ---- kdScatterActive + kdBuildTree + smInit + smDensityInit
    def smDensityInit(self, nSmooth=64, bGasAndDark=False, bGasOnly=False, want_arrays=True, keep_neighbors=False):
        self.nSmooth = nSmooth
        self._ck(self.lib.skidgpu_keep_neighbors(self.h, int(keep_neighbors)))
        rho = np.zeros(self.n, np.float32) if want_arrays else None
        b2 = np.zeros(self.n, np.float32) if want_arrays else None
        nx = C.c_int(0)
        self._ck(self.lib.skidgpu_density(self.h, nSmooth, int(bGasAndDark), int(bGasOnly), _ptr(rho), _ptr(b2),
                                          C.byref(nx)))
        self.nExtraScat = nx.value
        return rho, b2

    def neighbors(self):
        nbr = np.empty((self.n, self.nSmooth), np.int32)
        d2 = np.empty((self.n, self.nSmooth), np.float32)
        self._ck(self.lib.skidgpu_get_neighbors(self.h, _ptr(nbr), _ptr(d2)))
        return nbr, d2

    # ---- kdInitMove + flow loop
    def move(self, fDensMin=0.0, fTempMax=FLT_MAX, fMassMax=FLT_MAX, fCvg=0.0, fStep=0.0, bForceInitialCut=False,
             bNoPrune=False, keep_step0=False):
        self._ck(self.lib.skidgpu_keep_step0(self.h, int(keep_step0)))
        nm, ni = C.c_int(0), C.c_int(0)
        self._ck(self.lib.skidgpu_move(self.h, fDensMin, fTempMax, fMassMax, fCvg, fStep, int(bForceInitialCut),
                                       int(bNoPrune), self._cb, None, C.byref(nm), C.byref(ni)))
        self.nMove = nm.value
        return nm.value, ni.value

    def step0(self):
        iord = np.empty(self.nMove, np.int32)
        a = np.empty((self.nMove, 3), np.float32)
        alive = np.zeros(self.n, np.uint8)
        self._ck(self.lib.skidgpu_get_step0(self.h, _ptr(iord), _ptr(a), _ptr(alive)))
        return iord, a, alive

    def kdFoF(self, fTau):
        ng = C.c_int(0)
        self._ck(self.lib.skidgpu_fof(self.h, fTau, C.byref(ng)))
        self.nGroup = ng.value
        return ng.value

    def microstep(self, nSteps, fStep):
        self._ck(self.lib.skidgpu_microstep(self.h, nSteps, fStep, self._cb, None))

    def moved(self):
        iord = np.empty(self.nMove, np.int32)
        r = np.empty((self.nMove, 3), np.float32)
        self._ck(self.lib.skidgpu_get_moved(self.h, _ptr(iord), _ptr(r)))
        return iord, r

    def kdCalcCenter(self, fetch=True):
        if not fetch:
            self._ck(self.lib.skidgpu_centers(self.h, None, None))
            return None, None
        grp = np.empty(self.n, np.int32)
        cat = np.zeros(self.nGroup, PGROUP_DTYPE)
        self._ck(self.lib.skidgpu_centers(self.h, _ptr(grp), _ptr(cat)))
        return grp, cat

    def set_groups(self, piGroup, nGroup, centres=None):
        piGroup = np.ascontiguousarray(piGroup, np.int32)
        if centres is not None:
            centres = np.ascontiguousarray(centres, PGROUP_DTYPE)
        self.nGroup = nGroup
        self._ck(self.lib.skidgpu_set_groups(self.h, _ptr(piGroup), nGroup, _ptr(centres)))

    def kdUnbind(self, G=1.0, z=0.0, fCosmo=0.0, iSoftType=SPLINE, fScoop=0.0, bNoUnbind=False, nMaxMembers=INT_MAX,
                 nMinMembers=8, fetch=True, out_grp=None, out_cat=None):
        """fetch=False leaves labels and catalogue on the device (returned as None): counters only.
        out_grp / out_cat: caller-owned result arrays (e.g. pinned host memory) of at least n / nGroup entries."""
        grp = cat = None
        if fetch:
            grp = np.empty(self.n, np.int32) if out_grp is None else out_grp[:self.n]
            cat = np.zeros(max(self.nGroup, 1), PGROUP_DTYPE) if out_cat is None else out_cat[:max(self.nGroup, 1)]
        ng, nu, nb = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.skidgpu_unbind(self.h, G, z, fCosmo, iSoftType, fScoop, int(bNoUnbind), nMaxMembers,
                                         nMinMembers, _ptr(grp), _ptr(cat), C.byref(ng), C.byref(nu), C.byref(nb)))
        self.nGroup = ng.value
        return grp, (cat[:ng.value] if fetch else None), nu.value, nb.value

    def kdOutStats(self, G=1.0, z=0.0, fExpHub=0.0, fDensMin=0.0, fTempMax=FLT_MAX):
        """Accumulator rows behind the .stat file (kd.c:1703-1839), one per final group (row 0 unused)."""
        rows = np.zeros(max(self.nGroup, 1), STAT_ROW_DTYPE)
        self._ck(self.lib.skidgpu_stats(self.h, G, z, fExpHub, fDensMin, fTempMax, _ptr(rows)))
        return rows

    def debug_sort(self, keys, vals, bits):
        keys = np.ascontiguousarray(keys, np.uint64).copy()
        vals = np.ascontiguousarray(vals, np.uint32).copy()
        self._ck(self.lib.skidgpu_debug_sort(self.h, _ptr(keys), _ptr(vals), len(keys), bits))
        return keys, vals

    def debug_scan(self, a):
        a = np.ascontiguousarray(a, np.uint32)
        out = np.empty(len(a) + 1, np.uint32)
        self._ck(self.lib.skidgpu_debug_scan(self.h, _ptr(a), _ptr(out), len(a)))
        return out

    def stage_ms(self):
        names = ["density", "move", "fof", "microstep", "centers", "unbind"]
        return {k: self.lib.skidgpu_stage_ms(self.h, j) for j, k in enumerate(names)}

    def kernel_ms(self, which):
        nl = C.c_int(0)
        ms = self.lib.skidgpu_kernel_ms(self.h, which, C.byref(nl))
        return ms, nl.value

    def stream(self):
        return self.lib.skidgpu_stream(self.h)

    def counter(self, which):
        return int(self.lib.skidgpu_counter(self.h, which))


def run_skid(pinit, nGas, nDark, nStar, tau, nSmooth=64, fDensMin=0.0, fTempMax=FLT_MAX, fMassMax=FLT_MAX,
             fCvg=None, fScoop=None, nMembers=8, nMaxMembers=INT_MAX, bNoUnbind=False, bGasAndDark=False,
             bGasOnly=False, bForceInitialCut=False, bNoPrune=False, period=None, center=(0.0, 0.0, 0.0),
             z=0.0, Omega0=1.0, Lambda=0.0, Quintess=0.0, G=1.0, H0=0.0, iSoftType=SPLINE, fEps=None, device=0,
             want_arrays=True, ctx=None, want_stats=False, move_kernel=0):
    """The stage script of main.c:343-471 on one GPU.  Returns a dict of results."""
    tau = float(np.float32(tau))
    if fCvg is None:
        fCvg = float(np.float32(0.5 * tau))        # main.c:343
    if fScoop is None:
        fScoop = float(np.float32(2.0 * tau))      # main.c:344
    fStep = float(np.float32(0.5 * fCvg))          # main.c:345
    per = (FLT_MAX,) * 3 if period is None else (period,) * 3
    own = ctx is None
    sk = ctx or SkidGPU(per, center, bPeriodic=period is not None, device=device)
    out = {}
    try:
        sk.log = []
        sk.debug_move_kernel(move_kernel)
        sk.set_particles(pinit, nGas, nDark, nStar)
        rho, b2 = sk.smDensityInit(nSmooth, bGasAndDark, bGasOnly, want_arrays=want_arrays)
        out["rho"], out["ball2"], out["nExtraScat"] = rho, b2, sk.nExtraScat
        nMove, nIttr = sk.move(fDensMin, fTempMax, fMassMax, fCvg, fStep, bForceInitialCut, bNoPrune)
        out["nMove"], out["nIttr"] = nMove, nIttr
        nGroupFoF = sk.kdFoF(tau)
        sk.microstep(5, float(np.float32(0.1 * fStep)))  # main.c:14,435
        if want_arrays:
            out["moved_iOrder"], out["moved_r"] = sk.moved()
        grp0, cat0 = sk.kdCalcCenter()
        out["fof_grp"], out["fof_cat"] = grp0, cat0
        if fEps is not None:
            sk.kdSetSoft(fEps)
        f32 = lambda v: float(np.float32(v))        # main.c:56-58: z, Omega0, G, H0, Lambda, Q are floats
        z, G = f32(z), f32(G)
        a32 = f32(1.0 / (1.0 + z))                  # kd.c:1317: fShift is a float
        fCosmo = a32 * csmExp2Hub(a32, f32(H0), f32(Omega0), f32(Lambda), 0.0, f32(Quintess))
        grp, cat, nUnbound, nBefore = sk.kdUnbind(G, z, fCosmo, iSoftType, fScoop, bNoUnbind, nMaxMembers, nMembers)
        if want_stats:                               # main.c:482-484
            out["stat_rows"] = sk.kdOutStats(G, z, fCosmo, fDensMin, fTempMax)
            out["fCosmo"] = fCosmo
        out.update(grp=grp, cat=cat, nUnbound=nUnbound, nGroupBefore=nBefore, nGroup=sk.nGroup - 1,
                   log=list(sk.log), stage_ms=sk.stage_ms(), mover_steps=sk.counter(1), launches=sk.counter(0))
    finally:
        if own:
            sk.close()
    return out
