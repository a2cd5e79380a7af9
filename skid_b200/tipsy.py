"""TIPSY snapshot I/O (numpy), for tests and bench only — the product driver's I/O is host/tipsy_io.c.

Format (reference: tipsydefs.h:6-48, kd.c:16-28,122-222): header {double time; int nbodies, ndim,
nsph, ndark, nstar} (+ one pad int: explicit in the XDR "standard" format kd.c:26, struct tail
padding in the native format), then gas records (12 floats), dark (9), star (11).
Standard = big-endian (XDR), native = host endian.
"""
import numpy as np

GAS_FIELDS = 12   # mass pos3 vel3 rho temp hsmooth metals phi
DARK_FIELDS = 9   # mass pos3 vel3 eps phi
STAR_FIELDS = 11  # mass pos3 vel3 metals tform eps phi

PINIT_DTYPE = np.dtype([("r", "<f4", 3), ("v", "<f4", 3), ("fMass", "<f4"), ("fSoft", "<f4"),
                        ("fTemp", "<f4"), ("fBall2", "<f4"), ("fDensity", "<f4"), ("iOrder", "<i4")])
PGROUP_DTYPE = np.dtype([("rel", "<f4", 3), ("rCenter", "<f4", 3), ("rBound", "<f4", 3), ("vcm", "<f4", 3),
                         ("fMass", "<f4"), ("fRadius", "<f4"), ("nMembers", "<i4"), ("pStart", "<i4"),
                         ("pCurr", "<i4")])
assert PINIT_DTYPE.itemsize == 48 and PGROUP_DTYPE.itemsize == 68


def _hdr_dtype(standard):
    e = ">" if standard else "<"
    return np.dtype([("time", e + "f8"), ("nbodies", e + "i4"), ("ndim", e + "i4"), ("nsph", e + "i4"),
                     ("ndark", e + "i4"), ("nstar", e + "i4"), ("pad", e + "i4")])


def read_tipsy(path_or_bytes, standard=True):
    """Return dict(time, nGas, nDark, nStar, pinit) with pinit a PINIT_DTYPE array as kdReadTipsy fills it."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    h = np.frombuffer(buf, dtype=_hdr_dtype(standard), count=1)[0]
    ng, nd, ns = int(h["nsph"]), int(h["ndark"]), int(h["nstar"])
    n = ng + nd + ns
    e = ">f4" if standard else "<f4"
    off = 32
    p = np.zeros(n, dtype=PINIT_DTYPE)
    p["iOrder"] = np.arange(n, dtype=np.int32)
    if ng:
        a = np.frombuffer(buf, dtype=e, count=ng * GAS_FIELDS, offset=off).reshape(ng, GAS_FIELDS)
        off += ng * GAS_FIELDS * 4
        s = slice(0, ng)
        p["fMass"][s] = a[:, 0]; p["r"][s] = a[:, 1:4]; p["v"][s] = a[:, 4:7]
        p["fTemp"][s] = a[:, 8]; p["fSoft"][s] = a[:, 9]
    if nd:
        a = np.frombuffer(buf, dtype=e, count=nd * DARK_FIELDS, offset=off).reshape(nd, DARK_FIELDS)
        off += nd * DARK_FIELDS * 4
        s = slice(ng, ng + nd)
        p["fMass"][s] = a[:, 0]; p["r"][s] = a[:, 1:4]; p["v"][s] = a[:, 4:7]; p["fSoft"][s] = a[:, 7]
    if ns:
        a = np.frombuffer(buf, dtype=e, count=ns * STAR_FIELDS, offset=off).reshape(ns, STAR_FIELDS)
        s = slice(ng + nd, n)
        p["fMass"][s] = a[:, 0]; p["r"][s] = a[:, 1:4]; p["v"][s] = a[:, 4:7]; p["fSoft"][s] = a[:, 9]
    return dict(time=float(h["time"]), nGas=ng, nDark=nd, nStar=ns, pinit=p)


def write_tipsy(path, time, gas=None, dark=None, star=None, standard=True):
    """gas/dark/star: float32 arrays of shape (n, 12 / 9 / 11) in TIPSY field order."""
    e = ">f4" if standard else "<f4"
    ng = 0 if gas is None else len(gas)
    nd = 0 if dark is None else len(dark)
    ns = 0 if star is None else len(star)
    h = np.zeros(1, dtype=_hdr_dtype(standard))
    h["time"] = time; h["nbodies"] = ng + nd + ns; h["ndim"] = 3
    h["nsph"] = ng; h["ndark"] = nd; h["nstar"] = ns
    with open(path, "wb") as f:
        f.write(h.tobytes())
        for a, nf in ((gas, GAS_FIELDS), (dark, DARK_FIELDS), (star, STAR_FIELDS)):
            if a is not None and len(a):
                a = np.asarray(a, dtype=np.float32)
                assert a.shape[1] == nf
                f.write(a.astype(e).tobytes())


def pinit_to_records(p, nGas, nDark, nStar):
    """Inverse of read_tipsy for writing a PINIT array back to a TIPSY file (rho/metals/phi = 0)."""
    gas = dark = star = None
    if nGas:
        q = p[:nGas]
        gas = np.zeros((nGas, GAS_FIELDS), np.float32)
        gas[:, 0] = q["fMass"]; gas[:, 1:4] = q["r"]; gas[:, 4:7] = q["v"]; gas[:, 8] = q["fTemp"]; gas[:, 9] = q["fSoft"]
    if nDark:
        q = p[nGas:nGas + nDark]
        dark = np.zeros((nDark, DARK_FIELDS), np.float32)
        dark[:, 0] = q["fMass"]; dark[:, 1:4] = q["r"]; dark[:, 4:7] = q["v"]; dark[:, 7] = q["fSoft"]
    if nStar:
        q = p[nGas + nDark:]
        star = np.zeros((nStar, STAR_FIELDS), np.float32)
        star[:, 0] = q["fMass"]; star[:, 1:4] = q["r"]; star[:, 4:7] = q["v"]; star[:, 9] = q["fSoft"]
    return gas, dark, star


def read_array(path, dtype=np.float64):
    """TIPSY ASCII array (.grp/.den): first line N, then N values."""
    a = np.loadtxt(path, dtype=dtype)
    n = int(a[0])
    assert len(a) == n + 1, (len(a), n)
    return a[1:]


def read_vector(path):
    """TIPSY ASCII vector (.ray): N, then all x, all y, all z (kd.c:1570-1606)."""
    a = np.loadtxt(path, dtype=np.float64)
    n = int(a[0])
    assert len(a) == 3 * n + 1
    return a[1:].reshape(3, n).T


def read_gtp(path, standard=True):
    """Group catalogue file written by kdWriteGroup: star records (kd.c:1665-1682)."""
    with open(path, "rb") as f:
        buf = f.read()
    h = np.frombuffer(buf, dtype=_hdr_dtype(standard), count=1)[0]
    ns = int(h["nstar"])
    e = ">f4" if standard else "<f4"
    a = np.frombuffer(buf, dtype=e, count=ns * STAR_FIELDS, offset=32).reshape(ns, STAR_FIELDS)
    return dict(mass=a[:, 0].astype(np.float32), pos=a[:, 1:4].astype(np.float32), vel=a[:, 4:7].astype(np.float32),
                eps=a[:, 9].astype(np.float32), time=float(h["time"]))
