"""Readers for the binary dumps written by oracle/_ref/skid_ref_dump (see oracle/build_ref.sh),
and a helper that runs the reference binaries.  TEST INFRASTRUCTURE: imported by tests/,
tests/golden/make_golden.py and bench.py's reference/cpu_baseline arm only.
"""
import os
import subprocess
import time

import numpy as np

from skid_b200.tipsy import PGROUP_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
SKID_REF = os.path.join(REF_DIR, "skid_ref")
SKID_REF_DUMP = os.path.join(REF_DIR, "skid_ref_dump")


def have_ref():
    return os.path.exists(SKID_REF) and os.path.exists(SKID_REF_DUMP)


def run_ref(tipsy_path, args, outprefix, dump=False, noprune=False, cwd=None, timeout=3600):
    """Run the reference on a TIPSY file.  Returns (stdout text, wall seconds)."""
    exe = SKID_REF_DUMP if (dump or noprune) else SKID_REF
    env = dict(os.environ)
    if dump:
        env["SKID_DUMP"] = outprefix + ".dump"
    if noprune:
        env["SKID_NOPRUNE"] = "1"
    cmd = [exe] + [str(a) for a in args] + ["-o", outprefix]
    t0 = time.time()
    with open(tipsy_path, "rb") as fin:
        r = subprocess.run(cmd, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=cwd,
                           timeout=timeout)
    dt = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError(f"reference failed rc={r.returncode}: {r.stderr.decode()[-500:]}")
    return r.stdout.decode(), dt


def parse_log(text):
    """Extract the reference's progress lines (main.c:402,416,436; kd.c:158,1326,1461,1509; smooth1.c:304)."""
    out = dict(ittr=[], micro=[], times={})
    for ln in text.splitlines():
        ln = ln.strip()
        if ln.startswith("Ittr:"):
            p = dict(kv.split(":") for kv in ln.split())
            out["ittr"].append((int(p["Ittr"]), int(p["nActive"]), int(p["nScatter"])))
        elif ln.startswith("Microstep:"):
            p = dict(kv.split(":") for kv in ln.split())
            out["micro"].append((int(p["Microstep"]), int(p["nScatter"])))
        elif ln.startswith("nExtraScat:"):
            out["nExtraScat"] = int(ln.split(":")[1])
        elif ln.startswith("Groups before Unbind:"):
            out["nGroupBefore"] = int(ln.split(":")[1])
        elif ln.startswith("Number of particles Unbound:"):
            out["nUnbound"] = int(ln.split(":")[1])
        elif ln.startswith("Number of Groups:"):
            out["nGroup"] = int(ln.split(":")[1])
        elif ln.startswith(("Initial Density:", "Moving Particles:", "Friends of Friends:", "Microstepping:",
                            "Unbinding:")):
            k, v = ln.split(":")
            out["times"][k.strip()] = float(v)
    return out


def read_knn(path):
    """.knn: nInitActive, nSmooth, then per query (iOrder, fBall2, nSmooth x (iOrder_j, fKey_j))."""
    raw = np.fromfile(path, dtype=np.uint8)
    n, k = np.frombuffer(raw[:8].tobytes(), dtype=np.int32)
    rec = np.dtype([("iOrder", "<i4"), ("fBall2", "<f4"), ("nbr", [("iOrder", "<i4"), ("d2", "<f4")], (int(k),))])
    a = np.frombuffer(raw[8:].tobytes(), dtype=rec)
    assert len(a) == n
    return a


def read_step0(path):
    with open(path, "rb") as f:
        nm = int(np.fromfile(f, np.int32, 1)[0])
        mv = np.fromfile(f, np.dtype([("iOrder", "<i4"), ("a", "<f4", 3)]), nm)
        n, nact = np.fromfile(f, np.int32, 2)
        pi = np.fromfile(f, np.dtype([("iOrder", "<i4"), ("fBall2", "<f4"), ("fDensity", "<f4")]), int(n))
        nx = int(np.fromfile(f, np.int32, 1)[0])
        rep = np.fromfile(f, np.dtype([("iOrder", "<i4"), ("r", "<f4", 3)]), nx)
    return dict(movers=mv, pinit=pi, nInitActive=int(nact), replicas=rep)


def read_fof(path):
    with open(path, "rb") as f:
        n = int(np.fromfile(f, np.int32, 1)[0])
        a = np.fromfile(f, np.dtype([("iOrder", "<i4"), ("r", "<f4", 3), ("group", "<i4")]), n)
    return a


def read_groups(path):
    with open(path, "rb") as f:
        n, ng = np.fromfile(f, np.int32, 2)
        grp = np.fromfile(f, np.int32, int(n))
        cat = np.fromfile(f, PGROUP_DTYPE, int(ng))
    return grp, cat


def canonical_labels(grp):
    """Relabel a partition so groups are numbered by ascending smallest member index (0 stays 0)."""
    grp = np.asarray(grp)
    out = np.zeros_like(grp)
    idx = np.nonzero(grp)[0]
    if len(idx) == 0:
        return out
    g = grp[idx]
    first = np.full(g.max() + 1, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(first, g, idx)
    used = np.nonzero(first != np.iinfo(np.int64).max)[0]
    order = used[np.argsort(first[used], kind="stable")]
    remap = np.zeros(g.max() + 1, np.int64)
    remap[order] = np.arange(1, len(order) + 1)
    out[idx] = remap[g]
    return out
