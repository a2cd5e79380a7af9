#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference built by oracle/build_ref.sh.

Runs only in the build container (needs /root/reference/dark.std and oracle/_ref/).  The
demo command line is the reference's own (`demo:2`).  Outputs:

  demo_input.npz   the demo snapshot (config 1 input; the reference ships it as dark.std)
  demo_golden.npz  per-stage results of the reference on it (see keys below)
  demo_stdout.txt  the reference's progress lines

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402
from skid_b200 import tipsy  # noqa: E402

DEMO_ARGS = ["-std", "-tau", "9e-4", "-s", "64", "-d", "170", "-m", "8", "-H", "2.8944", "-p", "1",
             "-ray", "-den", "-stats"]


def main():
    src = "/root/reference/dark.std"
    snap = tipsy.read_tipsy(src, standard=True)
    p = snap["pinit"]
    np.savez_compressed(os.path.join(HERE, "demo_input.npz"), time=snap["time"], nGas=snap["nGas"],
                        nDark=snap["nDark"], nStar=snap["nStar"], mass=p["fMass"], r=p["r"], v=p["v"],
                        soft=p["fSoft"], temp=p["fTemp"])
    with tempfile.TemporaryDirectory() as td:
        pre = os.path.join(td, "dark")
        out, _ = refdump.run_ref(src, DEMO_ARGS, pre, dump=True)
        md5 = {e: hashlib.md5(open(pre + "." + e, "rb").read()).hexdigest() for e in ("grp", "den", "ray", "stat")}
        # sanity: these are the values SURVEY.md 8c lists for the unmodified reference
        assert md5["grp"] == "bbef8709daa0e9d5525981137cf1b02b", md5
        assert md5["den"] == "fc56c1b53a179c87654888d196cf5343", md5
        log = refdump.parse_log(out)
        knn = refdump.read_knn(pre + ".dump.knn")
        st0 = refdump.read_step0(pre + ".dump.step0")
        fof = refdump.read_fof(pre + ".dump.fof")
        grp0, cat0 = refdump.read_groups(pre + ".dump.ub0")
        grp1, cat1 = refdump.read_groups(pre + ".dump.ub1")
        den = tipsy.read_array(pre + ".den").astype(np.float32)
        grp = tipsy.read_array(pre + ".grp").astype(np.int32)
        ray = tipsy.read_vector(pre + ".ray").astype(np.float32)
        gtp = tipsy.read_gtp(pre + ".gtp", standard=True)
        n = len(p)
        ball2 = np.zeros(n, np.float32)
        ball2[knn["iOrder"]] = knn["fBall2"]
        # neighbour lists of a sample of queries (all 32768 x 64 would be 17 MB)
        sample = np.arange(0, n, 61, dtype=np.int32)
        row = np.full(n, -1, np.int64)
        row[knn["iOrder"]] = np.arange(len(knn))
        ks = knn[row[sample]]
        alive = np.zeros(n, np.uint8)
        pi = st0["pinit"]
        alive[pi["iOrder"][: st0["nInitActive"]]] = 1
        np.savez_compressed(
            os.path.join(HERE, "demo_golden.npz"),
            ball2=ball2, density=den,
            knn_sample=sample, knn_nbr=ks["nbr"]["iOrder"].astype(np.int32), knn_d2=ks["nbr"]["d2"].astype(np.float32),
            step0_iOrder=st0["movers"]["iOrder"], step0_a=st0["movers"]["a"],
            step0_alive=np.packbits(alive), step0_nReplicaAlive=len(st0["replicas"]),
            fof_iOrder=fof["iOrder"], fof_r=fof["r"], fof_group=fof["group"],
            ub0_grp=grp0.astype(np.int32), ub0_cat=cat0, ub1_grp=grp1.astype(np.int32), ub1_cat=cat1,
            grp=grp, ray=ray, gtp_mass=gtp["mass"], gtp_pos=gtp["pos"], gtp_vel=gtp["vel"], gtp_eps=gtp["eps"],
            ittr=np.array(log["ittr"], np.int32), micro=np.array(log["micro"], np.int32),
            nExtraScat=log["nExtraScat"], nGroupBefore=log["nGroupBefore"], nUnbound=log["nUnbound"],
            nGroup=log["nGroup"],
            md5_grp=md5["grp"], md5_den=md5["den"], md5_ray=md5["ray"], md5_stat=md5["stat"],
        )
        with open(os.path.join(HERE, "demo_stdout.txt"), "w") as f:
            f.write("\n".join(ln for ln in out.splitlines() if not ln.strip().startswith(
                ("Initial Density", "Moving Particles", "Friends of", "Microstepping", "Unbinding", "SKID CPU"))) + "\n")
        for e in ("stat",):
            with open(pre + "." + e) as fi, open(os.path.join(HERE, "demo." + e), "w") as fo:
                fo.write(fi.read())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
