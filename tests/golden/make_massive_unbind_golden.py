#!/usr/bin/env python
"""Golden results of the UNMODIFIED reference for the unbinding of MASSIVE groups (>= 20 000 members; BASELINE
configs[4]), which the full-size C5 golden had to skip with -maxgroup 20000 (kd.c:1330).

The serial reference needs O(n^2) pair evaluations per group (grav.c:8-36) plus O(n) per removal
(kd.c:1385-1447, kdSubPot grav.c:39-60), so the massive groups are isolated and run one process each through the
reference's own restart path (`skid -unbind <name>`, main.c:349-373):

  1. `skid_ref <flags> -nu -maxgroup 20000` on the box -> FoF catalogue (<name>.grp) and density centres (.gtp);
  2. for each of the K largest FoF groups: a .grp holding ONLY that group (label 1, everything else 0 - hence
     everything else is a scoop source, grav.c:63-135) and a one-row .gtp with its centre and vcm;
  3. `skid_ref <flags> -unbind <that>`: kdInGroup + kdInitpGroup + kdReadCenter + kdUnbind + kdTooSmall, all
     K processes in parallel;
  4. stored per group: member indices, centre row, bound flag per member, "Number of particles Unbound", and the
     mass of the surviving group from the reference's .gtp.

Usage:  python tests/golden/make_massive_unbind_golden.py <case> [K]
  case "m20": massive box 2^20 (seed 9), groups of 20 k ... 130 k members, minutes
  case "m20hot": the same box with hotter halos (velocity dispersion x 1.8): a large unbound fraction, so the
              removal loop (arg-max, rcm/vcm update, kdSubPot) runs tens of thousands of times per group
  case "C5":  massive box 2^24 (seed 7, the C5 box): 553 566 / 285 401 / 281 461 / ... members, CPU hours
  case "C5hot": the C5 box with hotter halos: tens of thousands of removals from groups of 10^5 .. 10^6 members
Output tests/golden/massive_unbind_<case>.npz.
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refdump  # noqa: E402
from skid_b200 import synth, tipsy  # noqa: E402

CASES = {"m20": (1 << 20, 9, 0.45), "m20hot": (1 << 20, 9, 0.8), "C5": (1 << 24, 7, 0.45), "C5hot": (1 << 24, 7, 0.8)}


def write_grp(path, labels):
    with open(path, "w") as f:
        f.write("%d\n" % len(labels))
        f.write("\n".join(map(str, labels.tolist())))
        f.write("\n")


def unbind_args(snap):
    fl = snap["flags"]
    return ["-std", "-tau", repr(fl["tau"]), "-m", "8", "-H", "2.8944", "-p", "1"]


def main():
    case = sys.argv[1]
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    n, seed, sigma = CASES[case]
    snap = synth.make_box(n, seed=seed, kind="massive", sigma_frac=sigma)
    work = os.environ.get("SKID_GOLDEN_TMP") or tempfile.mkdtemp(prefix="massive_unbind_")
    os.makedirs(work, exist_ok=True)
    f = os.path.join(work, "in.std")
    synth.write_std(snap, f)
    fof = os.path.join(work, "fof")
    if not os.path.exists(fof + ".grp"):
        t0 = time.time()
        refdump.run_ref(f, snap["ref_args"] + ["-nu", "-maxgroup", "20000"], fof, timeout=6 * 3600)
        print("FoF run %.0f s" % (time.time() - t0), flush=True)
    grp = tipsy.read_array(fof + ".grp").astype(np.int32)
    gtp = tipsy.read_gtp(fof + ".gtp", standard=True)
    sizes = np.bincount(grp)
    sizes[0] = 0
    top = np.argsort(-sizes, kind="stable")[:K]
    top = [int(g) for g in top if sizes[g] >= 20000]
    print("groups", top, "sizes", [int(sizes[g]) for g in top], flush=True)
    procs = []
    for k, g in enumerate(top):
        pre = os.path.join(work, "g%d" % k)
        write_grp(pre + ".grp", (grp == g).astype(np.int32))
        star = np.zeros((1, tipsy.STAR_FIELDS), np.float32)
        star[0, 0] = gtp["mass"][g - 1]
        star[0, 1:4] = gtp["pos"][g - 1]
        star[0, 4:7] = gtp["vel"][g - 1]
        tipsy.write_tipsy(pre + ".gtp", gtp["time"], star=star, standard=True)
        cmd = [refdump.SKID_REF] + unbind_args(snap) + ["-unbind", pre, "-o", pre + "_out"]
        procs.append((k, g, pre, time.time(),
                      subprocess.Popen(cmd, stdin=open(f, "rb"), stdout=open(pre + ".log", "w"), stderr=subprocess.STDOUT)))
    out = dict(n=n, seed=seed, kind="massive", sigma_frac=sigma, K=len(top), unbind_args=" ".join(unbind_args(snap)))
    for k, g, pre, t0, p in procs:
        rc = p.wait()
        assert rc == 0, (k, rc, open(pre + ".log").read()[-500:])
        wall = time.time() - t0
        log = refdump.parse_log(open(pre + ".log").read())
        res = tipsy.read_array(pre + "_out.grp").astype(np.int32)
        members = np.nonzero(grp == g)[0].astype(np.int32)
        bound = res[members] != 0
        assert not res[grp != g].any()
        out["g%d_members" % k] = members
        out["g%d_bound" % k] = np.packbits(bound)
        out["g%d_centre" % k] = np.concatenate([gtp["pos"][g - 1], gtp["vel"][g - 1], [gtp["mass"][g - 1]]]).astype(np.float32)
        out["g%d_log" % k] = np.array([log["nGroupBefore"], log["nUnbound"], log["nGroup"]], np.int64)
        rg = tipsy.read_gtp(pre + "_out.gtp", standard=True)
        out["g%d_gtp_mass" % k] = rg["mass"]
        out["g%d_gtp_vel" % k] = rg["vel"]
        out["g%d_wall_s" % k] = wall
        print("group %d: %d members, unbound %d, bound %d, gtp mass %s, %.0f s" %
              (k, len(members), log["nUnbound"], int(bound.sum()), rg["mass"], wall), flush=True)
    np.savez_compressed(os.path.join(HERE, "massive_unbind_%s.npz" % case), **out)


if __name__ == "__main__":
    main()
