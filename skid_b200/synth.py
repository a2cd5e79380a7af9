"""Synthetic clustered particle boxes for BASELINE.json configs 2-5.

All boxes: L = 1 centred on 0 (=> `-p 1`), positions in (-0.5, 0.5], equal masses 1/N (mean
density 1, so `-d 170` means what it means on the demo), no duplicate positions,
eps = tau = 0.0288 * N^(-1/3) (the demo's tau / mean spacing), `-H 2.8944 -G 1`.
35 % of the particles are a uniform background, the rest sit in Hernquist halos (c = 5,
truncated at r_vir of overdensity 200) with sizes drawn from dn/dM ~ M^-1.9, isotropic Gaussian
velocities.  kinds:
  "dark"     configs 2/4: dark only
  "gasdark"  config 3: first N/4 particles gas (T = 1e4, hsmooth = eps), flags -gd -O 0.3 -Lambda 0.7 -z 0.5 -t 30000
  "massive"  config 5: a few very massive halos (largest >= N/16), tau x 4
"""
import numpy as np

from .tipsy import PINIT_DTYPE


def _halo_sizes(rng, n_halo_particles, mmin, mmax, slope=1.9):
    e = 1.0 - slope
    sizes = []
    tot = 0
    while tot < n_halo_particles:
        u = rng.random(4096)
        m = (mmin ** e + u * (mmax ** e - mmin ** e)) ** (1.0 / e)
        for v in m.astype(np.int64):
            v = int(min(v, n_halo_particles - tot))
            if v < mmin:
                # remainder smaller than the smallest halo: dump it into the previous halo
                if sizes:
                    sizes[-1] += n_halo_particles - tot
                else:
                    sizes.append(n_halo_particles - tot)
                tot = n_halo_particles
                break
            sizes.append(v)
            tot += v
            if tot >= n_halo_particles:
                break
    return np.array(sizes, np.int64)


def make_box(n, seed=1234, kind="dark", sigma_frac=0.45):
    rng = np.random.default_rng(seed)
    n = int(n)
    n_bg = int(0.35 * n)
    n_h = n - n_bg
    mmin = 32
    if kind == "massive":
        # a few very massive halos: the largest holds N/8 particles
        big = [n // 8, n // 16, n // 16, n // 32, n // 32]
        rest = n_h - sum(big)
        sizes = np.concatenate([np.array(big, np.int64), _halo_sizes(rng, rest, mmin, max(n // 64, mmin + 1))])
    else:
        sizes = _halo_sizes(rng, n_h, mmin, max(n // 16, mmin + 1))
    nh = len(sizes)
    centres = rng.random((nh, 3)) - 0.5
    mass_h = sizes / n
    rvir = (3.0 * mass_h / (800.0 * np.pi)) ** (1.0 / 3.0)
    a = rvir / 5.0
    hid = np.repeat(np.arange(nh), sizes)
    # Hernquist radii truncated at r_vir: M(<r)/M = r^2/(r+a)^2
    umax = (rvir / (rvir + a)) ** 2
    u = rng.random(n_h) * umax[hid]
    su = np.sqrt(u)
    r = a[hid] * su / (1.0 - su)
    mu = 2.0 * rng.random(n_h) - 1.0
    ph = 2.0 * np.pi * rng.random(n_h)
    st = np.sqrt(1.0 - mu * mu)
    pos_h = centres[hid] + (r[:, None] * np.stack([st * np.cos(ph), st * np.sin(ph), mu], axis=1))
    sig = sigma_frac * np.sqrt(mass_h / rvir)  # fraction of the circular velocity at r_vir (G = 1)
    vel_h = rng.standard_normal((n_h, 3)) * sig[hid][:, None]
    pos = np.concatenate([rng.random((n_bg, 3)) - 0.5, pos_h])
    vel = np.concatenate([rng.standard_normal((n_bg, 3)) * 0.05, vel_h])
    # wrap into (-0.5, 0.5]
    pos = pos - np.floor(pos + 0.5)
    pos[pos <= -0.5] += 1.0
    perm = rng.permutation(n)
    pos = pos[perm].astype(np.float32)
    vel = vel[perm].astype(np.float32)
    pos[pos <= -0.5] = 0.5
    pos[pos > 0.5] = 0.5
    # no duplicate positions (the reference divides by fBall2 / r: SIGFPE, main.c:76).  The check sorts
    # all positions; above 2^22 particles it is skipped (a stray duplicate pair is harmless for both
    # codes - only >= nSmooth coincident particles make fBall2 zero - and the sort costs minutes at 2^27)
    for _ in range(8 if n <= (1 << 22) else 0):
        key = np.ascontiguousarray(pos).view([("", np.float32)] * 3).ravel()
        _, first = np.unique(key, return_index=True)
        if len(first) == n:
            break
        dup = np.ones(n, bool)
        dup[first] = False
        pos[dup] += (rng.random((int(dup.sum()), 3)).astype(np.float32) - 0.5) * 1e-5
        pos = np.clip(pos, np.float32(-0.4999999), np.float32(0.5))
    tau = float(np.float32(0.0288 * n ** (-1.0 / 3.0)))
    eps = tau
    if kind == "massive":
        tau = float(np.float32(4.0 * tau))
    p = np.zeros(n, PINIT_DTYPE)
    p["r"], p["v"] = pos, vel
    p["fMass"] = np.float32(1.0 / n)
    p["fSoft"] = np.float32(eps)
    p["iOrder"] = np.arange(n, dtype=np.int32)
    flags = dict(tau=tau, nSmooth=64, fDensMin=170.0, nMembers=8, H0=2.8944, period=1.0)
    ref_args = ["-std", "-tau", repr(tau), "-s", "64", "-d", "170", "-m", "8", "-H", "2.8944", "-p", "1"]
    nGas = 0
    if kind == "gasdark":
        nGas = n // 4
        p["fTemp"][:nGas] = 1.0e4
        flags.update(bGasAndDark=True, Omega0=0.3, Lambda=0.7, z=0.5, fTempMax=30000.0)
        ref_args += ["-gd", "-O", "0.3", "-Lambda", "0.7", "-z", "0.5", "-t", "30000"]
    return dict(pinit=p, nGas=nGas, nDark=n - nGas, nStar=0, time=1.0, flags=flags, ref_args=ref_args,
                kind=kind, seed=seed, n=n, n_halos=int(nh), largest_halo=int(sizes.max()))


def write_std(snap, path):
    """Write the snapshot as a TIPSY standard file (the reference's stdin)."""
    from .tipsy import pinit_to_records, write_tipsy
    gas, dark, star = pinit_to_records(snap["pinit"], snap["nGas"], snap["nDark"], snap["nStar"])
    write_tipsy(path, snap["time"], gas, dark, star, standard=True)
