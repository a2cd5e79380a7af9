"""Python model of the chunked scheme of skid_b200/csrc/stats.cu:k_stat_groups (a warp's 32 lanes = numpy
arrays of up to 32 entries): sort by (group, bits(r^2)), sequential float32 chains carried through chunks of 32
members, ballot-style selection of the max / half-mass circular velocity.  TEST TOOL: tests/test_oracle_cpu.py
checks it bit for bit against the oracle's orc_stats, which pins the kernel's LOGIC on the CPU; the CUDA kernel
itself is checked against orc_stats in tests/test_gpu_stats.py."""
import numpy as np

from oracle import orc

f32 = np.float32

def wrap(d, h):
    twoh = f32(2.0) * h
    if d > h: d = f32(d - twoh)
    if d <= -h: d = f32(d + twoh)
    return d

def model(p, rho, nGas, nDark, grp, nGroup, rc, vc, half, G, fExp, fExpHub, fDensMin, fTempMax):
    n = len(p)
    keys = np.zeros(n, np.uint64)
    for i in range(n):
        g = grp[i]
        if g > 0:
            d = [wrap(f32(p["r"][i][k] - rc[g][k]), half[k]) for k in range(3)]
            r2 = f32(f32(f32(f32(0) + f32(d[0]*d[0])) + f32(d[1]*d[1])) + f32(d[2]*d[2]))
            keys[i] = (np.uint64(g) << np.uint64(32)) | np.uint64(r2.view(np.uint32))
    order = np.argsort(keys, kind="stable")
    ks = keys[order]
    cnt = np.bincount(grp, minlength=nGroup)
    start = np.concatenate([[0], np.cumsum(cnt)])
    rows = np.zeros(nGroup, orc.STAT_ROW_DTYPE)
    for g in range(1, nGroup):
        s0, nm = start[g], cnt[g]
        if nm <= 0: continue
        fHalf = f32(0)
        for j in range(nm):
            fHalf = f32(fHalf + f32(f32(0.5) * p["fMass"][order[s0 + j]]))
        fTot = fGas = fStar = fVdisp = fVcirc = fmVcirc = fRVmax = fRhmass = f32(0)
        curD = 0.0
        for base in range(0, nm, 32):
            c = min(32, nm - base)
            idx = order[s0 + base: s0 + base + c]
            r2 = (ks[s0 + base: s0 + base + c] & np.uint64(0xffffffff)).astype(np.uint32).view(np.float32)
            m = p["fMass"][idx]; so = p["fSoft"][idx]
            dd = np.zeros((c, 3), np.float32)
            for l in range(c):
                for k in range(3):
                    dx = wrap(f32(p["r"][idx[l]][k] - rc[g][k]), half[k])
                    e = f32(f32(fExp * f32(p["v"][idx[l]][k] - vc[g][k])) + f32(fExpHub * dx))
                    dd[l, k] = f32(e * e)
            isGas = (idx < nGas) & (rho[idx] >= fDensMin) & (p["fTemp"][idx] <= fTempMax)
            isStar = idx >= nGas + nDark
            outside = r2.astype(np.float64) > 4.0 * so.astype(np.float64) * so.astype(np.float64)
            sq = np.sqrt(r2.astype(np.float64))
            myTot = np.zeros(c, np.float32)
            for j in range(c):
                fTot = f32(fTot + m[j]); myTot[j] = fTot
            for j in range(c):
                if isGas[j]: fGas = f32(fGas + m[j])
                if isStar[j]: fStar = f32(fStar + m[j])
            for j in range(c):
                for k in range(3): fVdisp = f32(fVdisp + dd[j, k])
            gm = (f32(G) * myTot).astype(np.float32)
            rv = sq.astype(np.float32)
            vcl = (gm / rv).astype(np.float32)
            cand = gm.astype(np.float64) / sq
            if np.any(outside & (cand > curD)):
                sc = np.where(outside, cand, -1.0)
                jacc = -1
                for j in range(c):
                    if sc[j] > curD:
                        curD = float(vcl[j]); jacc = j
                if jacc >= 0:
                    fVcirc = vcl[jacc]; fRVmax = rv[jacc]
            if fmVcirc == 0:
                for j in np.nonzero(myTot > fHalf)[0]:
                    if fmVcirc != 0: break
                    fRhmass = rv[j]; fmVcirc = vcl[j]
            r2Last = r2[c - 1]
        sqLast = np.sqrt(np.float64(r2Last))
        flV = f32(np.float64(f32(f32(G) * fTot)) / sqLast)
        if fVcirc == 0:
            fVcirc = flV; fRVmax = f32(sqLast)
        rows[g] = (nm, fTot, fGas, fStar, fVcirc, fmVcirc, flV, fRVmax, fRhmass, r2Last, fVdisp)
    return rows
