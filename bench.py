#!/usr/bin/env python
"""bench.py - SKID group-finding hot path on B200 (BASELINE.json metric: particles grouped/s,
density + move + group + unbind, at 2^24).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the whole hot path (tree + kNN density, move-to-convergence, FoF, micro
steps, centres, unbinding, too-small removal) over one synthetic snapshot.  Workload at N=1:
BASELINE.json configs[2] - synthetic gas+dark box, 2^24 particles, moving both species (-gd),
Lambda cosmology, unbinding on.  N>1 (torchrun, one rank per GPU), weak
scaling: ONE snapshot of N x 2^24 particles (N=8: the 2^27 box of configs[3]) sharded as the
north_star says - particles, tree boxes and scatterers replicated; sorts, kNN queries, movers and
groups shared between the ranks; the exchanges are issued by the library with NCCL (csrc/dist.cu;
DESIGN.md 6).  `--mode replicas` instead runs N independent 2^24 snapshots with no collective at all.

value     = particles / device time of the K timed steps with the snapshot already in HBM and the results
            left in HBM (CUDA events on the context's stream, max over ranks).
e2e       = same metric through the C-ABI with HOST buffers: pinned AoS snapshot -> device inside the
            timed region, labels + catalogue read back every step (N>1: every rank uploads 1/N of the
            snapshot, rank 0 reads the results back - what host/skid -gpus N does).
roofline  = the dominant kernel, k_tile_step (gradient walk + move): algorithmic bytes (SURVEY 8d:
            24*C+24 B per mover-step, C = 85) / its device time, measured in this run with CUDA events around
            every launch (skidgpu_set_profile); roofline_builds / roofline_knn: the tile-list builds and the
            kNN kernel the same way.
parity    = (N>1) same-group fraction of a 2^20 box run sharded over the N ranks against rank 0 running it
            alone, measured during warm-up; the run FAILS below 0.999.
cpu_baseline / --impl reference = the UNMODIFIED reference (oracle/_ref/skid_ref, serial: 1 core) on
            a bounded sample of the same generator, timed on this box's host.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particles grouped/sec (density+move+group+unbind)"
UNIT = "particles/s"
BYTES_PER_MOVER_STEP = 24 * 85 + 24  # SURVEY.md 8d / DESIGN.md: 24 B per containing scatterer (C = 85) + 24 B mover r/w
KF = {"tile_step": 0, "knn": 1, "builds": 2, "fallback": 3, "prune": 4}   # skidgpu_kernel_ms families


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """DRAM bytes per unit of the profiled kernels from this round's ncu --set full captures (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4) if r[5 + k].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference_sample(kind, log2n, seed, noprune=False):
    """Time the unmodified reference (serial) on a bounded sample.  Returns (particles/s, info)."""
    from oracle import refdump
    from skid_b200 import synth
    if not refdump.have_ref():
        raise RuntimeError("oracle/_ref/skid_ref not built (run `make ref` in the build container)")
    snap = synth.make_box(1 << log2n, seed=seed, kind=kind)
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "in.std")
        synth.write_std(snap, f)
        out, wall = refdump.run_ref(f, snap["ref_args"], os.path.join(td, "ref"), noprune=noprune)
    log = refdump.parse_log(out)
    stages = sum(log["times"].values())  # the reference's own "SKID CPU Time" lines (user CPU, I/O excluded)
    return (1 << log2n) / stages, dict(wall_s=wall, stage_s=stages, times=log["times"], groups=log.get("nGroup"))


def run_port_sample(kind, log2n, seed):
    """Fallback CPU arm when the compiled reference did not travel: the oracle's C restatement of the same stage
    script (oracle/pipeline.py; brute-force kNN, so the sample is smaller).  Returns (particles/s, info)."""
    from oracle import pipeline
    from skid_b200 import synth
    from skid_b200.api import csmExp2Hub
    snap = synth.make_box(1 << log2n, seed=seed, kind=kind)
    t0 = time.time()
    res = pipeline.run_port(snap, csmExp2Hub)
    wall = time.time() - t0
    stages = sum(res["times"].values())
    return (1 << log2n) / stages, dict(wall_s=wall, stage_s=stages, times=res["times"], groups=res["nGroup"])


def cpu_arm(kind, ref_log2n, seed):
    """(particles/s, info, kind, log2n): the unmodified reference if oracle/_ref is here, else the oracle port."""
    from oracle import refdump
    if refdump.have_ref():
        v, info = run_reference_sample(kind, ref_log2n, seed)
        return v, info, "reference", ref_log2n
    log2n = min(ref_log2n, 14)
    v, info = run_port_sample(kind, log2n, seed)
    return v, info, "port", log2n


def reference_arm(a, workload):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  The
    reference is strictly serial, so one core; its cost per particle is flat in N (measured 2^15 .. 2^24), so a
    bounded sample of the same generator stands for the workload: ONE run of a 2^20 box (the largest that fits the
    few-minute window; the full 2^24 box takes it ~40 minutes), reused for every warm-up and timed step."""
    v, info, cpu_kind, cpu_log2n = cpu_arm(a.kind, a.ref_log2n, seed=7)
    sec = (1 << cpu_log2n) / v
    sample = (f"ONE run of a 2^{cpu_log2n}-particle box of the same generator/flags, reused for all {a.warmup}+{a.steps} "
              f"steps (the full 2^{a.log2n} box needs ~{sec * (1 << (a.log2n - cpu_log2n)) / 60:.0f} CPU-minutes at this "
              f"rate); time = sum of the "
              + ("reference's own stage timers" if cpu_kind == "reference" else
                 "stage times of the oracle's C restatement (oracle/pipeline.py; oracle/_ref did not travel)"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": cpu_kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_detail": info,
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=24, help="particles = 2^log2n (BASELINE metric is quoted at 24)")
    ap.add_argument("--kind", default="gasdark", choices=["dark", "gasdark", "massive"])
    ap.add_argument("--cpu-log2n", type=int, default=17, help="size of the bounded CPU-baseline sample of the default run")
    ap.add_argument("--ref-log2n", type=int, default=20, help="size of the one sample of --impl reference")
    ap.add_argument("--parity-log2n", type=int, default=20, help="N>1: size of the sharded-vs-single parity box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="shard", choices=["shard", "replicas"], help="multi-GPU layout (N>1)")
    a = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "NONE"  # keep stdout to the one JSON line (NCCL prints its version banner there)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"synthetic {'gas+dark' if a.kind == 'gasdark' else a.kind} box 2^{a.log2n} particles, periodic L=1, "
                f"moving both species, Lambda cosmology, unbinding on (BASELINE configs[2])"
                if a.kind == "gasdark" else f"synthetic {a.kind} box 2^{a.log2n} particles, periodic L=1")

    if a.impl == "reference":
        return reference_arm(a, workload) if rank == 0 else 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    from skid_b200 import api, parallel, synth
    from skid_b200.tipsy import PINIT_DTYPE
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the B200 arm has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shard = world > 1 and a.mode == "shard"
    n = (1 << a.log2n) * (world if shard else 1)
    dev0 = torch.device("cuda", local)
    if shard:
        # one snapshot for everybody: rank 0 generates, NCCL broadcasts the SoA columns
        meta = torch.zeros(4, dtype=torch.int64, device=dev0)
        if rank == 0:
            snap = synth.make_box(n, seed=7, kind=a.kind)
            meta[:] = torch.tensor([snap["nGas"], snap["nDark"], snap["nStar"], n])
        dist.broadcast(meta, 0)
        nGas, nDark, nStar = int(meta[0]), int(meta[1]), int(meta[2])
        dev = []
        for k in range(9):
            if rank == 0:
                p = snap["pinit"]
                col = (p["r"][:, k] if k < 3 else p["v"][:, k - 3] if k < 6 else p[("fMass", "fSoft", "fTemp")[k - 6]])
                t = torch.from_numpy(np.ascontiguousarray(col)).to(dev0)
            else:
                t = torch.empty(n, dtype=torch.float32, device=dev0)
            dist.broadcast(t, 0)
            dev.append(t)
        fl = synth.make_box(1024, seed=7, kind=a.kind)["flags"]
        fl["tau"] = float(np.float32((4.0 if a.kind == "massive" else 1.0) * np.float32(0.0288 * n ** (-1.0 / 3.0))))
        p = np.zeros(n, PINIT_DTYPE)
        for k in range(9):
            h = dev[k].cpu().numpy()
            if k < 3:
                p["r"][:, k] = h
            elif k < 6:
                p["v"][:, k - 3] = h
            else:
                p[("fMass", "fSoft", "fTemp")[k - 6]] = h
        p["iOrder"] = np.arange(n, dtype=np.int32)
        snap = dict(pinit=p, nGas=nGas, nDark=nDark, nStar=nStar, flags=fl)
    else:
        snap = synth.make_box(n, seed=7 + rank, kind=a.kind)
        fl = snap["flags"]
        p = snap["pinit"]
        cols = [p["r"][:, 0], p["r"][:, 1], p["r"][:, 2], p["v"][:, 0], p["v"][:, 1], p["v"][:, 2], p["fMass"],
                p["fSoft"], p["fTemp"]]
        dev = [torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in cols]
    # pinned host AoS (e2e leg)
    try:
        pin = torch.empty(n * PINIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    except Exception:
        pin = torch.empty(n * PINIT_DTYPE.itemsize, dtype=torch.uint8)
    host_aos = pin.numpy().view(PINIT_DTYPE)
    host_aos[:] = p
    del p
    # pinned result buffers of the e2e leg (labels by iOrder, catalogue rows)
    from skid_b200.tipsy import PGROUP_DTYPE
    try:
        pin_grp = torch.empty(n, dtype=torch.int32).pin_memory()
        pin_cat = torch.empty((n // 16 + 1024) * PGROUP_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    except Exception:
        pin_grp = torch.empty(n, dtype=torch.int32)
        pin_cat = torch.empty((n // 16 + 1024) * PGROUP_DTYPE.itemsize, dtype=torch.uint8)
    out_grp, out_cat = pin_grp.numpy(), pin_cat.numpy().view(PGROUP_DTYPE)
    torch.cuda.synchronize()

    per = (fl["period"],) * 3
    sk = api.SkidGPU(per, (0.0, 0.0, 0.0), bPeriodic=True, device=local)
    sk.set_profile(True)
    reducer = None
    if shard:
        parallel.init_comm(sk, dist, rank, world)           # the library's own NCCL communicator
    tau = float(np.float32(fl["tau"]))
    fCvg = float(np.float32(0.5 * tau))
    fScoop = float(np.float32(2.0 * tau))
    fStep = float(np.float32(0.5 * fCvg))
    f32 = lambda v: float(np.float32(v))
    z = f32(fl.get("z", 0.0))
    a32 = f32(1.0 / (1.0 + z))
    fCosmo = a32 * api.csmExp2Hub(a32, f32(fl["H0"]), f32(fl.get("Omega0", 1.0)), f32(fl.get("Lambda", 0.0)))

    def one_pass(host):
        fetch = host and rank == 0 if shard else host
        if shard:
            return parallel.run_skid_sharded(sk, reducer, host_aos, snap["nGas"], snap["nDark"], snap["nStar"], fl,
                                             rank, world, host=host, dev_ptrs=[t.data_ptr() for t in dev], fetch=fetch,
                                             out_grp=out_grp, out_cat=out_cat)
        sk.log = []
        if host:
            sk.set_particles(host_aos, snap["nGas"], snap["nDark"], snap["nStar"])
        else:
            sk.set_particles_dev([t.data_ptr() for t in dev], n, snap["nGas"], snap["nDark"], snap["nStar"])
        sk.smDensityInit(fl["nSmooth"], fl.get("bGasAndDark", False), False, want_arrays=False)
        sk.move(fl["fDensMin"], fl.get("fTempMax", api.FLT_MAX), api.FLT_MAX, fCvg, fStep)
        sk.kdFoF(tau)
        sk.microstep(5, f32(0.1 * fStep))
        sk.kdCalcCenter(fetch=False)
        return sk.kdUnbind(1.0, z, fCosmo, api.SPLINE, fScoop, False, api.INT_MAX, fl["nMembers"], fetch=fetch,
                           out_grp=out_grp, out_cat=out_cat)

    stream = torch.cuda.ExternalStream(sk.stream(), device=torch.device("cuda", local))

    def timed(host, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = dict(fam_ms={k: 0.0 for k in KF}, fam_n={k: 0 for k in KF}, stage_ms={}, groups=0, d2h=0)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = sk.counter(0)
        ms0 = sk.counter(1)
        cb0 = sk.comm_bytes()
        ev0.record(stream)
        for _ in range(steps):
            grp, cat, nUnb, nBefore = one_pass(host)
            for k, w in KF.items():
                ms_, n_ = sk.kernel_ms(w)
                stats["fam_ms"][k] += ms_
                stats["fam_n"][k] += n_
            for k, v in sk.stage_ms().items():
                stats["stage_ms"][k] = stats["stage_ms"].get(k, 0.0) + v / steps
            stats["groups"] = sk.nGroup - 1
            stats["groups_before"] = nBefore
            stats["unbound"] = nUnb
            if grp is not None:
                stats["d2h"] = grp.nbytes + cat.nbytes
        ev1.record(stream)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        stats["launches"] = sk.counter(0) - l0
        stats["mover_steps"] = sk.counter(1) - ms0
        stats["nMove"] = sk.nMove
        cb1 = sk.comm_bytes()
        stats["comm_bytes"] = (cb1[0] - cb0[0]) / steps
        stats["comm_calls"] = (cb1[1] - cb0[1]) / steps
        if dist is not None:
            t = torch.tensor([ms, float(stats["mover_steps"]), stats["fam_ms"]["tile_step"], stats["fam_ms"]["knn"],
                              stats["fam_ms"]["builds"], float(stats["d2h"])], device="cuda", dtype=torch.float64)
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)      # slowest rank's times
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms = float(mx[0].item())
            stats["mover_steps"] = float(t[1].item())      # mover-steps of all shards
            stats["fam_ms"]["tile_step"], stats["fam_ms"]["knn"], stats["fam_ms"]["builds"] = (float(mx[k].item()) for k in (2, 3, 4))
            stats["d2h"] = float(mx[5].item())
        return ms, stats

    W = max(a.warmup, 3)
    parity = None
    if shard:
        # sharded vs single-GPU on the same small box, once, before the warm-up passes: the driver's evidence that
        # the N-rank run computes the single-GPU groups
        canonical_labels = parallel.canonical_labels  # (the oracle is only touched by the CPU arms)
        ps = synth.make_box(1 << a.parity_log2n, seed=11, kind=a.kind)
        g_sh, cat_sh, unb_sh, before_sh = parallel.run_skid_sharded(sk, None, ps["pinit"], ps["nGas"], ps["nDark"], ps["nStar"],
                                                                    ps["flags"], rank, world, host=True, fetch=True)
        it_sh = len([l for l in sk.log if l[0] == 0])
        flag = torch.ones(1, device=dev0)
        if rank == 0:
            single = api.run_skid(ps["pinit"], ps["nGas"], ps["nDark"], ps["nStar"], device=local, want_arrays=False, **ps["flags"])
            same = float(np.mean(canonical_labels(single["grp"]) == canonical_labels(g_sh)))
            parity = {"box": f"{a.kind} 2^{a.parity_log2n}, seed 11", "same_group": same,
                      "groups": [len(cat_sh) - 1, single["nGroup"]], "groups_before_unbind": [before_sh, single["nGroupBefore"]],
                      "unbound": [unb_sh, single["nUnbound"]], "ittr": [it_sh, single["nIttr"]]}
            if same < 0.999 or before_sh != single["nGroupBefore"]:
                flag[0] = 0
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) == 0:
            if rank == 0:
                print(json.dumps({"error": "sharded run does not reproduce the single-GPU groups", "parity": parity}), flush=True)
            sk.close()
            dist.destroy_process_group()
            return 3
    timed(False, W)                       # warm-up, device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, st = timed(False, a.steps)
    clocks = sampler.stop()
    timed(True, 1)                        # warm the host path (pinned staging, pool growth)
    ms_e2e, st_e = timed(True, a.steps)

    total_particles = n if shard else n * world
    value = total_particles * a.steps / (ms_dev * 1e-3)
    e2e_value = total_particles * a.steps / (ms_e2e * 1e-3)
    peak, peak_src = load_peaks()
    traffic = load_traffic()
    # per GPU: the mover-steps of all shards / N over the slowest shard's kernel time, against ONE GPU's peak
    nshare = world if shard else 1
    per_gpu_steps = st["mover_steps"] / nshare
    ts_ms, ts_n = st["fam_ms"]["tile_step"], max(st["fam_n"]["tile_step"], 1)
    achieved = per_gpu_steps * BYTES_PER_MOVER_STEP / (ts_ms * 1e-3) / 1e9 if ts_ms > 0 else 0.0
    knn_ms = st["fam_ms"]["knn"] / a.steps
    qbytes = 16 * fl["nSmooth"] + 24
    knn_ach = (n / nshare) * qbytes / (knn_ms * 1e-3) / 1e9 if knn_ms > 0 else None
    bld_ms = st["fam_ms"]["builds"]

    def per_launch(key, units):
        v = traffic.get(key)
        return None if v is None else float(v) * units

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "particles_per_gpu": 1 << a.log2n, "nSmooth": fl["nSmooth"], "tau": tau,
                   "parallelism": "1 GPU" if world == 1 else (
                       f"one {n}-particle snapshot sharded over {world} GPUs: replicated particles/boxes/scatterers; "
                       f"sorts, kNN queries, movers and groups shared; {st['comm_calls']:.0f} NCCL exchanges per step issued "
                       f"by the library" if shard else f"{world} independent snapshots, one per GPU"),
                   "l2": "inputs (604 MB SoA at 2^24) larger than the 126 MB L2; no explicit flush",
                   "movers": st["nMove"], "groups_before_unbind": st["groups_before"], "groups": st["groups"],
                   "unbound": st["unbound"]},
        "stage_ms": st["stage_ms"],
        "knn_queries_per_s": (n / nshare) * nshare / (knn_ms * 1e-3) if knn_ms > 0 else None,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_aos.nbytes),  # whole job: every rank uploads 1/N of the snapshot
                "d2h_bytes_per_step": int(st_e["d2h"]), "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(st["launches"]),
        "clocks": clocks,
        "roofline": {"kernel": "k_tile_step (gradient walk + move; one launch per step, timed alone)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": per_launch("tile_step_dram_bytes_per_mover_step", per_gpu_steps / ts_n),
                     "traffic_source": traffic.get("source"),
                     "peak_source": peak_src, "bytes_per_mover_step": BYTES_PER_MOVER_STEP,
                     "mover_steps_per_step": st["mover_steps"] / a.steps,
                     "launches_per_step": st["fam_n"]["tile_step"] / a.steps,
                     "avg_launch_ms": ts_ms / ts_n, "share_of_step": ts_ms / ms_dev},
        # the tile-list builds of the same stage (re-sort, k_super_walk, k_tile_filter, k_tile_walk), one span per rebuild:
        # algorithmic bytes = one more pass over the same scatterer records per mover and rebuild
        "roofline_builds": {"kernel": "k_super_walk + k_tile_filter + k_tile_walk (+ re-sort every 4th rebuild)", "bound": "hbm",
                            "unit": "GB/s", "peak": peak, "ms_per_step": bld_ms / a.steps,
                            "rebuilds_per_step": st["fam_n"]["builds"] / a.steps, "share_of_step": bld_ms / ms_dev,
                            "achieved": (per_gpu_steps / 5.0) * BYTES_PER_MOVER_STEP / (bld_ms * 1e-3) / 1e9 if bld_ms > 0 else None,
                            "note": "units = movers present at a rebuild (mover-steps / 5)"},
        "fallback_ms_per_step": st["fam_ms"]["fallback"] / a.steps, "prune_ms_per_step": st["fam_ms"]["prune"] / a.steps,
        # second kernel family, for the north_star's "kNN walk vs HBM roofline" line (16k+24 B per query)
        "roofline_knn": {"kernel": "k_knn_density", "bound": "hbm", "unit": "GB/s", "peak": peak, "bytes_per_query": qbytes,
                         "achieved": knn_ach, "frac": knn_ach / peak if knn_ach else None, "ms": knn_ms,
                         "traffic": per_launch("knn_dram_bytes_per_query", n / nshare),
                         "l2": traffic.get("knn_l2"),
                         "note": "on-chip bound by design: neighbouring queries share their neighbours through L1/L2, so DRAM "
                                 "traffic is a few per cent of the algorithmic bytes; `l2` holds the ncu L2 figures of this "
                                 "round's capture (profiles/)"},
    }
    if out["roofline_builds"]["achieved"]:
        out["roofline_builds"]["frac"] = out["roofline_builds"]["achieved"] / peak
    if shard:
        out["config"]["particles_total"] = n
        out["config"]["nccl_bytes_per_step"] = st["comm_bytes"]
        out["parity"] = parity
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            v, info, cpu_kind, cpu_log2n = cpu_arm(a.kind, a.cpu_log2n, seed=7)
            what = ("unmodified reference (oracle/_ref/skid_ref, serial)" if cpu_kind == "reference" else
                    "oracle C restatement (oracle/pipeline.py, serial; oracle/_ref did not travel)")
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                                   "sample": f"{what} on a 2^{cpu_log2n}-particle "
                                             f"box of the same generator/flags; sum of its stage timers "
                                             f"{info['stage_s']:.1f} s (wall {info['wall_s']:.1f} s); host has {os.cpu_count()} cores"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    del dev
    torch.cuda.empty_cache()  # nothing of torch's may outlive the context's stream
    sk.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
