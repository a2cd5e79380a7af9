#!/usr/bin/env python
"""bench.py - SKID group-finding hot path on B200 (BASELINE.json metric: particles grouped/s,
density + move + group + unbind, at 2^24).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the whole hot path (tree + kNN density, move-to-convergence, FoF, micro
steps, centres, unbinding, too-small removal) over one synthetic snapshot.  Workload at N=1:
BASELINE.json configs[2] - synthetic gas+dark box, 2^24 particles, moving both species (-gd),
Lambda cosmology, unbinding on.  N>1 (torchrun, one rank per GPU), weak
scaling: ONE snapshot of N x 2^24 particles (N=8: the 2^27 box of configs[3]) sharded as the
north_star says - particles, trees and scatterers replicated, kNN queries / movers / groups
sharded, NCCL only for the small agreement points (skid_b200/parallel.py).
`--mode replicas` instead runs N independent 2^24 snapshots with no collective at all.

value     = particles / device time of the K timed steps with the snapshot already in HBM
            (CUDA events on the context's stream, max over ranks).
e2e       = same metric through the C-ABI with HOST buffers: pinned AoS snapshot -> device inside the
            timed region, labels + catalogue read back every step.
roofline  = the dominant kernel (gradient walk + move, k_tile_step, timed together with the tile-list
            builds and the fallback walk that belong to a step): algorithmic bytes (24*C+24 B
            per mover-step, C = 85) / device time measured inside this run; one "launch" = one step.
cpu_baseline / --impl reference = the UNMODIFIED reference (oracle/_ref/skid_ref, serial: 1 core) on
            a bounded sample of the same generator (2^17 particles), timed on this box's host.
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
# the full 2^24 workload through the unmodified reference, measured once in the build container (one core)
FULL_SIZE_NOTE = ("; the FULL 2^24 box took the reference 2448 s of stage time = 6.85e3 particles/s with identical "
                  "counters (69 Ittr lines, 138065 groups before unbinding, 55419 groups; profiles/r01_reference_full_size.json)")
BYTES_PER_MOVER_STEP = 24 * 85 + 24  # SURVEY.md 8d / DESIGN.md: 24 B per containing scatterer (C = 85) + 24 B mover r/w


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=24, help="particles = 2^log2n (BASELINE metric is quoted at 24)")
    ap.add_argument("--kind", default="gasdark", choices=["dark", "gasdark", "massive"])
    ap.add_argument("--cpu-log2n", type=int, default=17, help="size of the bounded CPU-baseline sample")
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

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        W = max(a.warmup, 0)
        vals = []
        info = {}
        cpu_kind, cpu_log2n = "reference", a.cpu_log2n
        for it in range(W + a.steps):
            v, info, cpu_kind, cpu_log2n = cpu_arm(a.kind, a.cpu_log2n, seed=7)
            if it >= W:
                vals.append((1 << cpu_log2n) / v)
        sec = float(np.mean(vals))
        value = (1 << cpu_log2n) / sec
        sample = (f"2^{cpu_log2n}-particle box of the same generator/flags (full 2^{a.log2n} needs ~{140e-6 * (1 << a.log2n) / 60:.0f} "
                  f"CPU-minutes); time = sum of the "
                  + ("reference's own stage timers" if cpu_kind == "reference" else
                     "stage times of the oracle's C restatement (oracle/pipeline.py; oracle/_ref did not travel)"))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": cpu_kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_detail": info,
        }))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    from skid_b200 import api, synth
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
        from skid_b200 import parallel
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
        fl["tau"] = float(np.float32(0.0288 * n ** (-1.0 / 3.0)))
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
    torch.cuda.synchronize()

    per = (fl["period"],) * 3
    sk = api.SkidGPU(per, (0.0, 0.0, 0.0), bPeriodic=True, device=local)
    reducer = None
    if shard:
        sk.set_shard(rank, world)
        reducer = parallel.Reducer(dist, dev0, sk.stream())
        sk.set_reduce_cb(reducer.cb)
    tau = float(np.float32(fl["tau"]))
    fCvg = float(np.float32(0.5 * tau))
    fScoop = float(np.float32(2.0 * tau))
    fStep = float(np.float32(0.5 * fCvg))
    f32 = lambda v: float(np.float32(v))
    z = f32(fl.get("z", 0.0))
    a32 = f32(1.0 / (1.0 + z))
    fCosmo = a32 * api.csmExp2Hub(a32, f32(fl["H0"]), f32(fl.get("Omega0", 1.0)), f32(fl.get("Lambda", 0.0)))

    def one_pass(host):
        if shard:
            return parallel.run_skid_sharded(sk, reducer, host_aos, snap["nGas"], snap["nDark"], snap["nStar"], fl,
                                             rank, world, host=host, dev_ptrs=[t.data_ptr() for t in dev])
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
        grp, cat, nUnb, nBefore = sk.kdUnbind(1.0, z, fCosmo, api.SPLINE, fScoop, False, api.INT_MAX, fl["nMembers"])
        return grp, cat, nUnb, nBefore

    stream = torch.cuda.ExternalStream(sk.stream(), device=torch.device("cuda", local))

    def timed(host, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = dict(mover_steps=0, move_kernel_ms=0.0, move_launches=0, knn_ms=0.0, stage_ms={}, groups=0)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = sk.counter(0)
        ms0 = sk.counter(1)
        ev0.record(stream)
        for _ in range(steps):
            grp, cat, nUnb, nBefore = one_pass(host)
            kms, kl = sk.kernel_ms(0)
            stats["move_kernel_ms"] += kms
            stats["move_launches"] += kl
            stats["knn_ms"] += sk.kernel_ms(1)[0]
            for k, v in sk.stage_ms().items():
                stats["stage_ms"][k] = stats["stage_ms"].get(k, 0.0) + v / steps
            stats["groups"] = len(cat) - 1
            stats["groups_before"] = nBefore
            stats["unbound"] = nUnb
        ev1.record(stream)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        stats["launches"] = sk.counter(0) - l0
        stats["mover_steps"] = sk.counter(1) - ms0
        stats["nMove"] = sk.nMove
        stats["d2h"] = grp.nbytes + cat.nbytes
        if dist is not None:
            t = torch.tensor([ms, float(stats["mover_steps"]), stats["move_kernel_ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            tm = t[2:3].clone()
            dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)   # mover-steps of all shards
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)        # slowest shard's kernel time
            ms = float(t[0].item())
            stats["mover_steps"] = float(t[1].item())
            stats["move_kernel_ms"] = float(tm.item())
        return ms, stats

    W = max(a.warmup, 3)
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
    # per GPU: the mover-steps of all shards / N over the slowest shard's kernel time, against ONE GPU's peak
    per_gpu_steps = st["mover_steps"] / (world if shard else 1)
    achieved = per_gpu_steps * BYTES_PER_MOVER_STEP / (st["move_kernel_ms"] * 1e-3) / 1e9 if st["move_kernel_ms"] > 0 else 0.0
    traffic = None
    try:  # ncu dram__bytes_read+write of the step's kernels per mover-step (profiles/roofline_traffic.json)
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            per = float(json.load(f)["move_dram_bytes_per_mover_step"])
        traffic = per * per_gpu_steps / max(st["move_launches"], 1)
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "particles_per_gpu": 1 << a.log2n, "nSmooth": fl["nSmooth"], "tau": tau,
                   "parallelism": "1 GPU" if world == 1 else (
                       f"one {n}-particle snapshot sharded over {world} GPUs: replicated scatterers/trees, sharded "
                       f"kNN queries, movers and groups; NCCL all-reduce at {reducer.calls // max(1, (W + a.steps + 1 + a.steps))} agreement "
                       f"points per step" if shard else f"{world} independent snapshots, one per GPU"),
                   "l2": "inputs (604 MB SoA at 2^24) larger than the 126 MB L2; no explicit flush",
                   "movers": st["nMove"], "groups_before_unbind": st["groups_before"], "groups": st["groups"],
                   "unbound": st["unbound"]},
        "stage_ms": st["stage_ms"],
        "knn_queries_per_s": n / (st["knn_ms"] / a.steps * 1e-3) if st["knn_ms"] > 0 else None,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_aos.nbytes),
                "d2h_bytes_per_step": int(st_e["d2h"]), "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(st["launches"]),
        "clocks": clocks,
        "roofline": {"kernel": "k_tile_step (+ k_super_walk/k_tile_filter list builds and the k_move_step fallback: everything a step launches)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_mover_step": BYTES_PER_MOVER_STEP,
                     "mover_steps_per_step": st["mover_steps"] / a.steps,
                     "launches_per_step": st["move_launches"] / a.steps,
                     "avg_launch_ms": st["move_kernel_ms"] / max(st["move_launches"], 1),
                     "share_of_step": st["move_kernel_ms"] / ms_dev},
        # second kernel family, for the north_star's "kNN walk vs HBM roofline" line (16k+24 B per query)
        "roofline_knn": {"kernel": "k_knn_density", "bound": "hbm", "unit": "GB/s", "peak": peak,
                         "bytes_per_query": 16 * fl["nSmooth"] + 24,
                         "achieved": (n / (world if shard else 1)) * (16 * fl["nSmooth"] + 24) / (st["knn_ms"] / a.steps * 1e-3) / 1e9
                         if st["knn_ms"] > 0 else None,
                         "note": "on-chip bound by design: neighbouring queries share their neighbours through L1/L2 "
                                 "(ncu: 33 B of DRAM traffic per query, issue-active 77 %)"},
    }
    if out["roofline_knn"]["achieved"]:
        out["roofline_knn"]["frac"] = out["roofline_knn"]["achieved"] / peak
    try:
        # What actually bounds both kernel families (ncu: DRAM < 2 % of peak, sm__throughput 68-80 %): warp
        # instruction issue.  Instructions per unit are ncu counts (smsp__inst_executed.sum of the profiles named
        # below / units of that launch); peak = 148 SMs x 4 schedulers x 1 warp instruction per clock.
        issue_peak = 148 * 4 * float(clocks.get("sm_mhz") or 1965.0) * 1e6
        per_gpu_q = n / (world if shard else 1)
        knn_rate = per_gpu_q * 6800.0 / (st["knn_ms"] / a.steps * 1e-3) if st["knn_ms"] > 0 else None
        mv_rate = per_gpu_steps * 477.0 / (st["move_kernel_ms"] * 1e-3) if st["move_kernel_ms"] > 0 else None
        out["issue_roofline"] = {
            "unit": "warp instructions/s", "peak": issue_peak,
            "knn": {"inst_per_query": 6800, "achieved": knn_rate, "frac": knn_rate / issue_peak if knn_rate else None,
                    "source": "profiles/r01_v6_knn_2e22_lines.txt"},
            "move": {"inst_per_mover_step": 477, "achieved": mv_rate, "frac": mv_rate / issue_peak if mv_rate else None,
                     "source": "profiles/r01_v6_move_launches_2e24.csv (all kernels of a step)"}}
    except Exception:
        pass
    if shard:
        out["config"]["particles_total"] = n
        out["config"]["nccl_bytes_per_step"] = reducer.bytes // max(1, (W + a.steps + 1 + a.steps))
        out["config"]["reduce_callback_host_ms_per_step"] = 1e3 * reducer.host_s / max(1, (W + a.steps + 1 + a.steps))
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            v, info, cpu_kind, cpu_log2n = cpu_arm(a.kind, a.cpu_log2n, seed=7)
            what = ("unmodified reference (oracle/_ref/skid_ref, serial)" if cpu_kind == "reference" else
                    "oracle C restatement (oracle/pipeline.py, serial; oracle/_ref did not travel)")
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                                   "sample": f"{what} on a 2^{cpu_log2n}-particle "
                                             f"box of the same generator/flags; sum of its stage timers "
                                             f"{info['stage_s']:.1f} s (wall {info['wall_s']:.1f} s); host has {os.cpu_count()} cores"
                                             + (FULL_SIZE_NOTE if a.kind == "gasdark" and a.log2n == 24 else "")}
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
