"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the few exchange
points of the sharded pipeline.  The library calls back into `Reducer` at its agreement
points (include/skidgpu.h: skidgpu_reduce_cb); the all-gather of converged mover positions before
FoF / centres is driven from here.

What is sharded: kNN queries (contiguous Morton ranges; fBall2 and f64 density partials are summed),
movers (block-cyclic over the Morton-ordered mover list: contiguous ranges leave whole halos on one
rank and the others wait for it at every step) and groups for unbinding (g % nranks).
What is replicated: particles, trees, scatterers, FoF, catalogue bookkeeping.

The same code runs on CPU with the gloo backend on host buffers (tests/test_parallel_cpu.py).
"""
import ctypes as C

import numpy as np
import torch

REDUCE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int)

_NP = {0: np.int32, 1: np.uint8, 2: np.float32, 3: np.float64}
_TYPESTR = {0: "<i4", 1: "|u1", 2: "<f4", 3: "<f8"}


def shard_range(n, rank, nranks):
    """Contiguous range [lo, hi) of n items owned by `rank` (same rule as the library: kd shards are
    floor(n*rank/nranks) .. floor(n*(rank+1)/nranks))."""
    return (n * rank) // nranks, (n * (rank + 1)) // nranks


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can alias a raw device pointer."""

    def __init__(self, ptr, count, dtype):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": _TYPESTR[dtype],
                                         "data": (int(ptr), False), "version": 2}


def tensor_from_pointer(ptr, count, dtype, device):
    """Alias `count` elements at `ptr` (device pointer when device is CUDA, host pointer otherwise)."""
    if device.type == "cuda":
        return torch.as_tensor(_DevArray(ptr, count, dtype), device=device)
    buf = (C.c_char * (int(count) * np.dtype(_NP[dtype]).itemsize)).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=_NP[dtype]))


class Reducer:
    """skidgpu_reduce_cb implementation over torch.distributed.  `stream` = the context's CUDA stream
    (int handle) so that collectives are ordered with the library's kernels; None on CPU."""

    def __init__(self, dist, device, stream=None):
        self.dist, self.device = dist, device
        self.stream = torch.cuda.ExternalStream(stream, device=device) if (stream and device.type == "cuda") else None
        self.calls = 0
        self.bytes = 0
        self.host_s = 0.0      # host time spent inside the callback (per-step agreement points are latency bound)
        self._alias = {}       # (ptr, count, dtype) -> tensor: the library reuses a handful of buffers
        self.cb = REDUCE_CB(self._call)

    def reduce_tensor(self, t, op):
        d = self.dist
        rop = {0: d.ReduceOp.MIN, 1: d.ReduceOp.MAX, 2: d.ReduceOp.SUM}[op]
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                d.all_reduce(t, op=rop)
        else:
            d.all_reduce(t, op=rop)
        self.calls += 1
        self.bytes += t.numel() * t.element_size()

    def _call(self, user, ptr, count, dtype, op):
        import time
        t0 = time.perf_counter()
        try:
            key = (ptr, count, dtype)
            t = self._alias.get(key)
            if t is None:
                t = tensor_from_pointer(ptr, count, dtype, self.device)
                if count <= 16:  # only the small per-step buffers are worth caching (large ones get reallocated)
                    self._alias[key] = t
            if dtype == 1 and op != 2:
                # NCCL has no uint8 min/max guarantee across versions: widen flags through int32
                w = t.to(torch.int32)
                self.reduce_tensor(w, op)
                t.copy_(w.to(torch.uint8))
            else:
                self.reduce_tensor(t, op)
            self.host_s += time.perf_counter() - t0
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("skid_b200.parallel.Reducer:", e, flush=True)
            return 1

    def allgather_owned(self, arrays, lo, hi):
        """Every rank owns [lo,hi) of each 1-D tensor in `arrays`; afterwards all ranks hold all
        owned ranges.  Implemented as zero-the-rest + sum (ranges are disjoint), which needs no
        equal-size assumption."""
        for t in arrays:
            t[:lo] = 0
            t[hi:] = 0
            self.reduce_tensor(t, 2)


def merge_labels(dist, labels):
    """Labels after sharded unbinding: the owner of a group wrote 0 for unbound members, every other
    rank still holds the FoF label -> element-wise minimum (what the library asks the Reducer to do)."""
    dist.all_reduce(labels, op=dist.ReduceOp.MIN)
    return labels


def upload_sliced(reducer, pinit, rank, nranks):
    """Host AoS (PINIT records, 12 x 4 bytes) -> 9 replicated device SoA columns (x y z vx vy vz mass soft temp):
    slice upload + device transpose + all-gather.  Returns [slice buffer, 9 column tensors]; the caller keeps
    them alive until the library has copied the columns (skidgpu_set_particles_dev synchronises)."""
    dist, dev = reducer.dist, reducer.device
    n = len(pinit)
    per = -(-n // nranks)                       # equal chunks (all_gather_into_tensor), last one padded
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    raw = torch.from_numpy(np.ascontiguousarray(pinit).view(np.float32).reshape(n, 12))
    # host -> device on torch's own stream: pinned host blocks remember the streams that read them, and the
    # context's stream (owned by the library) may be gone by the time torch releases the host buffer
    mine = torch.zeros((per, 12), dtype=torch.float32, device=dev)
    if hi > lo:
        mine[:hi - lo].copy_(raw[lo:hi], non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    ctx = torch.cuda.stream(reducer.stream) if reducer.stream is not None else _nullcontext()
    cols = [mine]  # kept alive by the caller until the context's stream has been synchronised
    with ctx:
        full = torch.empty((nranks * per,), dtype=torch.float32, device=dev)
        for k in range(9):
            dist.all_gather_into_tensor(full, mine[:, k].contiguous())
            cols.append(full[:n].clone())
    reducer.bytes += 9 * 4 * n
    reducer.h2d_bytes = getattr(reducer, "h2d_bytes", 0) + 48 * (hi - lo)
    return cols


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def run_skid_sharded(sk, reducer, pinit, nGas, nDark, nStar, flags, rank, nranks, host=True, dev_ptrs=None):
    """The main.c stage script on a sharded snapshot.  sk: api.SkidGPU with set_shard/reduce cb applied.
    Returns (labels by iOrder, catalogue, nUnbound, nGroupBefore)."""
    from . import api
    f32 = lambda v: float(np.float32(v))
    tau = f32(flags["tau"])
    fCvg = f32(0.5 * tau)
    fScoop = f32(2.0 * tau)
    fStep = f32(0.5 * fCvg)
    z = f32(flags.get("z", 0.0))
    a32 = f32(1.0 / (1.0 + z))
    fCosmo = a32 * api.csmExp2Hub(a32, f32(flags["H0"]), f32(flags.get("Omega0", 1.0)), f32(flags.get("Lambda", 0.0)))
    sk.log = []
    if host and nranks > 1 and reducer.device.type == "cuda":
        # the snapshot is replicated on the devices but need not cross PCIe N times: every rank uploads its
        # 1/N slice of the host AoS, transposes it to SoA columns on the device and the columns are
        # all-gathered over NVLink (N x fewer host->device bytes per rank)
        cols = upload_sliced(reducer, pinit, rank, nranks)
        sk.set_particles_dev([t.data_ptr() for t in cols[1:]], len(pinit), nGas, nDark, nStar)  # copies and synchronises
        del cols  # released while the context's stream is still alive (NCCL recorded it on these blocks)
    elif host:
        sk.set_particles(pinit, nGas, nDark, nStar)
    else:
        sk.set_particles_dev(dev_ptrs, len(pinit), nGas, nDark, nStar)
    sk.smDensityInit(flags["nSmooth"], flags.get("bGasAndDark", False), False, want_arrays=False)
    sk.move(flags["fDensMin"], flags.get("fTempMax", api.FLT_MAX), api.FLT_MAX, fCvg, fStep)
    dev = reducer.device

    def gather_positions():
        if nranks == 1 or sk.nMove == 0:
            return
        px, py, pz, nm, lo, hi = sk.mover_arrays()
        sk.mask_unowned_movers()  # ownership is block-cyclic: the library zeroes what other ranks own
        for p in (px, py, pz):
            reducer.reduce_tensor(tensor_from_pointer(p, nm, 2, dev), 2)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    gather_positions()
    sk.kdFoF(tau)
    sk.microstep(5, f32(0.1 * fStep))
    gather_positions()
    sk.kdCalcCenter(fetch=False)
    return sk.kdUnbind(1.0, z, fCosmo, api.SPLINE, fScoop, False, api.INT_MAX, flags["nMembers"])
