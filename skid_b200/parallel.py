"""Multi-GPU plumbing for tests and bench.py: one process per GPU (torchrun).  The exchanges of the sharded
pipeline are issued by the LIBRARY with NCCL over NVLink on its own stream (csrc/dist.cu);
torch.distributed only carries the 128-byte NCCL id to the ranks (`init_comm`) and, in bench.py, the synthetic
snapshot.  host/skid -gpus N does the same from C with one host thread per GPU.

What is shared between the ranks: the sorts behind the tree builds (every rank sorts one key range of the
replicated input, the pieces are all-gathered), kNN queries (contiguous Morton ranges; fBall2 all-gathered, f64
density partials summed), movers (block-cyclic over the Morton-ordered mover list: contiguous ranges leave whole
halos on one rank and the others wait for it at every step) and groups for unbinding (g % nranks).
What is replicated: particles, tree boxes, scatterers, FoF, catalogue bookkeeping.

`Reducer` is the callback shim (skidgpu_set_reduce_cb) kept for tests; with the gloo backend on host buffers it
runs on CPU (tests/test_parallel_cpu.py), as do the numpy mirrors of the library's ownership and distributed-sort
layouts below.
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


OWN_BLOCK = 4096  # csrc/move.cu: movers are owned in blocks of this many consecutive (Morton-ordered) movers


def owned_ids(m, rank, nranks):
    """Mirror of k_owned_ids / owned_count (csrc/move.cu): mover ids owned by `rank`, block-cyclic."""
    ids = np.arange(m, dtype=np.int64)
    return ids[(ids // OWN_BLOCK) % nranks == rank]


def pack_owned(x, rank, nranks):
    """Mirror of k_pack_owned: this rank's slot of the position exchange buffer (one plane)."""
    m = len(x)
    nb = -(-m // OWN_BLOCK)
    per_blocks = -(-nb // nranks)
    slot = np.zeros(per_blocks * OWN_BLOCK, x.dtype)
    ids = owned_ids(m, rank, nranks)
    lb = (ids // OWN_BLOCK) // nranks
    slot[lb * OWN_BLOCK + ids % OWN_BLOCK] = x[ids]
    return slot


def unpack_all(slots, m, nranks):
    """Mirror of k_unpack_all: mover array from the all-gathered slots (slots[r] = rank r's plane)."""
    ids = np.arange(m, dtype=np.int64)
    blk = ids // OWN_BLOCK
    r = blk % nranks
    return np.stack(slots)[r, (blk // nranks) * OWN_BLOCK + ids % OWN_BLOCK]


def dist_sort_piece(keys, rank, nranks, samples=1 << 16):
    """Mirror of dist_sort_pairs (csrc/tree.cu): the (stable) sorted piece of the permutation that `rank`
    contributes; the concatenation over ranks equals one global stable sort."""
    n = len(keys)
    if nranks == 1 or n < 4 * samples:
        return np.argsort(keys, kind="stable") if rank == 0 else np.zeros(0, np.int64)
    samp = np.sort(keys[np.arange(samples) * (n // samples)], kind="stable")
    lo = 0 if rank == 0 else samp[samples * rank // nranks]
    sel = keys >= lo
    if rank < nranks - 1:
        sel &= keys < samp[samples * (rank + 1) // nranks]
    idx = np.nonzero(sel)[0]
    return idx[np.argsort(keys[idx], kind="stable")]


def canonical_labels(grp):
    """Relabel a partition so that groups are numbered by ascending smallest member index (0 stays 0): two
    catalogues describe the same partition iff their canonical labels are equal (bench.py's sharded-vs-single
    `parity`)."""
    grp = np.asarray(grp)
    out = np.zeros_like(grp)
    idx = np.nonzero(grp)[0]
    if len(idx) == 0:
        return out
    g = grp[idx]
    big = np.iinfo(np.int64).max
    first = np.full(int(g.max()) + 1, big, np.int64)
    np.minimum.at(first, g, idx)
    used = np.nonzero(first != big)[0]
    order = used[np.argsort(first[used], kind="stable")]
    remap = np.zeros(int(g.max()) + 1, np.int64)
    remap[order] = np.arange(1, len(order) + 1)
    out[idx] = remap[g]
    return out


def init_comm(sk, dist, rank, nranks):
    """Give the context its NCCL communicator: rank 0 makes the id, torch.distributed carries it."""
    from . import api
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sk.comm_init(box[0], rank, nranks)


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


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


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
                # NCCL has no uint8 min/max guarantee across versions: widen flags through int32.  The conversions
                # must run on the context's stream like the collective itself: on torch's default stream they raced
                # with the library's kernels before and after the exchange (the touched flags of the initial cut -
                # an intermittent 131 instead of 120 FoF groups on the demo through this shim)
                ctx = torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()
                with ctx:
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


def run_skid_sharded(sk, reducer, pinit, nGas, nDark, nStar, flags, rank, nranks, host=True, dev_ptrs=None,
                     fetch=True, out_grp=None, out_cat=None):
    """The main.c stage script on a sharded snapshot.  sk: api.SkidGPU with comm_init (or set_shard + reduce cb)
    applied.  Returns (labels by iOrder, catalogue, nUnbound, nGroupBefore)."""
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
    if host:
        # with a communicator the library uploads this rank's 1/N slice of the host AoS and all-gathers the
        # slices over NVLink (csrc/api.cu: skidgpu_set_particles)
        sk.set_particles(pinit, nGas, nDark, nStar)
    else:
        sk.set_particles_dev(dev_ptrs, len(pinit), nGas, nDark, nStar)
    sk.smDensityInit(flags["nSmooth"], flags.get("bGasAndDark", False), False, want_arrays=False)
    sk.move(flags["fDensMin"], flags.get("fTempMax", api.FLT_MAX), api.FLT_MAX, fCvg, fStep,
            bNoPrune=flags.get("bNoPrune", False))
    sk.kdFoF(tau)                      # the library exchanged the converged positions at the end of move
    sk.microstep(5, f32(0.1 * fStep))  # ... and the micro-stepped ones here
    sk.kdCalcCenter(fetch=False)
    return sk.kdUnbind(1.0, z, fCosmo, api.SPLINE, fScoop, False, flags.get("nMaxMembers", api.INT_MAX),
                       flags["nMembers"], fetch=fetch, out_grp=out_grp, out_cat=out_cat)
