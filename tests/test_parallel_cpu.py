"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in skid_b200/parallel.py: the reduce
callback the library calls at its agreement points, shard ranges, the all-gather of owned mover
ranges and the label / catalogue merge rules of the sharded unbinding."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skid_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        red = parallel.Reducer(dist, dev)
        res = {}
        # (1) the C callback on raw host pointers, every dtype/op the library uses
        n = 1000
        rng = np.random.default_rng(100 + rank)
        # fScatDens: float bits as int32, min; "no hit" = +inf bits
        t = np.array([np.float32(3.5 + rank).view(np.int32) if rank == 0 else 0x7f800000], np.int32)
        assert red.cb(None, t.ctypes.data, 1, 0, 0) == 0
        res["T"] = float(t.view(np.float32)[0])
        touched = (rng.random(n) < 0.3).astype(np.uint8)
        mine = touched.copy()
        assert red.cb(None, touched.ctypes.data, n, 1, 1) == 0
        res["touched_superset"] = bool(np.all(touched >= mine)) and set(np.unique(touched)) <= {0, 1}
        res["touched_sum"] = int(touched.sum())
        # sharded kNN: each rank fills only its query range; sum = full arrays on every rank
        lo, hi = parallel.shard_range(n, rank, world)
        full_b2 = np.linspace(1, 2, n).astype(np.float32)
        full_rho = np.linspace(5, 9, n)
        b2 = np.zeros(n, np.float32)
        b2[lo:hi] = full_b2[lo:hi]
        rho64 = full_rho * (0.25 if rank == 0 else 0.75)          # partial scatter sums
        assert red.cb(None, b2.ctypes.data, n, 2, 2) == 0 and red.cb(None, rho64.ctypes.data, n, 3, 2) == 0
        res["b2_ok"] = bool(np.array_equal(b2, full_b2))
        res["rho_ok"] = bool(np.allclose(rho64, full_rho, rtol=1e-15))
        # active count, sum
        act = np.array([7 + rank], np.int32)
        red.cb(None, act.ctypes.data, 1, 0, 2)
        res["active"] = int(act[0])
        # (2) all-gather of owned mover ranges (positions before FoF)
        m = 777
        truth = np.arange(m, dtype=np.float32) * 0.5
        lo, hi = parallel.shard_range(m, rank, world)
        x = torch.full((m,), -99.0)
        x[lo:hi] = torch.from_numpy(truth[lo:hi])
        red.allgather_owned([x], lo, hi)
        res["gather_ok"] = bool(np.array_equal(x.numpy(), truth))
        # (3) merge rules of the sharded unbinding: labels by min, catalogue rows by sum with zeros
        labels = torch.tensor([0, 1, 1, 2, 2, 2, 3], dtype=torch.int32)
        if rank == 1 % world:      # owner of group 1 unbinds particle 2
            labels[2] = 0
        if rank == 2 % world:      # owner of group 2 unbinds particle 3
            labels[3] = 0
        parallel.merge_labels(dist, labels)
        res["labels"] = labels.tolist()
        rows = torch.zeros(4, 8)
        for g in range(1, 4):
            if g % world == rank:
                rows[g] = float(g)
        red.reduce_tensor(rows, 2)
        res["rows_ok"] = bool(all(torch.all(rows[g] == float(g)) for g in range(1, 4)))
        res["calls"] = red.calls
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_shard_range_partition():
    for n in (0, 1, 7, 1000, 12190):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        res = out[r]
        assert res["T"] == 3.5                      # min over ranks, ignoring the "no hit" sentinel
        assert res["touched_superset"] and res["b2_ok"] and res["rho_ok"] and res["gather_ok"] and res["rows_ok"]
        assert res["active"] == 15
        assert res["labels"] == [0, 1, 0, 0, 2, 2, 3]
    assert out[0]["touched_sum"] == out[1]["touched_sum"]
