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


def test_block_cyclic_position_exchange_layout():
    """numpy mirrors of k_owned_ids / k_pack_owned / k_unpack_all (csrc/move.cu): every mover has exactly one
    owner, and packing the owned blocks + all-gather + unpacking restores every mover's value on every rank."""
    for m in (1, 4095, 4096, 4097, 50000, 3 * parallel.OWN_BLOCK * 8 + 17):
        for w in (1, 2, 3, 8):
            owners = np.zeros(m, np.int64)
            for r in range(w):
                ids = parallel.owned_ids(m, r, w)
                owners[ids] += 1
                # the library's owned_count: whole blocks, the last one may be short
                nb = -(-m // parallel.OWN_BLOCK)
                cnt = sum((m - j * parallel.OWN_BLOCK) if j == nb - 1 else parallel.OWN_BLOCK for j in range(r, nb, w))
                assert len(ids) == cnt
            assert np.all(owners == 1)
            x = np.random.default_rng(m + w).random(m).astype(np.float32)
            slots = [parallel.pack_owned(x, r, w) for r in range(w)]
            assert len({len(s) for s in slots}) == 1          # equal-sized all-gather blocks
            assert np.array_equal(parallel.unpack_all(slots, m, w), x)


def test_distributed_sort_pieces_equal_global_stable_sort():
    """numpy mirror of dist_sort_pairs (csrc/tree.cu): key ranges from a regular sample, stable local sorts,
    concatenation in rank order == one global stable sort, including heavy ties (group labels as keys)."""
    rng = np.random.default_rng(5)
    n = 1 << 19
    for keys in (rng.integers(0, 1 << 62, n, dtype=np.int64).astype(np.uint64),
                 rng.integers(0, 37, n).astype(np.uint64),                       # few distinct keys
                 np.zeros(n, np.uint64)):                                        # all equal
        want = np.argsort(keys, kind="stable")
        for w in (2, 3, 8):
            pieces = [parallel.dist_sort_piece(keys, r, w) for r in range(w)]
            assert np.array_equal(np.concatenate(pieces), want)
            if len(np.unique(keys)) > 1000:                   # balanced when the keys allow it
                assert max(len(p) for p in pieces) < 1.1 * n / w
