#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sharded pipeline
(skid_b200/parallel.py) must give the same groups as the single-GPU pipeline on the same snapshot,
and on the demo the reference's golden group count.  Exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_demo_input, DEMO  # noqa: E402
from oracle.refdump import canonical_labels  # noqa: E402
from skid_b200 import api, parallel, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cases = []
    p, ng, nd, ns, _ = load_demo_input()
    cases.append(("demo", p, ng, nd, ns, dict(DEMO)))
    s = synth.make_box(1 << 18, seed=1234, kind="dark")
    cases.append(("dark2^18", s["pinit"], s["nGas"], s["nDark"], s["nStar"], s["flags"]))
    s = synth.make_box(1 << 16, seed=7, kind="gasdark")
    cases.append(("gasdark2^16", s["pinit"], s["nGas"], s["nDark"], s["nStar"], s["flags"]))
    s = synth.make_box(1 << 20, seed=9, kind="massive")   # large enough for the distributed sorts; massive groups
    cases.append(("massive2^20", s["pinit"], s["nGas"], s["nDark"], s["nStar"], s["flags"]))
    # the demo once more through the callback shim (skidgpu_set_reduce_cb) instead of the library's own NCCL
    cases.append(("demo/callback-shim", p, ng, nd, ns, dict(DEMO)))
    ok = True
    for name, p, ng, nd, ns, fl in cases:
        sk = api.SkidGPU((fl["period"],) * 3, (0.0,) * 3, bPeriodic=True, device=local)
        red = parallel.Reducer(dist, dev, sk.stream())
        if name.endswith("callback-shim"):
            sk.set_shard(rank, world)
            sk.set_reduce_cb(red.cb)
        else:
            parallel.init_comm(sk, dist, rank, world)
        grp, cat, nUnb, nBefore = parallel.run_skid_sharded(sk, red, p, ng, nd, ns, fl, rank, world, host=True)
        log = [l for l in sk.log if l[0] == 0]
        cbytes, ccalls = sk.comm_bytes()
        sk.close()
        # all ranks must agree bit for bit
        t = torch.from_numpy(grp.astype(np.int64)).to(dev)
        t0 = t.clone()
        dist.broadcast(t0, 0)
        same_across = bool(torch.equal(t, t0))
        if rank == 0:
            ref = api.run_skid(p, ng, nd, ns, device=local, want_arrays=False, **fl)
            same = float(np.mean(canonical_labels(ref["grp"]) == canonical_labels(grp)))
            line = (f"{name}: world={world} groups {len(cat) - 1} vs single {ref['nGroup']}, before {nBefore} vs "
                    f"{ref['nGroupBefore']}, unbound {nUnb} vs {ref['nUnbound']}, same-group {same:.6f}, ittr {len(log)} vs "
                    f"{ref['nIttr']}, exchanges {ccalls} ({cbytes / 1e6:.1f} MB), ranks identical {same_across}")
            print(line, flush=True)
            good = (len(cat) - 1 == ref["nGroup"] and nBefore == ref["nGroupBefore"] and same >= 0.9999
                    and len(log) == ref["nIttr"])
            if name.startswith("demo"):
                good = good and nBefore == 120 and len(cat) - 1 == 68
            ok = ok and good
        ok = ok and same_across
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        print("MULTI-GPU PARITY FAILED", flush=True)
        return 1
    if rank == 0:
        print("MULTI-GPU PARITY OK", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
