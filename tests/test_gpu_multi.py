"""Sharded (one process per GPU, NCCL) pipeline vs the single-GPU pipeline.  Needs >= 2 GPUs;
on a 1-GPU box it is skipped (run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29511",
                        os.path.join(ROOT, "tools", "multi_check.py")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI-GPU PARITY OK" in r.stdout
