"""Slab decomposition over NCCL: one process per GPU (torchrun), neighbour exchange with torch.distributed send/recv.
Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,recut,exchange", [(2, 0, 1), (4, 0, 1), (2, 1, 1), (2, 0, 0)])
def test_slabs_over_nccl_match_one_gpu(world, recut, exchange):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    port = 29700 + world + 10 * recut + 20 * exchange
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "slab_nccl_worker.py"), "6" if recut else "5", str(recut), str(exchange)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SLAB_NCCL_OK" in r.stdout
