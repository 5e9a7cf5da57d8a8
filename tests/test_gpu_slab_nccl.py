"""Slab decomposition over NCCL: one process per GPU (torchrun), neighbour exchange with torch.distributed send/recv.
Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,recut,exchange,c_abi", [(2, 0, 1, 0), (4, 0, 1, 0), (2, 1, 1, 0), (2, 0, 0, 0), (2, 0, 1, 1), (2, 0, 0, 1), (4, 0, 1, 1), (2, 1, 1, 1)])
def test_slabs_over_nccl_match_one_gpu(world, recut, exchange, c_abi):
    """c_abi = 1: the whole slab step behind the C ABI (ps_comm_init / ps_comm_set_slab / ps_comm_step: NCCL send / recv issued from
    C++ on the context's stream), else particlesolver_b200/slab.py over torch.distributed"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    port = 29700 + world + 10 * recut + 20 * exchange + 40 * c_abi
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "slab_nccl_worker.py"), "6" if recut else "5", str(recut), str(exchange), str(c_abi)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SLAB_NCCL_OK" in r.stdout


def test_cli_runs_the_dam_break_over_two_ranks_like_one(tmp_path):
    """psolver_cli --scene c5: the C++ host runs BASELINE config C5 without Python — two processes, one GPU each, NCCL behind the C ABI
    (ps_comm_*), the id handed over through a file — and ends where the undecomposed run of the same scene ends."""
    import json
    import numpy as np
    from test_slab_cpu import _match
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs, have {torch.cuda.device_count()}")
    cli = os.path.join(os.path.dirname(HERE), "particlesolver_b200", "psolver_cli")
    common = ["--app", "gpu", "--scene", "c5", "--planes", "16", "--ny", "40", "--nz", "48", "--steps", "6", "--vx", "12"]
    one = subprocess.run([cli, *common, "--ranks", "1", "--dump-final", str(tmp_path / "one.bin")], capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stdout + one.stderr
    procs = [subprocess.Popen([cli, *common, "--ranks", "2", "--rank", str(r), "--id-file", str(tmp_path / "id"), "--dump-final", str(tmp_path / f"r{r}.bin")],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = [json.loads(o[0].strip().splitlines()[-1]) for o in outs]
    ref = json.loads(one.stdout.strip().splitlines()[-1])
    n = 16 * 40 * 48
    assert ref["particles_total"] == n and all(l["particles_total"] == n for l in lines)          # nothing lost in the exchange
    assert sum(l["particles_owned"] for l in lines) == n and sum(l["migrated_out"] for l in lines) > 0 and all(l["ghosts"] > 0 for l in lines)
    whole = np.fromfile(tmp_path / "one.bin", np.float32).reshape(-1, 8)
    got = np.concatenate([np.fromfile(tmp_path / f"r{r}.bin", np.float32).reshape(-1, 8) for r in range(2)])
    _match(whole[:, :4], whole[:, 4:], got[:, :4], got[:, 4:], tol=5e-5)
    assert abs(lines[0]["kinetic_energy_total"] - ref["kinetic_energy_total"]) <= 1e-4 * ref["kinetic_energy_total"]
