"""Slab decomposition, host side (CPU): cuts, halo selection, ghost bookkeeping, migration and the neighbour exchange of
particlesolver_b200.slab, with the oracle as the compute engine.  In-process lock-step (LocalCluster, 2 and 3 slabs) and
two real processes over gloo (DistComm); each must reproduce the single-domain oracle run."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle_py as orc
from particlesolver_b200 import slab
from slab_oracle_engine import OracleEngine

DT = 1.0 / 60.0


def _scene(nx=30, ny=8, nz=8):
    """A fluid block moving in +x at 6 units/s (0.1 per step: particles cross the cut planes) on a 64^3 grid."""
    pos, vel, w, ros, phase = slab.dam_break_block(nx, ny, nz, origin=(2.3, 1.3, 2.3))
    vel[:, 0] = 6.0
    p = orc.make_params(grid=(64, 64, 64), min_b=(0, 0, 0), max_b=(40, 30, 12))
    return p, pos, vel, w, phase, ros


def _single(p, pos, vel, w, phase, ros, steps):
    o = orc.OracleSystem(p, pos, vel, w, phase, ros, iterations=5)
    rands = np.full((5, 6), 0.5, np.float32)
    for _ in range(steps):
        o.step(DT, rands)
    return o.pos, o.vel


def _match(pos_a, vel_a, pos_b, vel_b, tol):
    """same particle SET: nearest-neighbour bijection within tol"""
    from scipy.spatial import cKDTree
    assert pos_a.shape == pos_b.shape
    d, idx = cKDTree(pos_b[:, :3]).query(pos_a[:, :3])
    assert d.max() <= tol, f"max position mismatch {d.max():.3e}"
    assert np.unique(idx).size == idx.size
    assert np.abs(vel_a[:, :3] - vel_b[idx, :3]).max() <= tol * 60 * 1.5


def _split(cuts, pos, *arrays):
    out = []
    for r in range(len(cuts) - 1):
        sel = (pos[:, 0] >= cuts[r]) & (pos[:, 0] < cuts[r + 1])
        out.append(tuple(a[sel] for a in (pos, *arrays)))
    return out


@pytest.mark.parametrize("exchange_lambda", [True, False], ids=["lambda-exchanged", "lambda-local"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_local_cluster_matches_single_domain(nranks, exchange_lambda):
    p, pos, vel, w, phase, ros = _scene()
    steps = 4
    ref_pos, ref_vel = _single(p, pos, vel, w, phase, ros, steps)
    cuts = slab.quantile_cuts(pos[:, 0], nranks)
    engines = [OracleEngine(p, a, b, c, d, e) for a, b, c, d, e in _split(cuts, pos, vel, w, phase, ros)]
    n0 = [e.n_owned for e in engines]
    cl = slab.LocalCluster(engines, cuts, exchange_lambda=exchange_lambda)
    assert cl.doms[0].halo == (2.25 if exchange_lambda else 4.5)
    for _ in range(steps):
        cl.step(DT)
    got_pos = np.concatenate([e.pos[:e.n_owned] for e in engines])
    got_vel = np.concatenate([e.vel for e in engines])
    assert got_pos.shape[0] == pos.shape[0]                       # nothing lost, nothing duplicated
    assert np.isfinite(got_pos).all()                             # every ghost received its lambda (the engine poisons them)
    assert sum(d.stats["migrated_out"] for d in cl.doms) > 0      # the block moves: particles did change owner
    assert [e.n_owned for e in engines] != n0
    assert all(d.stats["ghosts"] > 0 for d in cl.doms)
    # owned particles stay inside their slab (up to the drift allowance of one step)
    for d in cl.doms:
        x = d.eng.pos[:d.eng.n_owned, 0]
        assert (x >= d.x_lo - 0.25).all() and (x < d.x_hi + 0.25).all()
    _match(ref_pos, ref_vel, got_pos, got_vel, tol=2e-5)


def test_cuts():
    c = slab.uniform_cuts(0.0, 80.0, 4)
    assert c[0] == -np.inf and c[-1] == np.inf and c[1:-1] == [20.0, 40.0, 60.0]
    x = np.linspace(0, 1, 1001)
    q = slab.quantile_cuts(x, 4)
    assert np.allclose(q[1:-1], [0.25, 0.5, 0.75])
    assert slab.quantile_cuts(x, 1) == [-np.inf, np.inf]


def test_dam_break_block_is_rank_count_independent():
    whole = slab.dam_break_block(12, 3, 4)[0]
    parts = [slab.dam_break_block(12, 3, 4, ix0=a, ix1=b)[0] for a, b in ((0, 5), (5, 12))]
    a = whole[np.lexsort(whole[:, :3].T)]
    b = np.concatenate(parts); b = b[np.lexsort(b[:, :3].T)]
    assert np.array_equal(a, b)
    assert np.abs(whole[:, :3] - np.round((whole[:, :3] - 0.3125) / 0.625) * 0.625 - 0.3125).max() <= 0.0025 + 1e-6


def _gloo_worker(rank, world, port, steps, out_dir, exchange_lambda=True):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p, pos, vel, w, phase, ros = _scene()
    cuts = slab.quantile_cuts(pos[:, 0], world)
    mine = _split(cuts, pos, vel, w, phase, ros)[rank]
    eng = OracleEngine(p, *mine)
    dom = slab.SlabDomain(eng, rank, world, cuts, comm=slab.DistComm(eng), exchange_lambda=exchange_lambda)
    for _ in range(steps):
        dom.step(DT)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=eng.pos[:eng.n_owned], vel=eng.vel, migrated=dom.stats["migrated_out"],
             sent=dom.comm.bytes_sent, ghosts=dom.stats["ghosts"])
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange_lambda", [True, False], ids=["lambda-exchanged", "lambda-local"])
def test_two_processes_over_gloo_match_single_domain(tmp_path, exchange_lambda):
    import torch.multiprocessing as mp
    steps, world = 5, 2
    port = 29500 + (os.getpid() % 2000) + (2000 if exchange_lambda else 0)
    mp.spawn(_gloo_worker, args=(world, port, steps, str(tmp_path), exchange_lambda), nprocs=world, join=True)
    p, pos, vel, w, phase, ros = _scene()
    ref_pos, ref_vel = _single(p, pos, vel, w, phase, ros, steps)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    got_pos = np.concatenate([z["pos"] for z in parts])
    got_vel = np.concatenate([z["vel"] for z in parts])
    assert sum(int(z["migrated"]) for z in parts) > 0 and all(int(z["sent"]) > 0 for z in parts)
    _match(ref_pos, ref_vel, got_pos, got_vel, tol=2e-5)


def test_balanced_cuts_properties():
    hist = np.zeros(100, np.int64)
    hist[10:30] = 50      # all particles in [10, 30)
    old = [-np.inf, 40.0, 70.0, np.inf]
    new = slab.balanced_cuts(hist, 0.0, 100.0, old, min_width=5.0, max_shift=8.0)
    assert new[0] == -np.inf and new[-1] == np.inf
    assert new[1] == 32.0 and new[2] == 62.0                      # wanted 16.67 and 23.33, limited to a shift of 8 per re-cut
    for _ in range(10):                                           # repeated re-cuts converge to the quantiles, minimum width kept
        new = slab.balanced_cuts(hist, 0.0, 100.0, new, min_width=5.0, max_shift=8.0)
    assert abs(new[1] - (10 + 20 / 3)) < 1e-9 and abs(new[2] - (10 + 40 / 3)) < 1e-9 and new[2] - new[1] >= 5.0
    tight = slab.balanced_cuts(hist, 0.0, 100.0, new, min_width=9.0, max_shift=8.0)
    assert tight[2] - tight[1] >= 9.0 - 1e-12
    assert slab.balanced_cuts(hist, 0.0, 100.0, [-np.inf, np.inf], 5.0, 8.0) == [-np.inf, np.inf]


def test_recut_keeps_the_result_and_balances_the_slabs():
    """cut planes follow the moving block (re-cut every step from the summed x-histograms): same particles as the single
    domain, better balanced than with fixed planes"""
    p, pos, vel, w, phase, ros = _scene(nx=36)
    steps, nranks = 6, 3
    ref_pos, ref_vel = _single(p, pos, vel, w, phase, ros, steps)
    # deliberately unbalanced start: 2/3 of the block in the first slab
    x = pos[:, 0]
    cuts = [-np.inf, float(np.quantile(x, 0.66)), float(np.quantile(x, 0.83)), np.inf]
    counts = {}
    for recut in (0, 1):
        engines = [OracleEngine(p, a, b, c, d, e) for a, b, c, d, e in _split(cuts, pos, vel, w, phase, ros)]
        cl = slab.LocalCluster(engines, cuts, recut_every=recut, recut_range=(0.0, 40.0), recut_bins=512)
        for _ in range(steps):
            cl.step(DT)
        got_pos = np.concatenate([e.pos[:e.n_owned] for e in engines])
        got_vel = np.concatenate([e.vel for e in engines])
        _match(ref_pos, ref_vel, got_pos, got_vel, tol=2e-5)
        counts[recut] = [e.n_owned for e in engines]
        if recut:
            assert cl.doms[0].stats["recuts"] == steps - 1 and cl.doms[0].cuts != [float(c) for c in cuts]
            assert all(d.cuts == cl.doms[0].cuts for d in cl.doms)
    imbalance = {k: max(v) / (sum(v) / len(v)) for k, v in counts.items()}
    assert imbalance[1] < imbalance[0] - 0.3, (counts, imbalance)


def _gloo_recut_worker(rank, world, port, steps, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p, pos, vel, w, phase, ros = _scene(nx=36)
    cuts = [-np.inf, float(np.quantile(pos[:, 0], 0.7)), np.inf]
    mine = _split(cuts, pos, vel, w, phase, ros)[rank]
    eng = OracleEngine(p, *mine)
    dom = slab.SlabDomain(eng, rank, world, cuts, comm=slab.DistComm(eng), recut_every=1, recut_range=(0.0, 40.0), recut_bins=512)
    for _ in range(steps):
        dom.step(DT)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=eng.pos[:eng.n_owned], vel=eng.vel, cuts=np.array(dom.cuts[1:-1]), recuts=dom.stats["recuts"])
    dist.destroy_process_group()


def test_recut_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    steps, world = 5, 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gloo_recut_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    p, pos, vel, w, phase, ros = _scene(nx=36)
    ref_pos, ref_vel = _single(p, pos, vel, w, phase, ros, steps)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    _match(ref_pos, ref_vel, np.concatenate([z["pos"] for z in parts]), np.concatenate([z["vel"] for z in parts]), tol=2e-5)
    assert np.array_equal(parts[0]["cuts"], parts[1]["cuts"]) and int(parts[0]["recuts"]) == steps - 1
    n = [z["pos"].shape[0] for z in parts]
    assert max(n) / (sum(n) / 2) < 1.25, n                           # started at 70 / 30
