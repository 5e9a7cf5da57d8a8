"""BASELINE.json's headline size (config C3: reference GPU scene 7 scaled to 100^3 = 1,000,000 PBF particles, 256^3 grid)
through size-independent properties — the oracle finishes only small cases in seconds:
  grid:   keys sorted, `index` a permutation, ties in ascending original index (stable), sorted copies are exact gathers,
          the dense cell table is the lower bound of every key, cellStart/cellEnd delimit runs of equal keys;
  lambda: neighbour counts of interior lattice particles equal those of the same lattice at a size the oracle handles
          (the neighbourhood of an interior particle does not depend on the block size), never above the cap of 500;
  step:   finite, deterministic (two runs bit-identical), graph replay == stage-by-stage calls, and the fluid statistics
          (mean neighbour count, kinetic energy per particle) follow the small-scene oracle run."""
import numpy as np
import pytest

import helpers as H
import particlesolver_b200 as psb

pytestmark = pytest.mark.gpu
N_SIDE = 100


def build(side=N_SIDE):
    return psb.ParticleSystem.scene("c3", grid=256, max_particles=side ** 3 + 1024, side=side)


def test_grid_properties_at_one_million_particles():
    ps = build()
    sol = ps.solver
    assert sol.n == N_SIDE ** 3
    sol.predict(1 / 60)
    sol.build_grid()
    h, idx = sol.download(psb.ARR_HASH), sol.download(psb.ARR_INDEX)
    assert np.all(h[1:] >= h[:-1])                                    # sorted
    assert np.array_equal(np.sort(idx), np.arange(sol.n, dtype=np.uint32))  # a permutation
    same = h[1:] == h[:-1]
    assert np.all(idx[1:][same] > idx[:-1][same])                     # stable: ties keep ascending original index
    pos, spos = sol.download(psb.ARR_POS), sol.download(psb.ARR_SORTED_POS)
    assert np.array_equal(spos, pos[idx])                             # exact gather
    # keys are the reference's hash of the positions (integration_kernel.cuh:187-203): floor(p / cell) & (grid - 1), x fastest
    gp = np.floor(pos[:, :3] / np.float32(0.5)).astype(np.int64) & 255
    keys = ((gp[:, 2] * 256 + gp[:, 1]) * 256 + gp[:, 0]).astype(np.uint32)
    assert np.array_equal(h, keys[idx])
    cb = sol.download(psb.ARR_CELL_BEGIN)
    probe = np.random.default_rng(0).integers(0, sol.num_cells + 1, 200000)
    assert np.array_equal(cb[probe], np.searchsorted(h, probe.astype(np.uint32), side="left").astype(np.uint32))
    cs, ce = sol.download(psb.ARR_CELL_START), sol.download(psb.ARR_CELL_END)
    occ = np.nonzero(cs != 0xFFFFFFFF)[0]
    assert np.array_equal(np.unique(h), occ.astype(np.uint32))
    assert np.array_equal(cs[occ], np.searchsorted(h, occ.astype(np.uint32), side="left")) and np.array_equal(ce[occ], np.searchsorted(h, occ.astype(np.uint32), side="right"))
    ps.close()


def test_interior_neighbour_counts_match_the_oracle_lattice():
    small = build(22)                                                # 10,648 particles: the oracle's size
    o = H.oracle_from_solver(small.solver)
    o.build_grid()
    o.solve_fluids()
    n_small = np.asarray(o.nn)
    spos = np.asarray(o.spos)[:, :3]
    lo, hi = spos.min(0) + 2.1, spos.max(0) - 2.1
    interior_small = np.all((spos > lo) & (spos < hi), axis=1)
    small.close()
    ps = build()
    sol = ps.solver
    sol.build_grid()
    sol.find_neighbors()
    nn = sol.download(psb.ARR_NUM_NEIGHBORS)
    sp = sol.download(psb.ARR_SORTED_POS)[:, :3]
    lo, hi = sp.min(0) + 2.1, sp.max(0) - 2.1
    interior = np.all((sp > lo) & (sp < hi), axis=1)
    assert interior.sum() > 700_000 and interior_small.sum() > 2000   # 92^3 of 100^3, 14^3 of 22^3
    assert nn.max() <= 500
    # the jittered lattice (+-0.0025) leaves every interior particle the same neighbour shell up to pairs at |r| = H +- jitter
    assert abs(nn[interior].mean() - n_small[interior_small].mean()) < 0.5
    assert set(np.unique(nn[interior])) <= set(range(int(n_small[interior_small].min()) - 4, int(n_small[interior_small].max()) + 5))
    rows = sol.download(psb.ARR_NEIGHBOR_ROWS)
    assert not np.any(rows == 0xFFFFFFFF)                             # no warp overflowed its neighbour list
    ps.close()


def test_steps_are_finite_deterministic_and_graph_equals_stages():
    out = []
    for mode in ("graph", "graph", "stages"):
        ps = build()
        sol = ps.solver
        for _ in range(3):
            if mode == "graph":
                sol.step(1 / 60)
            else:
                sol.begin_step()
                sol.predict(1 / 60)
                for it in range(5):
                    sol.build_grid()
                    sol.solve_contacts()
                    sol.solve_fluid()
                    sol.collide_world(it)
                    sol.solve_distance()
                    sol.solve_point()
                sol.update_velocity(1 / 60)
        out.append((sol.download(psb.ARR_POS), sol.download(psb.ARR_VEL)))
        ps.close()
    assert np.isfinite(out[0][0]).all() and np.isfinite(out[0][1]).all()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])   # run-to-run
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])   # graph replay == stage calls
    v = out[0][1][:, :3].astype(np.float64)
    ke = 0.5 * (v * v).sum() / out[0][0].shape[0]
    # the same scene at the oracle's size, 3 steps on the oracle: kinetic energy per particle of the expanding blob agrees
    # (surface-to-volume differs: 22^3 vs 100^3, so only the order of magnitude is size-independent)
    small = build(22)
    o = H.oracle_from_solver(small.solver)
    rng = np.random.default_rng(1)
    for _ in range(3):
        o.step(1 / 60, rng.uniform(0, 1, (5, 6)).astype(np.float32))
    vs = np.asarray(o.vel)[:, :3].astype(np.float64)
    ke_small = 0.5 * (vs * vs).sum() / vs.shape[0]
    small.close()
    assert 0.2 * ke_small < ke < 5 * ke_small, (ke, ke_small)
