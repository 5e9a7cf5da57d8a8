"""Pins the oracle (oracle/gpu_step_oracle.c) against the reference itself: golden vectors dumped after every
wrapper call of one ParticleSystem::update by the reference's UNMODIFIED GPU sources on a B200
(tests/golden/make_golden.sh).  Runs on CPU.

Bars: integer grid output (sorted hash, sorted index, cellStart, cellEnd where valid, neighbour counts)
bit-exact; float stages within STAGE_ATOL when each stage starts from the reference's own inputs
(IEEE oracle vs the reference's -use_fast_math build); whole step without re-synchronisation within STEP_ATOL.
"""
import numpy as np
import pytest

import golden_io as G

STAGE_ATOL = 1e-5   # measured worst case 1.9e-6 (scene 1, distance stage)
STEP_ATOL = 1e-4    # SURVEY Appendix A.7 interior-particle bound after one full step
LAMBDA_ATOL = 2e-5  # absolute, lambda is O(0.1-1)


def _dev(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.abs(a - b).max()) if a.size else 0.0


@pytest.mark.parametrize("scene", G.SCENES)
def test_oracle_stage_by_stage_matches_reference_gpu(scene):
    g = G.load(scene)
    o = G.oracle_for(g)
    assert np.array_equal(o.occ, g["occurences"])  # updateOccurences, solver.cu:72-106
    o.predict(G.DT)
    assert _dev(o.pos, g["s0_predict_pos"]) <= STAGE_ATOL
    assert np.array_equal(o.prev.reshape(-1), g["s0_prev"])
    o.pos[:] = g["s0_predict_pos"].reshape(-1, 4)
    for it in range(o.iterations):
        T = f"s0_i{it}_"
        o.calc_hash()
        if it == 0:
            assert np.array_equal(o.hash, g[T + "hash_unsorted"])
        o.sort()
        o.reorder()
        assert np.array_equal(o.hash, g[T + "hash"]), f"iter {it}: sorted keys"
        assert np.array_equal(o.index, g[T + "index"]), f"iter {it}: sorted order"
        assert np.array_equal(o.cell_start, g[T + "cell_start"]), f"iter {it}: cellStart"
        valid = o.cell_start != 0xFFFFFFFF
        assert np.array_equal(o.cell_end[valid], g[T + "cell_end"][valid]), f"iter {it}: cellEnd"
        if it == 0:
            assert np.array_equal(o.spos.reshape(-1), g[T + "sorted_pos"])
            assert np.array_equal(o.sw, g[T + "sorted_w"]) and np.array_equal(o.sphase, g[T + "sorted_phase"])
        o.collide()
        solid = o.sphase >= 2
        assert np.array_equal(o.nn[solid], g[T + "collide_nn"][solid]), f"iter {it}: contact counts"
        assert _dev(o.pos, g[T + "collide_pos"]) <= STAGE_ATOL, f"iter {it}: contacts"
        o.pos[:] = g[T + "collide_pos"].reshape(-1, 4)
        o.solve_fluids()
        fluid = o.sphase == 0
        assert np.array_equal(o.nn[fluid], g[T + "fluid_nn"][fluid]), f"iter {it}: fluid neighbour counts"
        assert _dev(o.lam[fluid], g[T + "lambda"][fluid]) <= LAMBDA_ATOL, f"iter {it}: lambda"
        assert _dev(o.pos, g[T + "fluid_pos"]) <= STAGE_ATOL, f"iter {it}: fluid"
        o.lam[:] = g[T + "lambda"]
        o.nn[:] = g[T + "fluid_nn"]
        o.pos[:] = g[T + "fluid_pos"].reshape(-1, 4)
        o.collide_world(g[T + "rands"])
        assert _dev(o.pos, g[T + "world_pos"]) <= STAGE_ATOL, f"iter {it}: world"
        o.pos[:] = g[T + "world_pos"].reshape(-1, 4)
        o.solve_distance()
        assert _dev(o.pos, g[T + "dist_pos"]) <= STAGE_ATOL, f"iter {it}: distance"
        o.pos[:] = g[T + "dist_pos"].reshape(-1, 4)
        o.solve_point()
        assert np.array_equal(o.pos.reshape(-1), g[T + "point_pos"]), f"iter {it}: pins are exact copies"
    o.calc_velocity(G.DT)
    assert _dev(o.vel, g["s0_final_vel"]) <= STAGE_ATOL * 60
    assert o.dist_nonprefix == 0  # rank == key for every built-in scene


@pytest.mark.parametrize("scene", G.SCENES)
def test_oracle_whole_step_matches_reference_gpu(scene):
    g = G.load(scene)
    o = G.oracle_for(g)
    rands = np.stack([g[f"s0_i{it}_rands"] for it in range(o.iterations)])
    o.step(G.DT, rands)
    if scene == "9":
        # ropes resting on the pinned sphere: the friction term of the contact pass switches between its static and kinetic
        # branch at lt = S_FRICTION * dist (integration_kernel.cuh:450-459); for two rope particles of 3,653 the last-bit
        # difference accumulated over the iterations flips that branch.  Every stage agrees to 1e-5 from identical inputs
        # (test above); here: all but a handful within STEP_ATOL, those within 1e-3.
        d = np.abs(o.pos.reshape(-1, 4) - g["s0_final_pos"].reshape(-1, 4)).max(1)
        assert (d > STEP_ATOL).sum() <= 4 and d.max() <= 1e-3, (int((d > STEP_ATOL).sum()), float(d.max()))
        return
    assert _dev(o.pos, g["s0_final_pos"]) <= STEP_ATOL
    assert _dev(o.vel, g["s0_final_vel"]) <= STEP_ATOL * 60
    # the integer grid of the LAST iteration still matches bit for bit when the float error stayed below a cell edge
    T = f"s0_i{o.iterations - 1}_"
    if np.array_equal(o.hash, g[T + "hash"]):
        assert np.array_equal(o.index, g[T + "index"])


def test_scene_counts_match_survey_appendix_c():
    expect = {"1": (33, 32, 1), "2": (144, 264, 12), "3": (5632, 0, 0), "7": (5324, 0, 0), "8": (5822, 3735, 636)}
    for s, (n, m, p) in expect.items():
        g = G.load(s)
        assert int(g["meta_n"]) == n and g["dist_rest"].size == m and g["point_idx"].size == p


def test_sort_is_stable_and_handles_edge_cases():
    import oracle_py as orc
    rng = np.random.default_rng(3)
    for n in (1, 2, 255, 256, 257, 5000):
        keys = rng.integers(0, 7, size=n).astype(np.uint32)  # many ties
        idx = np.arange(n, dtype=np.uint32)
        k2, i2 = keys.copy(), idx.copy()
        orc.lib().or_sort(k2.ctypes.data, i2.ctypes.data, n)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(i2, order.astype(np.uint32)) and np.array_equal(k2, keys[order])


def test_oracle_omega_scales_the_averaged_deltas():
    """or_set_omega: the PBF delta-p and the distance-constraint correction are linear in omega (omega = 1 is the pinned reference
    path, every golden test above runs it); contacts are not (friction sees the scaled delta), so they are only checked to move."""
    import oracle_py as orc
    rng = np.random.default_rng(3)
    n = 4000
    pos = np.ones((n, 4), np.float32)
    pos[:, :3] = rng.uniform(2.0, 8.0, size=(n, 3)).astype(np.float32)
    p = orc.make_params(grid=(64, 64, 64))
    deltas = {}
    for om in (1.0, 0.5, 1.5):
        o = orc.OracleSystem(p, pos, np.zeros((n, 4), np.float32), np.ones(n, np.float32), np.zeros(n, np.int32), np.full(n, 8.0, np.float32))
        o.omega = om
        o.build_grid()
        o.solve_fluids()
        deltas[om] = (o.pos - pos)[:, :3].astype(np.float64)
    scale = np.abs(deltas[1.0]).max()
    assert scale > 1e-3
    for om in (0.5, 1.5):
        assert np.abs(deltas[om] - om * deltas[1.0]).max() <= 2e-6 * max(1.0, scale)
    # distance constraints: a stretched chain
    m = 50
    cpos = np.ones((m, 4), np.float32)
    cpos[:, 0] = np.arange(m) * 0.7
    cpos[:, 1] = 5.0
    idx = np.stack([np.arange(m - 1), np.arange(1, m)], 1).astype(np.uint32)
    rest = np.full(m - 1, 0.5, np.float32)
    d = {}
    for om in (1.0, 0.5):
        o = orc.OracleSystem(p, cpos, np.zeros((m, 4), np.float32), np.ones(m, np.float32), np.full(m, 2, np.int32), np.ones(m, np.float32),
                             dist_idx=idx, dist_rest=rest)
        o.omega = om
        o.solve_distance()
        d[om] = (o.pos - cpos)[:, :3].astype(np.float64)
    assert np.abs(d[1.0]).max() > 1e-2 and np.abs(d[0.5] - 0.5 * d[1.0]).max() <= 4e-6  # float32 positions near x = 35: ulp 3.8e-6
    orc.lib().or_set_omega(1.0)
