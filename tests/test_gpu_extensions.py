"""The solver parts north_star names that the reference does not contain — 3-D shape matching (K12), XSPH viscosity and
vorticity confinement (K13) — against float64 restatements (oracle/extensions_oracle.py).  PARITY UNPINNED: no reference
implementation exists; tolerances: positions 2e-4 (float32, fast-math, 20 quaternion iterations vs an SVD),
velocities 2e-4 relative to the largest velocity change."""
import os
import sys

import numpy as np
import pytest

import particlesolver_b200 as psb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import extensions_oracle as ext  # noqa: E402

pytestmark = pytest.mark.gpu


def rot(axis, angle):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def make_solver(pos, inv_mass=None, phase=None, vel=None, grid=64, lists=256, bounds=((-50, 0, -50), (50, 200, 50))):
    p = psb.default_params()
    p.grid_size[:] = (grid, grid, grid)
    p.min_bounds[:], p.max_bounds[:] = bounds
    p.neighbor_list_rows = lists
    n = pos.shape[0]
    sol = psb.Solver(p, max_particles=max(n, 1024))
    pos4 = np.concatenate([pos, np.ones((n, 1))], 1).astype(np.float32)
    vel4 = np.zeros((n, 4), np.float32)
    if vel is not None:
        vel4[:, :3] = vel
    sol.append(pos4, vel4, inv_mass if inv_mass is not None else np.ones(n), np.ones(n) * 1.5, phase if phase is not None else np.zeros(n))
    return sol


def cube(nx, ny, nz, spacing=0.5, origin=(0, 20, 0)):
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    return g * spacing + np.asarray(origin, np.float64)


def set_positions(sol, x):
    p4 = sol.download(psb.ARR_POS)
    p4[:, :3] = x
    sol.upload(psb.ARR_POS, p4)


@pytest.mark.parametrize("angle,shape", [(0.3, (4, 4, 4)), (1.0, (4, 4, 4)), (2.9, (3, 3, 3)), (1.2, (5, 5, 1)), (0.7, (2, 1, 1))])
def test_shape_matching_recovers_rigid_motions(angle, shape):
    rng = np.random.default_rng(3)
    x0 = cube(*shape)
    n = x0.shape[0]
    w = rng.uniform(0.5, 2.0, n)
    sol = make_solver(x0, inv_mass=w, phase=np.full(n, psb.RIGID))
    body = sol.add_rigid_body(np.arange(n))
    assert sol.num_rigid_bodies == 1
    mass = 1.0 / w.astype(np.float32).astype(np.float64)
    c0 = (mass[:, None] * x0).sum(0) / mass.sum()
    axis = (0, 0, 1) if shape[2] == 1 else (1, 2, 3)
    R = rot(axis, angle)
    x1 = (x0 - c0) @ R.T + c0 + np.array([1.5, -2.0, 0.75])
    set_positions(sol, x1)
    sol.solve_shapes()
    got = sol.download(psb.ARR_POS)[:, :3].astype(np.float64)
    assert np.abs(got - x1).max() < 2e-4, np.abs(got - x1).max()      # a rigid motion satisfies the constraint
    if n > 2:
        Rq = ext.quat_to_mat(sol.rigid_body_rotation(body))
        assert np.abs(Rq - R).max() < 2e-4
    # projecting again changes nothing (idempotent), now warm-started
    sol.solve_shapes()
    assert np.abs(sol.download(psb.ARR_POS)[:, :3] - got).max() < 2e-5
    sol.close()


@pytest.mark.parametrize("stiffness", [1.0, 0.5])
def test_shape_matching_of_deformed_bodies_matches_svd_polar_decomposition(stiffness):
    rng = np.random.default_rng(5)
    xa, xb = cube(4, 4, 4), cube(3, 2, 5, origin=(10, 30, 3))
    x0 = np.concatenate([xa, xb])
    n, na = x0.shape[0], xa.shape[0]
    w = rng.uniform(0.5, 2.0, n)
    phase = np.concatenate([np.full(na, psb.RIGID), np.full(n - na, psb.RIGID + 1)])
    sol = make_solver(x0, inv_mass=w, phase=phase)
    sol.add_rigid_body(np.arange(na), stiffness)
    sol.add_rigid_body(np.arange(na, n), stiffness)
    mass = 1.0 / w.astype(np.float32).astype(np.float64)
    x1 = x0.copy()
    expect = x0.copy()
    for sl, ang in ((slice(0, na), 0.8), (slice(na, n), 2.0)):
        c0 = (mass[sl, None] * x0[sl]).sum(0) / mass[sl].sum()
        rest = x0[sl] - c0
        x1[sl] = rest @ rot((3, 1, 2), ang).T + c0 + rng.normal(0, 0.08, rest.shape)
        expect[sl], R, _ = ext.shape_match(x1[sl].astype(np.float32).astype(np.float64), rest, mass[sl], stiffness)
        assert abs(np.linalg.det(R) - 1) < 1e-9
    set_positions(sol, x1)
    sol.solve_shapes()
    got = sol.download(psb.ARR_POS)[:, :3].astype(np.float64)
    assert np.abs(got - expect).max() < 2e-4, np.abs(got - expect).max()
    # mass-weighted centre of each body is preserved by the projection
    for sl in (slice(0, na), slice(na, n)):
        c1 = (mass[sl, None] * x1[sl].astype(np.float32)).sum(0) / mass[sl].sum()
        c2 = (mass[sl, None] * got[sl]).sum(0) / mass[sl].sum()
        assert np.abs(c1 - c2).max() < 1e-4
    sol.close()


def test_planar_shape_matching_agrees_with_the_reference_cpu_angle_estimator():
    """on a rigid planar motion the 3-D polar decomposition and the reference CPU solver's mass-weighted mean angle
    (Body::updateCOM, cpu/src/solver/particle.cpp:15-57, restated in oracle/cpu2d_full_oracle.py) find the same rotation"""
    import cpu2d_full_oracle as full
    x0 = cube(3, 2, 1, spacing=0.5, origin=(0, 10, 0))
    n = x0.shape[0]
    sol = make_solver(x0, phase=np.full(n, psb.RIGID))
    body = sol.add_rigid_body(np.arange(n))
    c0 = x0.mean(0)
    ang = 0.6
    x1 = (x0 - c0) @ rot((0, 0, 1), ang).T + c0
    set_positions(sol, x1)
    sol.solve_shapes()
    q = sol.rigid_body_rotation(body)
    gpu_angle = 2 * np.arctan2(q[2], q[3])
    scene = {"particles": [[x0[i, 0], x0[i, 1], 0, 0, 1.0, 0, 0, 0, 0] for i in range(n)], "xbounds": [-20, 20], "ybounds": [0, 100], "gravity": [0, -9.8],
             "bodies": [{"particles": list(range(n)), "rs": (x0 - c0)[:, :2].tolist(), "sdf": [[0, 1, .25]] * n, "imass": 1.0 / n, "center": c0[:2].tolist(),
                         "angle": 0.0, "stiffness": 1.0}], "standard": [], "rand_calls": 0}
    o = full.Cpu2dFullOracle(scene)
    o.ep = [[x1[i, 0], x1[i, 1]] for i in range(n)]
    o.update_com(o.bodies[0])
    assert abs(o.bodies[0]["angle"] - ang) < 1e-12
    assert abs(gpu_angle - o.bodies[0]["angle"]) < 1e-4
    sol.close()


def test_rigid_bodies_stay_rigid_through_whole_steps():
    xa, xb = cube(3, 3, 3, origin=(0, 3, 0)), cube(3, 3, 3, origin=(0.6, 6, 0.4))
    x0 = np.concatenate([xa, xb])
    n, na = x0.shape[0], xa.shape[0]
    phase = np.concatenate([np.full(na, psb.RIGID), np.full(n - na, psb.RIGID + 1)])
    sol = make_solver(x0, phase=phase)
    sol.add_rigid_body(np.arange(na))
    sol.add_rigid_body(np.arange(na, n))
    d0 = [np.linalg.norm(x0[sl][:, None] - x0[sl][None], axis=-1) for sl in (slice(0, na), slice(na, n))]
    for _ in range(120):
        sol.step(1 / 60)
    x = sol.download(psb.ARR_POS)[:, :3].astype(np.float64)
    assert np.isfinite(x).all()
    for sl, d in zip((slice(0, na), slice(na, n)), d0):
        d1 = np.linalg.norm(x[sl][:, None] - x[sl][None], axis=-1)
        assert np.abs(d1 - d).max() < 0.05, np.abs(d1 - d).max()   # contacts push single particles; the shape pulls them back every iteration
    assert x[:, 1].min() > -0.01 and x[:na, 1].mean() < 3.5        # fell onto the floor (y = 0) and stayed above it
    assert sol.launches_per_step > 0
    sol.close()


def fluid_block(side=12, seed=7):
    rng = np.random.default_rng(seed)
    x = cube(side, side, side, spacing=0.625, origin=(-3, 6, -3)) + rng.uniform(-0.05, 0.05, (side ** 3, 3))
    v = rng.normal(0, 1.0, x.shape)
    return x, v


@pytest.mark.parametrize("c_xsph,eps", [(0.01, 0.0), (0.0, 0.5), (0.05, 0.3)])
@pytest.mark.parametrize("lists", [256, 0, 40])
def test_viscosity_pass_matches_float64_restatement(c_xsph, eps, lists):
    x, v = fluid_block()
    n = x.shape[0]
    sol = make_solver(x, vel=v, lists=lists)
    sol.set_viscosity(c_xsph, eps)
    sol.build_grid()
    sol.find_neighbors()
    dt = 1 / 60
    sol.apply_viscosity(dt)
    got = sol.download(psb.ARR_VEL)[:, :3].astype(np.float64)
    x32, v32 = x.astype(np.float32).astype(np.float64), v.astype(np.float32).astype(np.float64)
    expect, _ = ext.viscosity(x32, v32, np.ones(n, bool), c_xsph, eps, dt)
    scale = np.abs(expect - v32).max()
    assert scale > 1e-4
    assert np.abs(got - expect).max() < 2e-4 * max(1.0, scale / 1e-2) , (np.abs(got - expect).max(), scale)
    if eps == 0.0:  # XSPH exchanges momentum symmetrically (equal masses): the total is conserved
        assert np.abs((got - v32).sum(0)).max() < 1e-3
    sol.close()


def test_viscosity_paths_agree_and_default_is_off():
    x, v = fluid_block(10)
    outs = []
    for lists in (256, 0):
        sol = make_solver(x, vel=v, lists=lists)
        sol.set_viscosity(0.02, 0.4)
        for _ in range(3):
            sol.step(1 / 60)
        outs.append((sol.download(psb.ARR_POS), sol.download(psb.ARR_VEL)))
        sol.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])  # same neighbour order either way
    sol = make_solver(x, vel=v)
    base = make_solver(x, vel=v)
    base.set_viscosity(0.0, 0.0)
    for _ in range(3):
        sol.step(1 / 60)
        base.step(1 / 60)
    assert np.array_equal(sol.download(psb.ARR_VEL), base.download(psb.ARR_VEL))
    assert not np.array_equal(sol.download(psb.ARR_VEL), outs[0][1])  # and switched on it does change the velocities
    sol.close()
    base.close()


def gas_scene(flags):
    x = cube(8, 8, 8, spacing=0.625, origin=(-2.5, 20, -2.5))
    n = x.shape[0]
    p = psb.default_params()
    p.flags = flags
    sol = psb.Solver(p, max_particles=1024)
    pos4 = np.concatenate([x, np.ones((n, 1))], 1).astype(np.float32)
    sol.append(pos4, np.zeros((n, 4), np.float32), np.ones(n), np.full(n, 4.1), np.full(n, psb.GAS))
    return sol, x


def test_gas_is_ignored_without_the_flag_like_the_reference():
    """the reference's GPU kernels skip phase GAS: such particles only fall ballistically"""
    sol, x = gas_scene(0)
    for _ in range(10):
        sol.step(1 / 60)
    got = sol.download(psb.ARR_POS)[:, :3].astype(np.float64)
    dt, g = 1 / 60, -9.8
    fall = sum(g * dt * dt * k for k in range(1, 11))
    assert np.abs(got[:, [0, 2]] - x[:, [0, 2]]).max() < 1e-5 and np.abs((got[:, 1] - x[:, 1]) - fall).max() < 1e-3
    sol.close()


def test_gas_with_the_flag_is_a_buoyant_pbf_fluid():
    sol, x = gas_scene(psb.FLAG_GAS)
    sol.predict(1 / 60)
    y1 = sol.download(psb.ARR_POS)[:, 1].astype(np.float64)
    assert np.allclose(y1 - x[:, 1], -9.8 * -0.2 / 3600, atol=1e-5)          # gravity x ALPHA (-0.2): the gas rises
    sol.build_grid()
    assert np.all(sol.download(psb.ARR_SORTED_PHASE) == psb.FLUID)           # the neighbour kernels see it as fluid ...
    assert np.all(sol.download(psb.ARR_PHASE) == psb.GAS)                     # ... the particle keeps its phase
    sol.find_neighbors()
    nn = sol.download(psb.ARR_NUM_NEIGHBORS)
    assert nn.min() > 20 and nn.max() <= 146 and np.abs(sol.download(psb.ARR_LAMBDA)).max() > 0
    for _ in range(60):
        sol.step(1 / 60)
    got = sol.download(psb.ARR_POS)[:, :3].astype(np.float64)
    assert np.isfinite(got).all() and got[:, 1].mean() > x[:, 1].mean() + 0.5  # a second later the cloud has risen
    spread = np.linalg.norm(got - got.mean(0), axis=1).mean() / np.linalg.norm(x - x.mean(0), axis=1).mean()
    assert 0.5 < spread < 3.0                                                 # and is held together / apart by the density constraint
    sol.close()


@pytest.mark.parametrize("scene", ["7", "3", "8"])
def test_fluid_stats_match_the_oracle_estimator(scene):
    """ps_fluid_stats (mean / max density error, kinetic energy — the quantities of the 1000-step parity bar) against the
    oracle's estimator on the same state; positions and velocities must come out untouched"""
    import helpers as H
    ps = psb.ParticleSystem.scene(scene)
    sol = ps.solver
    for _ in range(5):
        ps.update(1 / 60)
    pos, vel = sol.download(psb.ARR_POS), sol.download(psb.ARR_VEL)
    mean, mx, ke = sol.fluid_stats()
    o = H.oracle_from_solver(sol)
    omean, omx, oke = o.fluid_stats()
    assert abs(mean - omean) <= 1e-4 * max(omean, 1e-3) and abs(mx - omx) <= 1e-4 * max(omx, 1e-3), (mean, omean, mx, omx)
    assert abs(ke - oke) <= 1e-6 * max(oke, 1.0)
    assert np.array_equal(sol.download(psb.ARR_POS), pos) and np.array_equal(sol.download(psb.ARR_VEL), vel)
    ps.update(1 / 60)   # and the run goes on (graph replay after the diagnostics pass)
    ps.close()
