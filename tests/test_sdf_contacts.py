"""SDF contacts between rigid bodies in the contact pass K5 (ps_set_rigid_body_sdf, include/psolver.h): the reference CPU
app's RigidContactConstraint (cpu/src/constraint/rigidcontactconstraint.cpp:13-96, 2-D) lifted to 3-D.

The reference's GPU solver has no rigid bodies at all, so PARITY IS UNPINNED in 3-D.  Checked here: (CPU) the C
oracle's K5 with the SDF rule against an independent float64 all-pairs restatement (oracle/extensions_oracle.py), 1e-5, and
that both branches of the rule (surface layers / interior) occur in the scene; (GPU) the CUDA path against the C oracle
within helpers.POS_ATOL with exact contact counts, whole steps, and a checkpoint that carries the SDF data."""
import os
import sys

import numpy as np
import pytest

import helpers as H
import oracle_py as orc
import particlesolver_b200 as psb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import extensions_oracle as ext  # noqa: E402

DT = 1.0 / 60.0
R = 0.25


def lattice(nx, ny, nz, origin):
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    return g * (2 * R) + np.asarray(origin, np.float64)


def rotation(axis, angle):
    a = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


ROT_B = rotation((0.3, 1.0, 0.2), 0.35)


def scene():
    """A: 4x4x4 box with SDF.  B: 4x4x4 box with SDF, turned by ROT_B and pushed more than two layers deep into A (interior rule).
    C: 3x3x3 body WITHOUT SDF resting in A's top (plain contacts).  Returns rest poses, posed positions, phases, body slices."""
    a = lattice(4, 4, 4, (10, 10, 10))
    b_rest = lattice(4, 4, 4, (10.95, 10.13, 10.21))
    cb = b_rest.mean(0)
    b = (b_rest - cb) @ ROT_B.T + cb
    c = lattice(3, 3, 3, (10.2, 11.9, 10.3))
    rest = np.concatenate([a, b_rest, c])
    posed = np.concatenate([a, b, c])
    phase = np.concatenate([np.full(64, psb.RIGID + 1), np.full(64, psb.RIGID + 2), np.full(27, psb.RIGID + 3)]).astype(np.int32)
    return rest, posed, phase, (slice(0, 64), slice(64, 128), slice(128, 155))


def world_sdf(n, slices, rot_b):
    s = np.zeros((n, 4), np.float32)
    s[:, 3] = -1.0
    box = psb.box_sdf(4, 4, 4, R)
    s[slices[0]] = box
    s[slices[1], :3] = (box[:, :3].astype(np.float64) @ np.asarray(rot_b, np.float64).T).astype(np.float32)
    s[slices[1], 3] = box[:, 3]
    return s


def test_box_sdf_follows_the_reference_builders():
    s = psb.box_sdf(4, 4, 4, R)
    corner, face, inner = s[0], s[1 * 16 + 1 * 4 + 0], s[1 * 16 + 1 * 4 + 1]
    assert np.allclose(corner[:3], -np.ones(3) / np.sqrt(3)) and np.isclose(corner[3], R * np.sqrt(3))   # simulation.cpp:667: corners radius * sqrt(k)
    assert np.allclose(face[:3], (0, 0, -1)) and np.isclose(face[3], R)                                    # :669: faces radius
    assert np.isclose(inner[3], 3 * R * np.sqrt(3)) and np.isclose(np.linalg.norm(inner[:3]), 1.0)         # second layer, three faces equally near
    assert np.allclose(np.linalg.norm(s[:, :3], axis=1), 1.0)


def test_oracle_contact_pass_with_sdf_matches_the_float64_restatement():
    rest, posed, phase, sl = scene()
    n = posed.shape[0]
    rng = np.random.default_rng(5)
    pos4 = np.concatenate([posed, np.ones((n, 1))], 1).astype(np.float32)
    w = rng.uniform(0.5, 2.0, n).astype(np.float32)
    o = orc.OracleSystem(orc.make_params(), pos4, np.zeros((n, 4), np.float32), w, phase, np.ones(n, np.float32))
    o.prev[:, :3] = (posed + rng.normal(0, 0.01, posed.shape)).astype(np.float32)   # some tangential motion for the friction terms
    sdf = world_sdf(n, sl, ROT_B)
    for use_sdf in (False, True):
        o.pos[:] = pos4
        o.sdf_world = sdf if use_sdf else None
        o.build_grid()
        o.collide()
        want, counts = ext.contact_pass(pos4, o.prev, w, phase, R, sdf_world=sdf if use_sdf else None)
        got_counts = np.zeros(n, np.int64)
        got_counts[o.index] = o.nn
        assert np.array_equal(got_counts, counts)
        assert np.abs(o.pos[:, :3] - want).max() < 1e-5
        if use_sdf:
            with_sdf = o.pos.copy()
        else:
            plain = o.pos.copy()
    changed = np.abs(with_sdf - plain).max(1) > 1e-6
    assert changed[sl[0]].any() and changed[sl[1]].any()           # the rule acts between the two SDF bodies ...
    a_only_c = [i for i in range(64) if changed[i]]
    assert len(a_only_c) < 64 and not changed[sl[2]].any()         # ... and nowhere else: C (no SDF) is treated as before
    # both branches of the rule occur: surface-layer pairs (overlap depth, mirrored normal) and interior pairs (SDF depth and normal)
    x = pos4[:, :3].astype(np.float64)
    kinds = set()
    for i in range(64):
        for j in range(64, 128):
            d = np.linalg.norm(x[i] - x[j])
            if d < 2.001 * R:
                kinds.add("interior" if min(sdf[i, 3], sdf[j, 3]) >= 2 * R + ext.EPS else "surface")
    assert kinds == {"interior", "surface"}


def gpu_scene():
    rest, posed, phase, sl = scene()
    n = rest.shape[0]
    p = psb.default_params()
    sol = psb.Solver(p, max_particles=1024)
    pos4 = np.concatenate([rest, np.ones((n, 1))], 1).astype(np.float32)
    w = np.random.default_rng(5).uniform(0.5, 2.0, n).astype(np.float32)
    sol.append(pos4, np.zeros((n, 4), np.float32), w, np.ones(n), phase)
    bodies = [sol.add_rigid_body(np.arange(s.start, s.stop)) for s in sl]
    box = psb.box_sdf(4, 4, 4, R)
    sol.set_rigid_body_sdf(bodies[0], box)
    sol.set_rigid_body_sdf(bodies[1], box)
    p4 = sol.download(psb.ARR_POS)
    p4[:, :3] = posed
    sol.upload(psb.ARR_POS, p4)
    return sol, bodies, sl, posed


@pytest.mark.gpu
def test_gpu_contact_pass_with_sdf_vs_oracle():
    sol, bodies, sl, posed = gpu_scene()
    n = posed.shape[0]
    sol.begin_step()
    sol.solve_shapes()                                  # finds body B's rotation (the members are already in a rigid pose: they do not move)
    assert np.abs(sol.download(psb.ARR_POS)[:, :3] - posed).max() < 2e-4
    rot_b = ext.quat_to_mat(sol.rigid_body_rotation(bodies[1]))
    assert np.abs(rot_b - ROT_B).max() < 1e-3
    prev = sol.download(psb.ARR_PREV)
    prev[:, :3] = posed + np.random.default_rng(6).normal(0, 0.01, posed.shape)
    sol.upload(psb.ARR_PREV, prev)
    o = H.oracle_from_solver(sol)
    o.prev[:] = sol.download(psb.ARR_PREV)
    o.sdf_world = world_sdf(n, sl, rot_b)
    sol.build_grid(); o.build_grid()
    H.assert_grid_equal(sol, o)
    before = o.pos.copy()
    sol.solve_contacts(); o.collide()
    assert np.array_equal(sol.download(psb.ARR_NUM_NEIGHBORS), o.nn)
    assert H.max_abs(sol.download(psb.ARR_POS), o.pos) <= H.POS_ATOL
    assert np.abs(o.pos - before).max() > 0.05          # a real correction, not a no-op
    o.sdf_world = None
    o.pos[:] = before
    o.collide()
    assert H.max_abs(sol.download(psb.ARR_POS), o.pos) > 10 * H.POS_ATOL   # and not the centre-to-centre rule
    sol.close()


@pytest.mark.gpu
def test_interpenetrating_boxes_with_sdf_come_apart_and_stay_rigid(tmp_path):
    sol, bodies, sl, posed = gpu_scene()
    for _ in range(40):
        sol.step(DT)
    path = os.path.join(tmp_path, "sdf.ckpt")
    sol.save(path)
    twin = psb.Solver.load(path)
    for _ in range(20):
        sol.step(DT); twin.step(DT)
    x = sol.download(psb.ARR_POS)[:, :3].astype(np.float64)
    assert np.array_equal(sol.download(psb.ARR_POS), twin.download(psb.ARR_POS))   # the checkpoint carries the SDF data
    twin.close()
    assert np.isfinite(x).all()
    a, b = x[sl[0]], x[sl[1]]
    gap = np.linalg.norm(a[:, None, :] - b[None, :, :], axis=2).min()
    assert gap > 0.8 * 2 * R                                                       # from > 2 layers of overlap to (soft) touching
    for s in sl[:2]:                                                                # shape matching keeps the boxes boxes
        d0 = np.linalg.norm(posed[s][:, None] - posed[s][None], axis=2)
        d1 = np.linalg.norm(x[s][:, None] - x[s][None], axis=2)
        assert np.abs(d1 - d0).max() < 0.05
    sol.close()


@pytest.mark.gpu
def test_sdf_argument_checks():
    sol, bodies, sl, posed = gpu_scene()
    bad = psb.box_sdf(4, 4, 4, R)
    bad[3, :3] = 0
    with pytest.raises(psb.PsError):
        sol.set_rigid_body_sdf(bodies[0], bad)          # a depth without a gradient
    with pytest.raises(psb.PsError):
        sol.set_rigid_body_sdf(99, bad)
    none = np.zeros((27, 4), np.float32); none[:, 3] = -1
    sol.set_rigid_body_sdf(bodies[2], none)             # "no data" is fine
    sol.step(DT)
    sol.close()


@pytest.mark.gpu
def test_rigid_box_scene_of_the_host_class_stacks_and_stays_rigid():
    """extension scene "r" (psb200::build_rigid_scene, addRigidBox): shape-matched boxes with SDF data through the C++ host class"""
    ps = psb.ParticleSystem.scene("r")
    sol = ps.solver
    assert sol.num_rigid_bodies == 4 and ps.getNumParticles() == 4 * 125 + 3 * 5 * 3
    x0 = ps.getPositions()[:, :3].astype(np.float64)
    for _ in range(360):
        ps.update(DT)
    x = ps.getPositions()[:, :3].astype(np.float64)
    assert np.isfinite(x).all()
    boxes = [slice(125 * k, 125 * (k + 1)) for k in range(4)]
    for s in boxes:                                   # every box is still a box
        d0 = np.linalg.norm(x0[s][:, None] - x0[s][None], axis=2)
        d1 = np.linalg.norm(x[s][:, None] - x[s][None], axis=2)
        assert np.abs(d1 - d0).max() < 0.05
    for a in range(4):                                # and no two boxes sit inside each other
        for b in range(a + 1, 4):
            gap = np.linalg.norm(x[boxes[a]][:, None] - x[boxes[b]][None], axis=2).min()
            assert gap > 0.8 * 2 * R, (a, b, gap)
    y = [x[s][:, 1].mean() for s in boxes[:3]]
    assert abs(y[0] - 1.25) < 0.1 and abs(y[1] - 3.75) < 0.15 and abs(y[2] - 6.25) < 0.2   # the tower of 2.5-unit boxes stands, at rest
    assert np.abs(ps.getVelocities()[:500, :3]).max() < 0.05
    assert x[:, 1].min() > 0.2                        # nothing went through the floor
    ps.close()
