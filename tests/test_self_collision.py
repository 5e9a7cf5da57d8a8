"""Opt-in self-collision of constrained bodies in the contact pass K5 (PS_FLAG_SELF_COLLISION, include/psolver.h).

NOT in the reference: there particles of one phase > SOLID always skip each other (integration_kernel.cuh:336-337), so the
cloth of GPU scene 2 — which BASELINE.json describes as "with particle self-collision" — never touches itself.
PARITY UNPINNED.  What is checked: (CPU) the oracle's rule against a plain numpy statement of it, and that without the
adjacency the oracle is the reference's rule; (GPU) the CUDA path against the oracle stage by stage within
helpers.POS_ATOL with exact contact counts, both with and without the flag, and the behaviour over whole steps."""
import types

import numpy as np
import pytest

import helpers as H
import oracle_py as orc
import particlesolver_b200 as psb

DT = 1.0 / 60.0
SP = 0.5  # particle diameter = cell size


def folded_scene(nx=16, nz=6, gap=0.3, y0=6.0):
    """One cloth (phase RIGID+1, structural distance constraints) folded back over itself `gap` apart, a compressed block of
    a shape-matched body (phase RIGID+2, NO distance constraints) and a loose solid particle resting in the upper layer."""
    pos, phase, pairs, rest = [], [], [], []
    half = nx // 2
    for i in range(nx):
        for k in range(nz):
            if i < half:
                pos.append((10 + i * SP, y0, 10 + k * SP))
            else:
                pos.append((10 + (nx - 1 - i) * SP + 0.1, y0 + gap, 10 + k * SP))
            phase.append(psb.RIGID + 1)
    idx = lambda i, k: i * nz + k
    for i in range(nx):
        for k in range(nz):
            if i + 1 < nx: pairs.append((idx(i, k), idx(i + 1, k))); rest.append(SP)
            if k + 1 < nz: pairs.append((idx(i, k), idx(i, k + 1))); rest.append(SP)
    n_cloth = len(pos)
    for a in range(3):
        for b in range(3):
            for c in range(3):
                pos.append((20 + a * 0.4, y0 + b * 0.4, 20 + c * 0.4))  # members overlap: 0.4 < 2.001 r
                phase.append(psb.RIGID + 2)
    pos.append((10 + 2 * SP + 0.05, y0 + gap + 0.35, 10 + 2 * SP))
    phase.append(psb.SOLID)
    pos = np.asarray(pos, np.float64)
    return pos, np.asarray(phase, np.int32), np.asarray(pairs, np.uint32), np.asarray(rest, np.float32), n_cloth


def make_oracle(self_collision):
    pos, phase, pairs, rest, n_cloth = folded_scene()
    n = pos.shape[0]
    pos4 = np.concatenate([pos, np.ones((n, 1))], 1).astype(np.float32)
    o = orc.OracleSystem(orc.make_params(), pos4, np.zeros((n, 4), np.float32), np.ones(n, np.float32), phase, np.ones(n, np.float32),
                         pairs, rest)
    o.self_collision = self_collision
    o.prev[:] = o.pos
    return o, pairs, n_cloth


def expected_contact_counts(pos, phase, pairs, self_collision, radius=0.25):
    """the rule in plain numpy (float64 all-pairs)"""
    n = pos.shape[0]
    d = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=2)
    touch = (d < radius * 2.001) & ~np.eye(n, dtype=bool)
    linked = np.zeros((n, n), bool)
    linked[pairs[:, 0], pairs[:, 1]] = linked[pairs[:, 1], pairs[:, 0]] = True
    constrained = linked.any(1)
    same = (phase[:, None] == phase[None, :]) & (phase[:, None] > psb.SOLID)
    skip = same & ~(self_collision & constrained[:, None] & constrained[None, :] & ~linked)
    return (touch & ~skip).sum(1), touch & ~skip


@pytest.mark.parametrize("self_collision", [False, True])
def test_oracle_rule_matches_the_plain_statement(self_collision):
    o, pairs, n_cloth = make_oracle(self_collision)
    before = o.pos.copy()
    o.build_grid()
    o.collide()
    x = before[:, :3].astype(np.float64)
    want, contact = expected_contact_counts(x, o.phase, pairs.astype(np.int64), self_collision)
    got = np.zeros(o.n, np.int64)
    got[o.index] = o.nn                       # counts live by sorted slot
    assert np.array_equal(got, want)
    moved = np.abs(o.pos[:, :3] - before[:, :3]).max(1) > 0
    assert np.array_equal(moved, want > 0)    # exactly the particles with a contact move
    body = slice(n_cloth, n_cloth + 27)
    assert not moved[body].any()              # members of an unconstrained (shape-matched) body never collide with each other
    if self_collision:
        rows = want[:n_cloth].reshape(16, 6)  # every cloth particle touches the other layer (the first row past the fold only meets
        assert rows[8].max() == 0 and np.delete(rows, 8, 0).min() >= 1  # its constrained neighbour, which is skipped) ...
        i, j = np.nonzero(np.triu(contact[:n_cloth, :n_cloth]))
        d0 = np.linalg.norm(x[i] - x[j], axis=1)
        d1 = np.linalg.norm(o.pos[i, :3].astype(np.float64) - o.pos[j, :3], axis=1)
        assert (d1 > d0).all()                # ... and the layers are pushed apart
    else:
        assert 1 <= want[:n_cloth].sum() <= 2  # the reference: only the loose solid particle touches the cloth


def gpu_solver(flags):
    pos, phase, pairs, rest, n_cloth = folded_scene()
    n = pos.shape[0]
    p = psb.default_params()
    p.flags = flags
    sol = psb.Solver(p, max_particles=1024)
    pos4 = np.concatenate([pos, np.ones((n, 1))], 1).astype(np.float32)
    sol.append(pos4, np.zeros((n, 4), np.float32), np.ones(n), np.ones(n), phase)
    sol.add_distance_constraints(pairs, rest)
    return sol, pairs, n_cloth


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, psb.FLAG_SELF_COLLISION], ids=["reference_rule", "self_collision"])
def test_contact_stage_and_whole_steps_vs_oracle(flags):
    from test_gpu_parity import staged_compare
    sol, pairs, n_cloth = gpu_solver(flags)
    o = H.oracle_from_solver(sol)
    assert o.self_collision == bool(flags)
    sol.begin_step(); sol.build_grid(); o.build_grid()
    H.assert_grid_equal(sol, o)
    sol.solve_contacts(); o.collide()
    ph = o.sphase >= psb.CLOTH
    assert np.array_equal(sol.download(psb.ARR_NUM_NEIGHBORS)[ph], o.nn[ph])
    assert H.max_abs(sol.download(psb.ARR_POS), o.pos) <= H.POS_ATOL
    sol.close()
    sol, _, _ = gpu_solver(flags)
    worst = staged_compare(types.SimpleNamespace(solver=sol), steps=3)   # every stage of three steps, contact counts implied by positions
    assert worst["contacts"] <= H.POS_ATOL
    sol.close()


@pytest.mark.gpu
def test_a_folded_cloth_keeps_its_layers_apart_only_with_the_flag():
    layer_gap = {}
    for flags in (0, psb.FLAG_SELF_COLLISION):
        sol, pairs, n_cloth = gpu_solver(flags)
        for _ in range(90):   # falls 6 units onto the floor and settles
            sol.step(DT)
        x = sol.download(psb.ARR_POS)[:n_cloth, :3].astype(np.float64)
        assert np.isfinite(x).all()
        nz, nx = 6, 16
        y = x[:, 1].reshape(nx, nz)
        layer_gap[flags] = float(np.median(y[nx // 2 + 2:, :]) - np.median(y[:nx // 2 - 2, :]))
        sol.close()
    assert layer_gap[0] < 0.15                      # the reference's rule: the upper layer sinks into the lower one
    assert layer_gap[psb.FLAG_SELF_COLLISION] > 0.3  # with self-collision it rests on it (contact distance 0.5, soft Jacobi contacts)


@pytest.mark.gpu
def test_flag_changes_nothing_without_distance_constraints():
    """fluid / unconstrained scenes carry no adjacency: the flag must leave them bit-identical"""
    outs = []
    for flags in (0, psb.FLAG_SELF_COLLISION):
        ps = psb.ParticleSystem.scene("5")
        p = ps.solver.params
        p.flags = flags
        ps.solver.set_params(p)
        for _ in range(3):
            ps.update(DT)
        outs.append(ps.solver.download(psb.ARR_POS).copy())
        ps.close()
    assert np.array_equal(outs[0], outs[1])
