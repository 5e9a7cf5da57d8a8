"""Checkpoints (ps_save / ps_load, ps2d_save / ps2d_load) and the headless runner (psolver_cli): SURVEY §8f row 1.
A run continued from a checkpoint must be BIT-IDENTICAL to the uninterrupted run — particles, constraints, rigid bodies,
emitters and both random streams (cuRAND wall jitter in 3-D, glibc rand() in 2-D) included."""
import json
import os
import subprocess

import numpy as np
import pytest

import particlesolver_b200 as psb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "particlesolver_b200", "psolver_cli")


@pytest.mark.parametrize("scene", ["8", "3", "7"])
def test_3d_checkpoint_continues_bit_identically(scene, tmp_path):
    ps = psb.ParticleSystem.scene(scene)
    sol = ps.solver
    if scene == "7":
        sol.set_viscosity(0.01, 0.2)
    if scene == "8":   # a rigid body over the first solid block of the combo scene
        ph = sol.download(psb.ARR_PHASE)
        solid = np.nonzero(ph == psb.SOLID)[0][:27]
        sol.add_rigid_body(solid, 0.8)
    for _ in range(4):
        ps.update(1 / 60)
    path = str(tmp_path / "c.psb")
    sol.save(path)
    for _ in range(4):
        ps.update(1 / 60)
    a_pos, a_vel = sol.download(psb.ARR_POS), sol.download(psb.ARR_VEL)
    re = psb.Solver.load(path)
    assert re.n == sol.n and re.num_rigid_bodies == sol.num_rigid_bodies
    i0, r0 = sol.distance_constraints()
    i1, r1 = re.distance_constraints()
    assert np.array_equal(i0, i1) and np.array_equal(r0, r1)
    for _ in range(4):
        re.step(1 / 60)
    assert np.array_equal(re.download(psb.ARR_POS), a_pos)
    assert np.array_equal(re.download(psb.ARR_VEL), a_vel)
    re.close()
    ps.close()


def test_3d_checkpoint_carries_the_stale_lambdas_of_solids_next_to_fluid(tmp_path):
    """K6 leaves lambda of non-fluid sorted slots untouched and K7 reads lambda_j of every neighbour (the reference's behaviour,
    integration_kernel.cuh:596-642): with fluid resting against solids those values are state that crosses steps, so the
    checkpoint (format v3) carries the lambda array.  Fluid block between two solid blocks, in contact from the first step."""
    def block(x0, x1, phase, rho0):
        g = np.mgrid[x0:x1:0.5, 0.25:4.0:0.5, -2.0:2.0:0.5].reshape(3, -1).T.astype(np.float32)
        pos = np.ones((g.shape[0], 4), np.float32)
        pos[:, :3] = g
        k = g.shape[0]
        return pos, np.zeros((k, 4), np.float32), np.ones(k, np.float32), np.full(k, rho0, np.float32), np.full(k, phase, np.int32)
    parts = [block(-4.0, -2.0, psb.SOLID, 1.0), block(-2.0, 2.0, psb.FLUID, 1.5), block(2.0, 4.0, psb.SOLID, 1.0)]
    n = sum(p[0].shape[0] for p in parts)
    prm = psb.default_params()
    sol = psb.Solver(prm, max_particles=n)
    for p in parts:
        sol.append(*p)
    for _ in range(6):
        sol.step(1 / 60)
    # the premise: some solid slot holds a non-zero (stale) lambda, and it neighbours fluid
    lam, sph = sol.download(psb.ARR_LAMBDA), sol.download(psb.ARR_SORTED_PHASE)
    assert np.any(lam[sph != psb.FLUID] != 0.0)
    path = str(tmp_path / "mixed.psb")
    sol.save(path)
    for _ in range(6):
        sol.step(1 / 60)
    re = psb.Solver.load(path)
    for _ in range(6):
        re.step(1 / 60)
    assert np.array_equal(re.download(psb.ARR_POS), sol.download(psb.ARR_POS))
    assert np.array_equal(re.download(psb.ARR_VEL), sol.download(psb.ARR_VEL))
    re.close()
    sol.close()


@pytest.mark.parametrize("key", ["w", "8", "0", "2", "v"])
def test_2d_checkpoint_continues_bit_identically(key, tmp_path):
    sim = psb.Simulation2D.scene(key)
    for _ in range(12):
        sim.tick(.01)
    path = str(tmp_path / "c.ps2")
    sim.save(path)
    for _ in range(10):
        sim.tick(.01)
    re = psb.Simulation2D.load(path)
    assert re.getNumBodies() == sim.getNumBodies()
    for _ in range(10):
        re.tick(.01)
    assert re.getNumParticles() == sim.getNumParticles()
    assert np.array_equal(re.positions(), sim.positions()) and np.array_equal(re.velocities(), sim.velocities())
    assert re.rand_calls == sim.rand_calls
    re.close()
    sim.close()


def test_cli_runs_both_apps_and_round_trips_checkpoints(tmp_path):
    g6 = np.load(os.path.join(ROOT, "tests", "golden", "ref_cpu_scene6.npz"))
    out = subprocess.run([CLI, "--app", "cpu", "--scene", "6", "--ticks", "100", "--json"], capture_output=True, text=True, check=True).stdout
    r = json.loads(out)
    assert r["particles"] == 432 and r["scene_name"] == "FLUID_TEST"
    ke100 = float(g6["ke_every_100"][0])
    assert abs(r["kinetic_energy"] - ke100) <= 1e-4 * ke100    # the reference CPU solver's kinetic energy after 100 ticks
    ck = str(tmp_path / "g.psb")
    a = json.loads(subprocess.run([CLI, "--app", "gpu", "--scene", "7", "--steps", "6", "--json"], capture_output=True, text=True, check=True).stdout)
    subprocess.run([CLI, "--app", "gpu", "--scene", "7", "--steps", "3", "--save", ck], capture_output=True, text=True, check=True)
    b = json.loads(subprocess.run([CLI, "--app", "gpu", "--load", ck, "--steps", "3", "--json", "--dump-every", "3", "--out", str(tmp_path / "d")],
                                  capture_output=True, text=True, check=True).stdout)
    assert a["particles"] == b["particles"] == 5324
    assert a["position_checksum"] == b["position_checksum"] and a["kinetic_energy"] == b["kinetic_energy"]
    raw = open(tmp_path / "d" / "step000003.bin", "rb").read()
    assert raw[:7] == b"PSDUMP1" and len(raw) == 24 + 16 * 5324
    bad = subprocess.run([CLI, "--app", "cpu", "--scene", "x"], capture_output=True, text=True)
    assert bad.returncode == 1 and "unknown scene" in bad.stderr


def test_cli_switches_the_optional_contact_rules_on():
    """--self-collision sets PS_FLAG_SELF_COLLISION: scene 6 (solids falling on a held cloth) ends elsewhere than with the
    reference's rule; scene 7 (fluid only, no distance constraints) is untouched by it"""
    run = lambda *a: json.loads(subprocess.run([CLI, "--app", "gpu", "--steps", "40", "--json", *a], capture_output=True, text=True, check=True).stdout)
    assert run("--scene", "7")["position_checksum"] == run("--scene", "7", "--self-collision")["position_checksum"]
    a, b = run("--scene", "2"), run("--scene", "2", "--self-collision")
    assert a["particles"] == b["particles"] and np.isfinite(b["kinetic_energy"])


def test_2d_tick_graph_replay_equals_eager_issue():
    """the 2-D tick is replayed as a CUDA graph once its key has stood still for a few ticks; PS_NO_GRAPH=1 issues every launch
    eagerly — the two must be the same computation (scenes with walls + jitter draws, rigid bodies, an emitter that changes n)"""
    seq = dict(os.environ, PS2D_FUSED_MAX_N="0")   # the launch sequence (these scenes would otherwise run their tick as one fused kernel)
    for key in ("6", "w", "s"):
        a = subprocess.run([CLI, "--app", "cpu", "--scene", key, "--ticks", "120", "--json"], capture_output=True, text=True, check=True, env=seq).stdout
        b = subprocess.run([CLI, "--app", "cpu", "--scene", key, "--ticks", "120", "--json"], capture_output=True, text=True, check=True,
                           env=dict(seq, PS_NO_GRAPH="1")).stdout
        a, b = json.loads(a), json.loads(b)
        assert a["launches_per_tick"] > 1 and b["launches_per_tick"] > 1
        assert a["kinetic_energy"] == b["kinetic_energy"] and a["rand_calls"] == b["rand_calls"] and a["particles"] == b["particles"], key


def test_cli_emitter_and_shooting_grow_the_system_until_it_is_full():
    """The GPU app's run-time emission paths, headless: the fluid-emitter toggle (particleapp.cpp:74-79: addFluid of a
    3 x 1 x 3 block every 0.1 s) and the mouse shot (:91-96: setParticleToAdd).  The particle count changes between steps, so the step's
    CUDA graph is re-captured again and again; a batch that would reach maxParticles is dropped ('>=', particlesystem.cpp:335) and the
    run carries on."""
    def run(*extra):
        r = subprocess.run([CLI, "--app", "gpu", "--scene", "1", "--json", *extra], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        return json.loads(r.stdout.strip().splitlines()[-1])
    base = run("--steps", "60")
    n0 = base["particles"]
    assert n0 == base["particles_at_start"] and 30 <= n0 <= 34     # the rope of makeInitScene
    grown = run("--steps", "60", "--emit", "--shoot", "7")
    # 60 steps of 1/60 s: a block at step 1 and then every 0.1 s (9 fluid particles each: ceil(2) / 0.625 = 3 per axis in x and z, ceil(1) / 0.625 = 1 in y), a shot every 7 steps
    assert grown["particles"] > n0 + 9 * 9 and grown["emitter_hit_capacity"] is False
    assert np.isfinite(grown["kinetic_energy"]) and np.isfinite(grown["position_checksum"])
    full = run("--steps", "120", "--emit", "--max-particles", "80")
    assert full["emitter_hit_capacity"] is True and n0 < full["particles"] < 80      # the batch that would reach 80 was dropped
    assert np.isfinite(full["kinetic_energy"])


def test_host_class_emission_matches_between_graph_and_eager_issue():
    """growing n mid-run through the host class: the graph-replayed run (re-captured whenever n changes) equals the eagerly issued one"""
    import os as _os
    outs = []
    for no_graph in (False, True):
        env = dict(_os.environ)
        if no_graph:
            env["PS_NO_GRAPH"] = "1"
        r = subprocess.run([CLI, "--app", "gpu", "--scene", "7", "--steps", "40", "--emit", "--shoot", "5", "--json"], capture_output=True, text=True,
                           timeout=300, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0]["particles"] == outs[1]["particles"] > outs[0]["particles_at_start"]
    assert outs[0]["position_checksum"] == outs[1]["position_checksum"] and outs[0]["kinetic_energy"] == outs[1]["kinetic_energy"]
