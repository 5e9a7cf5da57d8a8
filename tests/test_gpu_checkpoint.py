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
    for key in ("6", "w", "s"):
        a = subprocess.run([CLI, "--app", "cpu", "--scene", key, "--ticks", "120", "--json"], capture_output=True, text=True, check=True).stdout
        b = subprocess.run([CLI, "--app", "cpu", "--scene", key, "--ticks", "120", "--json"], capture_output=True, text=True, check=True,
                           env=dict(os.environ, PS_NO_GRAPH="1")).stdout
        a, b = json.loads(a), json.loads(b)
        assert a["kinetic_energy"] == b["kinetic_energy"] and a["rand_calls"] == b["rand_calls"] and a["particles"] == b["particles"], key
