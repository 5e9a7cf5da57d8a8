"""The 2-D tick as one kernel (k2d_tick_fused, scenes of a few dozen particles by default) against the launch sequence it replaces:
the same device functions in the same order, so every scene must evolve bit for bit the same way — walls with jitter draws, rigid
bodies with SDF contacts, ropes (level-scheduled distance runs), fluids, gas with an emitter that changes n, the fluid emitter that
reads the kept lambdas.  PS2D_FUSED_MAX_N moves the size threshold (0: never fused; the kernel's shared memory allows 2048)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "particlesolver_b200", "psolver_cli")


def run(key, max_n, ticks=90, extra=()):
    env = dict(os.environ, PS2D_FUSED_MAX_N=str(max_n))
    r = subprocess.run([CLI, "--app", "cpu", "--scene", key, "--ticks", str(ticks), "--json", *extra], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("key", ["8", "6", "1", "2", "7", "0", "w", "v", "s"])
def test_fused_tick_equals_the_launch_sequence(key):
    a, b = run(key, 0), run(key, 2048)
    assert a["launches_per_tick"] > 1
    if b["particles"] <= 2048:
        assert b["launches_per_tick"] == 1
    assert a["particles"] == b["particles"] and a["rand_calls"] == b["rand_calls"], (a, b)
    assert a["kinetic_energy"] == b["kinetic_energy"], (a, b)


def test_fused_tick_with_the_stabilization_pass():
    a, b = run("2", 0, extra=("--stabilization", "2")), run("2", 2048, extra=("--stabilization", "2"))
    assert a["kinetic_energy"] == b["kinetic_energy"] and a["rand_calls"] == b["rand_calls"]


def test_small_scenes_are_fused_by_default():
    env = {k: v for k, v in os.environ.items() if k != "PS2D_FUSED_MAX_N"}
    r = subprocess.run([CLI, "--app", "cpu", "--scene", "8", "--ticks", "30", "--json"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["particles"] <= 800 and out["launches_per_tick"] == 1
