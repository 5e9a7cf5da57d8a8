"""Pins oracle/cpu2d_full_oracle.py (every constraint group of the reference's 2-D CPU solver) to
tests/golden/ref_cpu_scenes.npz: states written by the reference's own unmodified CPU solver for all 15 key-bound scenes.
The restatement follows the reference expression by expression and uses the same libm, so the bar is 1e-12 (observed: 0)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cpu2d_full_oracle as full  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "ref_cpu_scenes.npz"))
SCENES = sorted(k[:-6] for k in G.files if k.endswith("_scene"))
# how many of the kept ticks the (pure-Python, O(N^2)) oracle replays per scene in the CPU suite
BUDGET = {"wrecking_ball": 2, "wall": 3, "fluid_solid": 4, "balloon": 4}


def test_all_scenes_present():
    assert len(SCENES) == 16


@pytest.mark.parametrize("name", SCENES)
def test_full_oracle_reproduces_reference_cpu_solver(name):
    scene = json.loads(str(G[f"{name}_scene"]))
    o = full.Cpu2dFullOracle(scene)
    t = int(G[f"{name}_t0"])
    ticks = [int(x) for x in G[f"{name}_ticks"]][: BUDGET.get(name, 5)]
    saw_contacts = 0
    for target in ticks:
        while t < target:
            o.tick(.01)
            t += 1
            saw_contacts += o.num_contacts
        p, v = G[f"{name}_p{t}"], G[f"{name}_v{t}"]
        assert o.n == p.shape[0], f"{name} tick {t}: particle count {o.n} vs {p.shape[0]}"
        dp, dv = np.abs(o.positions() - p).max(), np.abs(o.velocities() - v).max()
        assert dp <= 1e-12 and dv <= 1e-10, f"{name} tick {t}: |dp| {dp:.3e} |dv| {dv:.3e}"
        assert o.rng.calls == int(G[f"{name}_rand{t}"]), f"{name} tick {t}: rand() draws differ"
        assert abs(o.kinetic_energy() - float(G[f"{name}_ke{t}"])) <= 1e-12 * max(1., abs(float(G[f"{name}_ke{t}"])))
    if name in ("granular", "stacks", "wall", "friction", "sdf", "wrecking_ball", "fluid_solid", "balloon", "rope"):
        assert saw_contacts > 0, f"{name}: the replayed ticks exercise no contact constraint"


# ---- the stabilization pass: the reference compiled with its option USE_STABILIZATION (oracle/_ref/ref_cpu_stab) ----
GS = np.load(os.path.join(ROOT, "tests", "golden", "ref_cpu_scenes_stab.npz"))
STAB_SCENES = sorted(k[:-6] for k in GS.files if k.endswith("_scene"))


def test_stabilization_fixture_differs_from_the_default_build():
    assert STAB_SCENES == ["fluid_solid", "friction", "sdf", "stacks", "wall", "wrecking_ball"]
    for name in STAB_SCENES:
        t = int(GS[f"{name}_ticks"][-1])
        assert np.abs(GS[f"{name}_p{t}"] - G[f"{name}_p{t}"]).max() > 1e-3      # a different trajectory ...
    assert int(GS["fluid_solid_rand20"]) > int(G["fluid_solid_rand20"])           # ... and more jitter draws (stabile wall constraints draw too)


@pytest.mark.parametrize("name", STAB_SCENES)
def test_full_oracle_reproduces_reference_cpu_solver_with_stabilization(name):
    scene = json.loads(str(GS[f"{name}_scene"]))
    o = full.Cpu2dFullOracle(scene, stabilization_iterations=2)
    t = int(GS[f"{name}_t0"])
    ticks = [int(x) for x in GS[f"{name}_ticks"]][: BUDGET.get(name, 5)]
    for target in ticks:
        while t < target:
            o.tick(.01)
            t += 1
        p, v = GS[f"{name}_p{t}"], GS[f"{name}_v{t}"]
        dp, dv = np.abs(o.positions() - p).max(), np.abs(o.velocities() - v).max()
        assert dp <= 1e-12 and dv <= 1e-10, f"{name} tick {t}: |dp| {dp:.3e} |dv| {dv:.3e}"
        assert o.rng.calls == int(GS[f"{name}_rand{t}"]), f"{name} tick {t}: rand() draws differ"
