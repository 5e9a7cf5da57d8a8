"""1000-step parity bar of north_star: density-error and energy statistics against the reference's own GPU solver.
Per-particle comparison is meaningless after ~10 steps (chaotic divergence), so the bar is statistical:
every 100 steps, mean |rho/rho0 - 1| over the fluid and the kinetic energy must stay inside a band around the series of
the reference's UNMODIFIED CUDA code (tests/golden/ref_gpu_stats.json, generated on a B200 by
tests/golden/make_stats_golden.py).  Bands: mean density error within 35 % + 0.005 absolute; kinetic energy within a
factor [0.6, 1.6] (the series decays over two orders of magnitude; the oracle itself sits within [0.78, 1.05] of it)."""
import json
import os

import numpy as np
import pytest

import golden_io as G
import oracle_py as orc

REF = json.load(open(os.path.join(G.GOLDEN_DIR, "ref_gpu_stats.json")))
EVERY, STEPS = REF["every"], REF["steps"]


def check_series(scene, series):
    ref = REF["scenes"][scene]["series"]
    assert len(series) == len(ref) == STEPS // EVERY
    for k, ((err, mx, ke), (rerr, rmx, rke)) in enumerate(zip(series, ref)):
        step = (k + 1) * EVERY
        assert abs(err - rerr) <= 0.35 * rerr + 0.005, f"scene {scene} step {step}: mean density error {err:.4f} vs reference {rerr:.4f}"
        assert 0.6 * rke <= ke <= 1.6 * rke, f"scene {scene} step {step}: kinetic energy {ke:.1f} vs reference {rke:.1f}"
        assert np.isfinite(mx) and mx < 1.5


def test_oracle_1000_steps_scene7_statistics():
    """pins the oracle's long-run behaviour to the reference's (CPU, ~25 s)"""
    g = G.load("7")
    o = G.oracle_for(g)
    rng = np.random.default_rng(1234)
    series = []
    for s in range(1, STEPS + 1):
        o.step(G.DT, rng.uniform(0, 1, (o.iterations, 6)).astype(np.float32))
        if s % EVERY == 0:
            series.append(o.fluid_stats())
    check_series("7", series)


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["7", "3"])
def test_gpu_1000_steps_statistics(scene):
    """the CUDA path (whole steps, CUDA graph, its own cuRAND wall jitter) over 1000 steps"""
    import particlesolver_b200 as psb
    import helpers as H
    ps = psb.ParticleSystem.scene(scene)
    sol = ps.solver
    o = H.oracle_from_solver(sol)  # used for the statistics only
    series = []
    for s in range(1, STEPS + 1):
        ps.update(G.DT)
        if s % EVERY == 0:
            o.pos[:] = sol.download(psb.ARR_POS)
            o.vel[:] = sol.download(psb.ARR_VEL)
            series.append(o.fluid_stats())
    assert np.isfinite(sol.download(psb.ARR_POS)).all()
    check_series(scene, series)
    ps.close()
