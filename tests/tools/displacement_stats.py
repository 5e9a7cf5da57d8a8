#!/usr/bin/env python
"""How far do particles move between the first grid build of a step and the later solver iterations?  (why a
candidate superset kept from iteration 1 and re-tested in iterations 2-5 — a Verlet list with a skin — was not built.)
CPU only (the oracle): python tests/tools/displacement_stats.py [side] [steps]   — a side^3 block of the c3 scene (spacing 2.5 r, rho0 1.5)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/: the oracle binding (test infrastructure)
import oracle_py as orc  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
rng = np.random.default_rng(1)
g = np.arange(side) * 0.625
X, Y, Z = np.meshgrid(g - side * 0.3125, g + 6, g - side * 0.3125, indexing="ij")
pos = np.stack([X.ravel(), Y.ravel(), Z.ravel(), np.ones(X.size)], 1).astype(np.float32)
pos[:, :3] += rng.uniform(-0.0025, 0.0025, (pos.shape[0], 3)).astype(np.float32)
n = pos.shape[0]
p = orc.make_params(grid=(128, 128, 128), min_b=(-50, 0, -50), max_b=(50, 200, 50), origin=(-32, 0, -32))
o = orc.OracleSystem(p, pos, np.zeros((n, 4), np.float32), np.ones(n, np.float32), np.zeros(n, np.int32), np.full(n, 1.5, np.float32))
rs = np.random.default_rng(2)
print("step  neighbours  displacement since the step's first grid build, at the last iteration: max / p99.9 / p99, share > 0.1, share > 0.2")
for s in range(steps):
    o.predict(1 / 60)
    x1, last = None, None
    for it in range(5):
        o.build_grid()
        if it == 0:
            x1 = o.pos.copy()
        else:
            d = np.linalg.norm(o.pos[:, :3] - x1[:, :3], axis=1)
            last = (d.max(), np.percentile(d, 99.9), np.percentile(d, 99), (d > 0.1).mean(), (d > 0.2).mean())
        o.solve_fluids()
        o.collide_world(rs.uniform(0, 1, 6).astype(np.float32))
    o.calc_velocity(1 / 60)
    if s % 5 == 0 or s < 6:
        print(f"{s:4d}  {o.nn.mean():8.0f}   max {last[0]:.3f}  p99.9 {last[1]:.3f}  p99 {last[2]:.3f}  >0.1: {last[3]:.4f}  >0.2: {last[4]:.4f}")
