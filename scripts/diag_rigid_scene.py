"""tuning aid: the rigid-box tower of extension scene "r" over time, with and without SDF data (PYTHONPATH=. python scripts/diag_rigid_scene.py)"""
import sys
import numpy as np, particlesolver_b200 as psb
for sdf in (True, False):
    ps = psb.ParticleSystem()
    for ll, ur in (((0, 1, 0), (3, 4, 3)), ((0, 5, 0), (3, 8, 3)), ((0, 9, 0), (3, 12, 3))):
        ps.addRigidBox(ll, ur, 1.0, sdf=sdf)
    x0 = ps.getPositions()[:, :3].astype(np.float64)
    boxes = [slice(125 * k, 125 * (k + 1)) for k in range(3)]
    print("sdf", sdf)
    for step in range(1, 481):
        ps.update(1/60)
        if step % 60 == 0:
            x = ps.getPositions()[:, :3].astype(np.float64); v = ps.getVelocities()[:, :3]
            gaps = [round(float(np.linalg.norm(x[boxes[a]][:, None] - x[boxes[b]][None], axis=2).min()), 3) for a in range(3) for b in range(a+1, 3)]
            print(step, "y", [round(float(x[s][:,1].mean()),2) for s in boxes], "x", [round(float(x[s][:,0].mean()),2) for s in boxes], "gaps", gaps, "ymin", round(float(x[:,1].min()),3), "vmax", round(float(np.abs(v).max()),2))
    ps.close()
