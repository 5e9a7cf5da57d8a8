"""Loader for the committed fixtures tests/golden/ref_gpu_scene*.npz (dumps of the reference's unmodified GPU
solver, produced by tests/golden/make_golden.sh on a B200)."""
import os

import numpy as np

import oracle_py as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = ["1", "2", "3", "4", "5", "6", "7", "8", "9"]  # every key-bound scene of the GPU app (particleapp.cpp:141-215)
DT = np.float32(1.0 / 60.0)


def load(scene, root=GOLDEN_DIR):
    z = np.load(os.path.join(root, f"ref_gpu_scene{scene}.npz"))
    d = {k: z[k] for k in z.files if not k.endswith("__same_as")}
    for k in z.files:
        if k.endswith("__same_as"):
            d[k[: -len("__same_as")]] = d[str(z[k])]
    return d


def oracle_for(g):
    p = orc.make_params(radius=float(g["meta_radius"]), grid=tuple(int(x) for x in g["meta_grid"]),
                        min_b=tuple(int(x) for x in g["meta_min"]), max_b=tuple(int(x) for x in g["meta_max"]))
    return orc.OracleSystem(p, g["init_pos"], g["init_vel"], g["init_w"], g["init_phase"], g["init_ros"], g["dist_idx"], g["dist_rest"],
                            g["point_idx"], g["point_xyz"], iterations=int(g["meta_iters"]))
