"""Long-run statistics of the reference's UNMODIFIED GPU solver (oracle/_ref/ref_gpu) for the 1000-step parity bar of
north_star: mean / max density error |rho/rho0 - 1| and kinetic energy every 100 steps, for the fluid scenes 7 and 3.
Run on a GPU box from the repo root:   python tests/golden/make_stats_golden.py [out_dir]
                                       python tests/golden/make_stats_golden.py c3 [out_dir]   -> ref_gpu_stats_c3.json: the headline
                                       workload (scene 7 scaled to 1,000,000 particles, 256^3 grid), 100 steps, statistics every 20
The fixture (tests/golden/ref_gpu_stats.json) is small; per-particle comparison is meaningless after ~10 steps (chaotic
divergence), so the long-run bar is statistical."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import oracle_py as orc  # noqa: E402
import pack_golden  # noqa: E402

STEPS, EVERY = 1000, 100
SCENES = {"7": dict(min_b=(-50, 0, -50), max_b=(50, 200, 50)), "3": dict(min_b=(-7, 0, -5), max_b=(7, 20, 5))}


def stats_series(p, series, w, phase, ros):
    out = []
    for pos, vel in series:
        o = orc.OracleSystem(p, pos, vel, w, phase, ros)
        out.append([float(x) for x in o.fluid_stats()])
    return out


def main_c3(out_dir):
    """BASELINE config C3 at full size: the reference's own CUDA code for 100 steps (the blob expands, falls and meets the floor)"""
    steps, every, side, grid = 100, 20, 100, 256
    os.makedirs(out_dir, exist_ok=True)
    raw = "/tmp/ref_stats_c3"
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_gpu"), "--scene", "c3", "--grid", str(grid), "--side", str(side), "--max", str(side ** 3 + 1024),
                    "--mode", "whole", "--steps", str(steps), "--dump-every", str(every), "--out", raw], check=True)
    arrs, meta = pack_golden.load_dump(raw)
    p = orc.make_params(grid=tuple(int(x) for x in meta["grid"]), **SCENES["7"])
    series = [(arrs[f"w{s}_pos"], arrs[f"w{s}_vel"]) for s in range(every, steps + 1, every)]
    res = {"steps": steps, "every": every, "dt": 1.0 / 60.0, "side": side, "grid": grid, "n": int(meta["n"][0]),
           "columns": ["mean_density_error", "max_density_error", "kinetic_energy"],
           "series": stats_series(p, series, arrs["init_w"], arrs["init_phase"], arrs["init_ros"])}
    print("c3", res["series"])
    json.dump(res, open(os.path.join(out_dir, "ref_gpu_stats_c3.json"), "w"), indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "c3":
        return main_c3(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "golden"))
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    res = {"steps": STEPS, "every": EVERY, "dt": 1.0 / 60.0, "columns": ["mean_density_error", "max_density_error", "kinetic_energy"], "scenes": {}}
    for scene, b in SCENES.items():
        raw = f"/tmp/ref_stats_scene{scene}"
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_gpu"), "--scene", scene, "--mode", "whole", "--steps", str(STEPS), "--dump-every",
                        str(EVERY), "--out", raw], check=True)
        arrs, meta = pack_golden.load_dump(raw)
        p = orc.make_params(grid=tuple(int(x) for x in meta["grid"]), **b)
        series = [(arrs[f"w{s}_pos"], arrs[f"w{s}_vel"]) for s in range(EVERY, STEPS + 1, EVERY)]
        res["scenes"][scene] = {"n": int(meta["n"][0]), "series": stats_series(p, series, arrs["init_w"], arrs["init_phase"], arrs["init_ros"])}
        print(scene, res["scenes"][scene]["series"][0], res["scenes"][scene]["series"][-1])
    json.dump(res, open(os.path.join(out_dir, "ref_gpu_stats.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
