"""tests/golden/ref_cpu_scenes.npz — scene descriptions and states of the reference's UNMODIFIED 2-D CPU solver
(oracle/_ref/ref_cpu, built from /root/reference/cpu/src by `make -C oracle ref`) for every key-bound scene of its app
(cpu/src/view.cpp:129-177): SURVEY §8 row a19.
Runs anywhere (no GPU):   python tests/golden/make_cpu_scenes_golden.py
Per scene NAME: NAME_scene = the full restart state as JSON (particles incl. friction and force accumulators, rigid
bodies with r vectors / SDF / centre / angle, the STANDARD constraint list in order, smoke emitters, position of the glibc
rand() stream) taken after tick T0 (0 = the freshly built scene; later for scenes whose rigid contacts start late), and
NAME_p{t}, NAME_v{t}, NAME_rand{t} for a few ticks t > T0 (particle counts may grow: smoke emitters); NAME_scene0 = the
freshly built scene where T0 > 0; NAME_key = the app's key for the scene."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# name: (key, T0, ticks kept)
SCENES = {
    "granular": ("1", 0, [1, 2, 3, 10, 20]),
    "stacks": ("2", 100, [101, 102, 103, 105, 110]),
    "wall": ("3", 0, [1, 2, 3, 5, 10]),
    "pendulum": ("4", 0, [1, 2, 3, 10, 50]),
    "rope": ("5", 0, [1, 2, 3, 10, 20]),
    "fluid_solid": ("7", 0, [1, 2, 3, 10, 20]),
    "gas_rope": ("8", 0, [1, 2, 3, 10, 20]),
    "friction": ("9", 98, [99, 100, 101, 105, 110]),
    "balloon": ("0", 0, [1, 2, 3, 10, 20]),
    "cradle": ("n", 0, [1, 2, 3, 50, 100]),
    "smoke_open": ("s", 0, [1, 2, 3, 10, 20]),
    "smoke_closed": ("d", 0, [1, 2, 3, 10, 20]),
    "sdf": (".", 46, [47, 48, 49, 52, 56]),
    "wrecking_ball": ("w", 0, [1, 2, 3, 5, 10]),
    "volcano": ("v", 0, [1, 2, 3, 10, 20]),
    "volcano_freezing": ("v", 240, [241, 242, 243, 250, 260]),   # fluid particles freeze into solids from tick ~150 on
}


def read_tick(raw_dir, t):
    raw = np.fromfile(os.path.join(raw_dir, f"tick{t:05d}.bin"), dtype=np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    rec = raw[4:].view(np.float64).reshape(n, 8)
    meta = {}
    for line in open(os.path.join(raw_dir, f"tick{t:05d}.txt")):
        k, *v = line.split()
        meta[k] = [float(x) for x in v]
    return rec, meta


# --stab: the same for the reference compiled with its option USE_STABILIZATION (oracle/_ref/ref_cpu_stab; simulation.h:17,
# simulation.cpp:249-271), scenes with rigid contacts and / or walls -> tests/golden/ref_cpu_scenes_stab.npz
STAB_SCENES = ["stacks", "wall", "fluid_solid", "friction", "sdf", "wrecking_ball"]


def main():
    stab = "--stab" in sys.argv
    keep = {}
    for name, (key, t0, ticks) in SCENES.items():
        if stab and name not in STAB_SCENES:
            continue
        raw_dir = f"/tmp/ref_cpu_scene_{name}" + ("_stab" if stab else "")
        cmd = [os.path.join(ROOT, "oracle", "_ref", "ref_cpu_stab" if stab else "ref_cpu"), "--scene", key, "--ticks", str(max(ticks)), "--dump", raw_dir, "--dump-every", "1"]
        if t0:
            cmd += ["--scene-at", str(t0)]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        text = open(os.path.join(raw_dir, f"scene_t{t0:05d}.json" if t0 else "scene.json")).read()
        json.loads(text)
        keep[f"{name}_scene"] = np.array(text)
        if t0:  # the freshly built scene as well (what the scene builders are compared with)
            keep[f"{name}_scene0"] = np.array(open(os.path.join(raw_dir, "scene.json")).read())
        keep[f"{name}_key"] = np.array(key)
        keep[f"{name}_t0"] = np.array(t0)
        keep[f"{name}_ticks"] = np.array(ticks)
        for t in ticks:
            rec, m = read_tick(raw_dir, t)
            keep[f"{name}_p{t}"], keep[f"{name}_v{t}"] = rec[:, 0:2].copy(), rec[:, 2:4].copy()
            keep[f"{name}_rand{t}"] = np.array(int(m["rand_calls"][0]))
            keep[f"{name}_ke{t}"] = np.array(m["ke"][0])
        print(name, "n", json.loads(text)["n"], "->", keep[f"{name}_p{ticks[-1]}"].shape[0])
    out = os.path.join(ROOT, "tests", "golden", "ref_cpu_scenes_stab.npz" if stab else "ref_cpu_scenes.npz")
    np.savez_compressed(out, **keep)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
