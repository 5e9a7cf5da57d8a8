"""tests/golden/ref_cpu_scene6.npz — states of the reference's UNMODIFIED 2-D CPU solver (oracle/_ref/ref_cpu, built from
/root/reference/cpu/src by `make -C oracle ref`) on its scene 6 (FLUID_TEST: two fluids, 432 particles), config C1.
Runs anywhere (no GPU):   python tests/golden/make_cpu_golden.py
Kept: the initial state, the position of the glibc rand() stream at scene start (the solver draws wall jitter from it),
full states after ticks 1, 2, 3, 10, 100, and kinetic energy / mean density error every 100 ticks up to 1000."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
RAW = "/tmp/ref_cpu_scene6"
TICKS = [1, 2, 3, 10, 100]


def read_tick(t):
    raw = np.fromfile(os.path.join(RAW, f"tick{t:05d}.bin"), dtype=np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    rec = raw[4:].view(np.float64).reshape(n, 8)
    meta = {}
    for line in open(os.path.join(RAW, f"tick{t:05d}.txt")):
        k, *v = line.split()
        meta[k] = [float(x) for x in v]
    return rec, meta


def main():
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_cpu"), "--scene", "6", "--ticks", "1000", "--dump", RAW, "--dump-every", "1"], check=True)
    rec0, m0 = read_tick(0)
    keep = {"n": np.array(rec0.shape[0]), "p0": rec0[:, 0:2], "v0": rec0[:, 2:4], "imass": rec0[:, 4], "phase": rec0[:, 5].astype(np.int32),
            "fluid": rec0[:, 7].astype(np.int32), "rho0": np.array(m0["fluids"][1:]), "xbounds": np.array(m0["xbounds"]),
            "ybounds": np.array(m0["ybounds"]), "gravity": np.array(m0["gravity"]), "rand_calls0": np.array(int(m0["rand_calls"][0])), "dt": np.array(0.01)}
    for t in TICKS:
        rec, m = read_tick(t)
        keep[f"p{t}"], keep[f"v{t}"] = rec[:, 0:2], rec[:, 2:4]
        keep[f"rand_calls{t}"] = np.array(int(m["rand_calls"][0]))
    ke = []
    for t in range(100, 1001, 100):
        _, m = read_tick(t)
        ke.append(m["ke"][0])
    keep["ke_every_100"] = np.array(ke)
    keep["ke1"] = np.array(read_tick(1)[1]["ke"][0])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_cpu_scene6.npz"), **keep)
    print("ke after 1 tick", keep["ke1"], "rand calls at start", keep["rand_calls0"], "ke series", ke)


if __name__ == "__main__":
    main()
