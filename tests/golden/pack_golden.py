"""Pack raw dumps of oracle/_ref/ref_gpu (the reference's unmodified GPU solver run on a B200) into the small
.npz fixtures committed under tests/golden/.

Generation (on a GPU box; /root/reference is NOT needed there, the binary was built here by `make -C oracle ref`):
    bash tests/golden/make_golden.sh        # runs ref_gpu per scene, then this script
Arrays identical to the previous stage's are stored once (key "<name>__same_as").  cellEnd is stored with the
entries of empty cells zeroed: the reference leaves them unspecified (integration.cu:199).
"""
import os
import sys

import numpy as np

DT = {"f32": np.float32, "u32": np.uint32, "i32": np.int32}


def load_dump(d):
    out = {}
    for line in open(os.path.join(d, "manifest.txt")):
        name, dt, cnt = line.split()
        a = np.fromfile(os.path.join(d, name + ".bin"), dtype=DT[dt])
        assert a.size == int(cnt), (name, a.size, cnt)
        out[name] = a
    meta = {}
    for line in open(os.path.join(d, "scene.txt")):
        k, *v = line.split()
        meta[k] = v
    return out, meta


def pack(d, dst):
    arrs, meta = load_dump(d)
    keep = {}
    keep["meta_n"] = np.array(int(meta["n"][0]))
    keep["meta_radius"] = np.array(float(meta["radius"][0]), np.float32)
    keep["meta_grid"] = np.array([int(x) for x in meta["grid"]], np.uint32)
    keep["meta_min"] = np.array([int(x) for x in meta["min"]], np.int32)
    keep["meta_max"] = np.array([int(x) for x in meta["max"]], np.int32)
    keep["meta_iters"] = np.array(int(meta["iters"][0]))
    last_pos_name, last_pos = None, None
    for name, a in arrs.items():  # manifest order == execution order
        if name.endswith("cell_end"):
            cs = arrs[name[:-len("cell_end")] + "cell_start"]
            a = np.where(cs != 0xFFFFFFFF, a, 0).astype(np.uint32)
        if name.endswith("sorted_pos") or name.endswith("sorted_w") or name.endswith("sorted_phase") or name.endswith("hash_unsorted"):
            if "_i0_" not in name:
                continue  # exact gathers / recomputable: keep the first iteration only
        if name.endswith("_pos") and not name.endswith("sorted_pos") and a.dtype == np.float32:
            if last_pos is not None and a.shape == last_pos.shape and np.array_equal(a.view(np.uint32), last_pos.view(np.uint32)):
                keep[name + "__same_as"] = np.array(last_pos_name)
                continue
            last_pos_name, last_pos = name, a
        keep[name] = a
    np.savez_compressed(dst, **keep)
    print(dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    src_root, dst_root = sys.argv[1], sys.argv[2]
    os.makedirs(dst_root, exist_ok=True)
    for s in sorted(os.listdir(src_root)):
        d = os.path.join(src_root, s)
        if os.path.exists(os.path.join(d, "manifest.txt")) and os.path.exists(os.path.join(d, "scene.txt")):
            pack(d, os.path.join(dst_root, f"ref_gpu_{s}.npz"))
