"""Build libpsolver.so (the product: hand-written sm_100a CUDA + C++ host) in-tree with nvcc.

Run as `python -m particlesolver_b200.build` or through `__graft_entry__.build()`.  The library is built for
sm_100a ONLY (`-gencode arch=compute_100a,code=sm_100a`); there is no other backend and no CPU path.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpsolver.so")
OBJ = os.path.join(HERE, "build")
CU_SOURCES = ["ps_stream_kernels.cu", "ps_grid_kernels.cu", "ps_sort_kernels.cu", "ps_neighbor_kernels.cu", "ps_fluid_staged.cu", "ps_slab_kernels.cu", "ps_shape_kernels.cu", "ps_context.cu", "ps_extensions.cu", "ps_checkpoint.cu", "ps_stream_io.cu", "ps_comm.cu",
              "ps_reference_abi.cu", "ps2d.cu"]
CPP_SOURCES = ["particle_system.cpp", "scenes2d.cpp"]
# -use_fast_math mirrors the reference's own build flags (gpu/particles_cuda.pro:153-158): div.approx / sqrt.approx /
# ftz are parity-relevant, see SURVEY Appendix A.1
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-use_fast_math", "-lineinfo", "-std=c++17", "-Xcompiler",
              "-fPIC,-fno-strict-aliasing", "-Xptxas", "-v"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src, *extra) if os.path.exists(p))


def build_variant(name, defines, verbose=False):
    """Tuning aid: libpsolver_<name>.so with extra -D flags (loaded through PS_LIBRARY, see __init__.lib)."""
    global OUT, OBJ, NVCC_FLAGS
    saved = (OUT, OBJ, NVCC_FLAGS)
    try:
        OUT = os.path.join(HERE, f"libpsolver_{name}.so")
        OBJ = os.path.join(HERE, "build", "variant_" + name)
        NVCC_FLAGS = NVCC_FLAGS + ["-D" + d for d in defines]
        return build_all(force=False, verbose=verbose)
    finally:
        OUT, OBJ, NVCC_FLAGS = saved


def build_all(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(OUT):
            return OUT  # GPU box without a toolchain: use the library that travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libpsolver.so")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", h) for h in os.listdir(os.path.join(HERE, "..", "include"))]
    objs, rebuilt = [], False
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)  # nvcc picks gcc from PATH; the image's $CC wrapper lacks libgomp specs
    for src in CU_SOURCES + CPP_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _newer(s, o, headers):
            flags = list(NVCC_FLAGS)
            if src == "ps2d.cu":  # double-precision parity path: IEEE arithmetic without contraction, like the x86-64 reference
                flags = [f for f in flags if f != "-use_fast_math"] + ["-fmad=false"]
            cmd = [nvcc, *flags, "-I", os.path.join(HERE, "..", "include"), "-c", s, "-o", o]
            if src.endswith(".cpp"):  # host class: plain g++, IEEE arithmetic without contraction (scene parity)
                cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
                cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-I",
                       os.path.join(HERE, "..", "include"), "-I", cuda_inc, "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on " + src)
            rebuilt = True
    if rebuilt or force or not os.path.exists(OUT):
        cmd = [nvcc, "-shared", "-o", OUT, *objs, "-lcurand", "-ldl", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    if os.path.basename(OUT) == "libpsolver.so":  # tuning variants (build_variant) leave the CLI on the default library
        build_cli(env)
    return OUT


def build_cli(env=None):
    """psolver_cli: the headless runner (csrc/psolver_cli.cpp), linked against the in-tree libpsolver.so"""
    cli, src = os.path.join(HERE, "psolver_cli"), os.path.join(CSRC, "psolver_cli.cpp")
    if not _newer(src, cli, [OUT]):
        return cli
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), "-I", cuda_inc, src, "-o", cli,
           "-L", HERE, "-l:" + os.path.basename(OUT), "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("psolver_cli link failed")
    return cli


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":  # python -m particlesolver_b200.build --variant q16 PS_KQ=16 ...
        print(build_variant(sys.argv[2], sys.argv[3:], verbose=False))
    else:
        print(build_all(force="--force" in sys.argv, verbose=True))
