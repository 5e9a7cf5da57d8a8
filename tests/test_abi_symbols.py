"""The C-ABI shared library loads on a GPU-less machine and exports every symbol the public headers declare.
No compute calls here; creating a context without a device must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import HAS_GPU

import particlesolver_b200 as psb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, pattern):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(pattern, src, flags=re.M)))


def test_library_is_built_in_tree():
    assert os.path.exists(psb.LIB_PATH), "run `python -m particlesolver_b200.build`"


def test_psolver_h_symbols_exported():
    L = psb.lib()
    names = _declared("psolver.h", r"\b(ps_[a-z_0-9]+)\s*\(")
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/psolver.h but not exported: {missing}"


def test_psolver2d_h_symbols_exported():
    L = psb.lib()
    names = _declared("psolver2d.h", r"\b(ps2d_[a-z_0-9]+)\s*\(")
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/psolver2d.h but not exported: {missing}"


def test_reference_abi_symbols_exported():
    """Same names as the reference's wrappers.cuh:12-97, shared_variables.cuh:13-36 and the non-GL part of util.cuh:6-25."""
    L = psb.lib()
    names = _declared("ps_reference_abi.h", r"^\s*(?:void|int|float|uint)\s*\*?\s*([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    names = [n for n in names if n not in ("defined",)]
    reference_wrappers = ["initIntegration", "freeIntegrationVectors", "appendIntegrationParticle", "setParameters", "integrateSystem",
                          "calcHash", "sortParticles", "reorderDataAndFindCellStart", "collideWorld", "collide", "sortByType",
                          "calcVelocity", "solveFluids", "appendSolverParticle", "addPointConstraint", "addDistanceConstraint",
                          "freeSolverVectors", "solvePointConstraints", "solveDistanceConstraints", "freeSharedVectors",
                          "appendPhaseAndMass", "copyToXstar", "getPhaseRawPtr", "getXstarRawPtr", "getWRawPtr", "printXstar",
                          "cudaInit", "allocateArray", "freeArray", "copyArrayToDevice", "copyArrayFromDevice", "iDivUp", "computeGridSize"]
    for n in reference_wrappers:
        assert n in names, f"{n} not declared in ps_reference_abi.h"
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared but not exported: {missing}"


def test_sim_params_layout_matches_reference():
    # SimParams, gpu/src/cuda/kernel.cuh:9-22: 3+1+1 floats, 3+1 uints, 3+3 floats, 2 uints = 68 bytes
    class PsRefSimParams(ctypes.Structure):
        _fields_ = [("gravity", ctypes.c_float * 3), ("globalDamping", ctypes.c_float), ("particleRadius", ctypes.c_float),
                    ("gridSize", ctypes.c_uint * 3), ("numCells", ctypes.c_uint), ("worldOrigin", ctypes.c_float * 3),
                    ("cellSize", ctypes.c_float * 3), ("numBodies", ctypes.c_uint), ("maxParticlesPerCell", ctypes.c_uint)]
    assert ctypes.sizeof(PsRefSimParams) == 68


def test_defaults_are_the_reference_constants():
    p = psb.default_params()
    assert tuple(p.gravity) == (0.0, pytest.approx(-9.8), 0.0)          # particlesystem.cpp:65
    assert p.particle_radius == 0.25 and tuple(p.cell_size) == (0.5, 0.5, 0.5)  # particleapp.cpp:24, particlesystem.cpp:62
    assert tuple(p.grid_size) == (64, 64, 64) and p.solver_iterations == 5  # particleapp.cpp:25,38
    assert tuple(p.min_bounds) == (-50, 0, -50) and tuple(p.max_bounds) == (50, 200, 50)
    assert p.omega == 1.0 and p.flags == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(psb.PsError):
        psb.Solver(psb.default_params(), max_particles=16)
    with pytest.raises(psb.PsError):
        psb.ParticleSystem()
    with pytest.raises(psb.PsError):
        psb.Simulation2D()


def test_bad_params_rejected_before_touching_the_device():
    p = psb.default_params()
    p.grid_size[0] = 100  # not a power of two: the reference's '&' wrap would alias garbage
    with pytest.raises(psb.PsError) as e:
        psb.Solver(p, max_particles=16)
    assert e.value.code == psb.PS_ERR_INVALID


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "particlesolver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in src and "libpsoracle" not in src and "gpu_step_oracle" not in src and "cpu2d_oracle" not in src, f


def test_ps_scenes2d_h_symbols_exported():
    names = _declared("ps_scenes2d.h", r"\b(ps2d_[a-z_0-9]+)\s*\(")
    assert {"ps2d_build_scene", "ps2d_scene_name"} <= set(names)
    missing = [n for n in names if not hasattr(psb.lib(), n)]
    assert not missing, f"declared in include/ps_scenes2d.h but not exported: {missing}"


def test_checkpoint_loaders_reject_foreign_files_before_touching_the_gpu(tmp_path):
    import ctypes as C
    bad = tmp_path / "not_a_checkpoint.bin"
    bad.write_bytes(b"hello world, definitely not a checkpoint" * 4)
    h = C.c_void_p()
    assert psb.lib().ps_load(str(bad).encode(), 0, C.byref(h)) == psb.PS_ERR_INVALID and not h.value
    assert b"not a libpsolver checkpoint" in psb.lib().ps_last_error()
    assert psb.lib().ps2d_load(str(bad).encode(), 0, C.byref(h)) == psb.PS_ERR_INVALID and not h.value
    assert psb.lib().ps_load(str(tmp_path / "missing").encode(), 0, C.byref(h)) == psb.PS_ERR_INVALID


def test_headless_cli_is_built_and_fails_loudly_without_a_gpu():
    import subprocess
    cli = os.path.join(ROOT, "particlesolver_b200", "psolver_cli")
    assert os.path.exists(cli), "python -m particlesolver_b200.build builds it next to libpsolver.so"
    assert subprocess.run([cli, "--help"], capture_output=True).returncode == 0
    if not HAS_GPU:
        r = subprocess.run([cli, "--app", "cpu", "--scene", "6", "--ticks", "1"], capture_output=True, text=True)
        assert r.returncode == 1 and "ps2d_build_scene" in r.stderr   # no CPU fallback
