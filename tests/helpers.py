"""Shared helpers of the parity tests: scene -> oracle mirror, comparison with stated tolerances."""
import os

import numpy as np

import oracle_py as orc
import particlesolver_b200 as psb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Tolerances.  Integer grid output: exact.  Positions after a stage / a step: the CUDA
# path uses approximate divide / rsqrt / FMA contraction (the reference's own -use_fast_math build does too),
# the oracle IEEE float32 without contraction; both accumulate ~150 neighbour terms of magnitude <= 1.
POS_ATOL = 1e-4          # per-particle |dx|_inf after one stage or one step, world units (particle radius 0.25)
LAMBDA_RTOL = 2e-3       # lambda is a ratio of O(1e2) sums; compare relatively
VEL_ATOL = POS_ATOL * 60 * 1.01  # v = dx / dt with dt = 1/60


def oracle_from_solver(sol: "psb.Solver"):
    """Mirror the state of a libpsolver context into an OracleSystem (same inputs, same constraint order)."""
    p = sol.params
    op = orc.make_params(radius=p.particle_radius, grid=tuple(p.grid_size), min_b=tuple(p.min_bounds), max_b=tuple(p.max_bounds),
                         gravity=tuple(p.gravity), origin=tuple(p.world_origin), cell=tuple(p.cell_size))
    didx, drest = sol.distance_constraints()
    pidx, pxyz = sol.point_constraints()
    o = orc.OracleSystem(op, sol.download(psb.ARR_POS), sol.download(psb.ARR_VEL), sol.download(psb.ARR_INV_MASS),
                         sol.download(psb.ARR_PHASE), sol.download(psb.ARR_REST_DENSITY), didx, drest, pidx, pxyz,
                         iterations=int(p.solver_iterations))
    o.self_collision = bool(p.flags & psb.FLAG_SELF_COLLISION)
    o.omega = float(p.omega)
    return o


def max_abs(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)))


def assert_grid_equal(sol, o, tag=""):
    """hash / index / cellStart / cellEnd bit-exact; cellEnd only where cellStart marks the cell non-empty
    (the reference never clears it: integration.cu:199)."""
    h, i = sol.download(psb.ARR_HASH), sol.download(psb.ARR_INDEX)
    cs, ce = sol.download(psb.ARR_CELL_START), sol.download(psb.ARR_CELL_END)
    assert np.array_equal(h, o.hash), f"{tag} sorted hash differs"
    assert np.array_equal(i, o.index), f"{tag} sorted index differs"
    assert np.array_equal(cs, o.cell_start), f"{tag} cellStart differs"
    valid = o.cell_start != 0xFFFFFFFF
    assert np.array_equal(ce[valid], o.cell_end[valid]), f"{tag} cellEnd differs"
    # the sorted copies are exact gathers
    assert np.array_equal(sol.download(psb.ARR_SORTED_POS), o.spos), f"{tag} sortedPos differs"
    assert np.array_equal(sol.download(psb.ARR_SORTED_INV_MASS), o.sw)
    assert np.array_equal(sol.download(psb.ARR_SORTED_PHASE), o.sphase)
    # dense table: lower bound of every cell key in the sorted hash array
    cb = sol.download(psb.ARR_CELL_BEGIN)
    expect = np.searchsorted(o.hash, np.arange(o.num_cells + 1, dtype=np.uint64), side="left").astype(np.uint32)
    assert np.array_equal(cb, expect), f"{tag} cell_begin table differs"
