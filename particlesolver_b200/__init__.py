"""particlesolver_b200 — B200-native unified particle solver step (PBF fluids, contacts, cloth/rope constraints).

Python here is plumbing only: it loads the in-tree `libpsolver.so` (hand-written sm_100a CUDA kernels + the C++
host class) through its C ABI (include/psolver.h) with ctypes.  There is NO CPU path: importing works without a
GPU (so that symbols can be checked), but creating a solver without an sm_100 device raises.

Mirrors the reference's host interface (ebirenbaum/ParticleSolver gpu/src/particlesystem.h:22-52):
`ParticleSystem(radius, grid, max_particles, min_bounds, max_bounds, iterations)` with `addFluid`,
`addParticleGrid`, `addHorizCloth`, `addRope`, `addStaticSphere`, `update(dt)`, ... implemented by the C++ class
in csrc/particle_system.cpp (this module only forwards).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsolver.so")

# phase codes, reference gpu/src/cuda/shared_variables.cuh:4-9
NO_COLLIDE, FLUID, GAS, CLOTH, SOLID, RIGID = -1, 0, 1, 2, 3, 4

PS_OK, PS_ERR_INVALID, PS_ERR_CUDA, PS_ERR_CAPACITY, PS_ERR_STATE = 0, 1, 2, 3, 4
FLAG_ZERO_NONFLUID_LAMBDA = 1
FLAG_GAS = 2
FLAG_SELF_COLLISION = 4
FLAG_STAGED_LAMBDA = 8
NUM_STAGES = 12

(ARR_POS, ARR_VEL, ARR_PREV, ARR_INV_MASS, ARR_PHASE, ARR_REST_DENSITY, ARR_HASH, ARR_INDEX, ARR_CELL_START, ARR_CELL_END,
 ARR_SORTED_POS, ARR_SORTED_INV_MASS, ARR_SORTED_PHASE, ARR_LAMBDA, ARR_NUM_NEIGHBORS, ARR_RANDS, ARR_OCCURRENCES,
 ARR_CELL_BEGIN, ARR_NEIGHBOR_ROWS) = range(19)
_ARR_DTYPE = {ARR_POS: np.float32, ARR_VEL: np.float32, ARR_PREV: np.float32, ARR_INV_MASS: np.float32, ARR_PHASE: np.int32,
              ARR_REST_DENSITY: np.float32, ARR_HASH: np.uint32, ARR_INDEX: np.uint32, ARR_CELL_START: np.uint32,
              ARR_CELL_END: np.uint32, ARR_SORTED_POS: np.float32, ARR_SORTED_INV_MASS: np.float32, ARR_SORTED_PHASE: np.int32,
              ARR_LAMBDA: np.float32, ARR_NUM_NEIGHBORS: np.uint32, ARR_RANDS: np.float32, ARR_OCCURRENCES: np.uint32,
              ARR_CELL_BEGIN: np.uint32, ARR_NEIGHBOR_ROWS: np.uint32}
_ARR_WIDTH = {ARR_POS: 4, ARR_VEL: 4, ARR_PREV: 4, ARR_SORTED_POS: 4}


class PsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpsolver error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """PsParams (include/psolver.h) — superset of the reference's SimParams (gpu/src/cuda/kernel.cuh:9-22)."""
    _fields_ = [("gravity", C.c_float * 3), ("global_damping", C.c_float), ("particle_radius", C.c_float),
                ("grid_size", C.c_uint32 * 3), ("world_origin", C.c_float * 3), ("cell_size", C.c_float * 3),
                ("min_bounds", C.c_int32 * 3), ("max_bounds", C.c_int32 * 3), ("solver_iterations", C.c_uint32),
                ("omega", C.c_float), ("flags", C.c_uint32), ("neighbor_list_rows", C.c_uint32)]


class Params2D(C.Structure):
    """Ps2dParams (include/psolver2d.h): bounds, gravity and iteration count of the reference's CPU Simulation."""
    _fields_ = [("x_bounds", C.c_double * 2), ("y_bounds", C.c_double * 2), ("gravity", C.c_double * 2), ("solver_iterations", C.c_uint32),
                ("stabilization_iterations", C.c_uint32)]


_lib = None


def lib():
    """The loaded libpsolver.so.  Raises if the extension has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("PS_LIBRARY", LIB_PATH)  # tuning builds (build.build_variant); always a libpsolver build
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `python -m particlesolver_b200.build` "
                              "(sm_100a CUDA extension; there is no CPU or PyTorch fallback)")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
        L.ps_last_error.restype = C.c_char_p
        L.ps_version.restype = C.c_char_p
        L.ps_default_params.argtypes = [C.POINTER(Params)]
        L.ps_default_params.restype = None
        L.ps_create.argtypes = [i32, C.POINTER(Params), u64, C.POINTER(vp)]
        L.ps_destroy.argtypes = [vp]
        L.ps_set_params.argtypes = [vp, C.POINTER(Params)]
        L.ps_get_params.argtypes = [vp, C.POINTER(Params)]
        for f in ("ps_num_particles", "ps_num_cells", "ps_num_owned"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u64
        L.ps_launches_per_step.argtypes = [vp]
        L.ps_launches_per_step.restype = u32
        L.ps_append_particles.argtypes = [vp, vp, vp, vp, vp, vp, u64]
        L.ps_add_distance_constraints.argtypes = [vp, vp, vp, u64]
        L.ps_add_point_constraints.argtypes = [vp, vp, vp, u64]
        L.ps_num_distance_constraints.argtypes = [vp]
        L.ps_num_distance_constraints.restype = u64
        L.ps_num_point_constraints.argtypes = [vp]
        L.ps_num_point_constraints.restype = u64
        L.ps_copy_distance_constraints.argtypes = [vp, vp, vp]
        L.ps_copy_point_constraints.argtypes = [vp, vp, vp]
        L.ps_timer_start.argtypes = [vp]
        L.ps_timer_stop.argtypes = [vp, C.POINTER(f32)]
        L.ps_step_profiled.argtypes = [vp, f32, vp, vp]
        L.ps_stage_name.argtypes = [i32]
        L.ps_stage_name.restype = C.c_char_p
        L.ps_step.argtypes = [vp, f32]
        L.ps_step_streamed.argtypes = [vp, f32, vp, vp, vp, vp]
        L.ps_io_wait.argtypes = [vp, u32]
        L.ps_io_begin.argtypes = [vp, vp, vp]
        L.ps_io_prefetch.argtypes = [vp, vp, vp]
        L.ps_io_end.argtypes = [vp, vp, vp]
        L.ps_sync.argtypes = [vp]
        L.ps_last_step_ms.argtypes = [vp, C.POINTER(f32)]
        for f in ("ps_begin_step", "ps_build_grid", "ps_solve_contacts", "ps_solve_fluid", "ps_solve_fluid_lambda", "ps_solve_fluid_delta", "ps_solve_distance",
                  "ps_solve_point"):
            getattr(L, f).argtypes = [vp]
        L.ps_predict.argtypes = [vp, f32]
        L.ps_update_velocity.argtypes = [vp, f32]
        L.ps_collide_world.argtypes = [vp, u32]
        for f in ("ps_download", "ps_upload", "ps_download_async", "ps_upload_async"):
            getattr(L, f).argtypes = [vp, i32, vp, u64, u64]
        L.ps_device_ptr.argtypes = [vp, i32]
        L.ps_device_ptr.restype = vp
        L.ps_stream.argtypes = [vp]
        L.ps_stream.restype = vp
        L.ps_set_ghost_count.argtypes = [vp, u64]
        L.ps_slab_pack_halo.argtypes = [vp, f32, f32, f32, vp, vp, u64, C.POINTER(u32 * 2)]
        L.ps_slab_set_ghosts.argtypes = [vp, vp, u64, vp, u64]
        L.ps_slab_pack_migrants.argtypes = [vp, f32, f32, vp, vp, u64, C.POINTER(u32 * 2)]
        L.ps_slab_append_migrants.argtypes = [vp, vp, u64, vp, u64]
        L.ps_slab_set_lambda_range.argtypes = [vp, f32, f32]
        L.ps_slab_pack_lambda.argtypes = [vp, vp, vp, u64, C.POINTER(u32 * 2)]
        L.ps_slab_set_ghost_lambda.argtypes = [vp, vp, u64, vp, u64]
        L.ps_slab_set_lambda_sinks.argtypes = [vp, vp, vp, u64]
        L.ps_slab_x_histogram.argtypes = [vp, f32, f32, u32, vp]
        L.ps_comm_get_unique_id.argtypes = [vp]
        L.ps_comm_init.argtypes = [vp, vp, i32, i32]
        L.ps_comm_destroy.argtypes = [vp]
        L.ps_comm_set_slab.argtypes = [vp, f32, f32, f32, i32, u64, u64]
        L.ps_comm_step.argtypes = [vp, f32]
        L.ps_comm_set_recut.argtypes = [vp, vp, u32, f32, f32, u32]
        L.ps_comm_get_cuts.argtypes = [vp, vp, vp]
        L.ps_comm_stats.argtypes = [vp, vp]
        L.ps_comm_allreduce_sum.argtypes = [vp, vp, u32]
        # C++ host class (csrc/particle_system.cpp)
        L.pshost_create.argtypes = [f32, u32, u32, u32, u32, vp, vp, i32]
        L.pshost_create.restype = vp
        L.pshost_build_scene.argtypes = [C.c_char_p, i32, u32, i32, i32, i32]
        L.pshost_build_scene.restype = vp
        L.pshost_destroy.argtypes = [vp]
        L.pshost_destroy.restype = None
        L.pshost_ctx.argtypes = [vp]
        L.pshost_ctx.restype = vp
        L.pshost_error.argtypes = [vp]
        L.pshost_error.restype = C.c_char_p
        L.pshost_num_particles.argtypes = [vp]
        L.pshost_num_particles.restype = u32
        L.pshost_update.argtypes = [vp, f32]
        L.pshost_update.restype = None
        L.pshost_add_fluid.argtypes = [vp, vp, vp, f32, f32]
        L.pshost_add_particle_grid.argtypes = [vp, vp, vp, f32, i32]
        L.pshost_add_rigid_box.argtypes = [vp, vp, vp, f32, i32, f32]
        L.pshost_add_rigid_box.restype = i32
        L.pshost_add_horiz_cloth.argtypes = [vp, vp, vp, vp, vp, f32, i32]
        L.pshost_add_rope.argtypes = [vp, vp, vp, f32, i32, f32, i32]
        L.pshost_add_static_sphere.argtypes = [vp, vp, vp, f32]
        L.pshost_set_particle_to_add.argtypes = [vp, vp, vp, f32]
        L.pshost_set_fluid_to_add.argtypes = [vp, vp, vp, f32, f32]
        L.pshost_make_point_constraint.argtypes = [vp, u32, vp]
        L.pshost_make_distance_constraint.argtypes = [vp, u32, u32, f32]
        L.pshost_get_positions.argtypes = [vp, vp]
        L.pshost_get_velocities.argtypes = [vp, vp]
        for f in ("pshost_add_fluid", "pshost_add_particle_grid", "pshost_add_horiz_cloth", "pshost_add_rope", "pshost_add_static_sphere",
                  "pshost_set_particle_to_add", "pshost_set_fluid_to_add", "pshost_make_point_constraint",
                  "pshost_make_distance_constraint", "pshost_get_positions", "pshost_get_velocities"):
            getattr(L, f).restype = None
        L.ps_fluid_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ps_add_rigid_body.argtypes = [vp, vp, u64, f32, C.POINTER(u32)]
        L.ps_num_rigid_bodies.argtypes = [vp]
        L.ps_num_rigid_bodies.restype = u64
        L.ps_solve_shapes.argtypes = [vp]
        L.ps_rigid_body_rotation.argtypes = [vp, u32, vp]
        L.ps_set_rigid_body_sdf.argtypes = [vp, u32, vp]
        L.ps_set_viscosity.argtypes = [vp, f32, f32]
        L.ps_find_neighbors.argtypes = [vp]
        L.ps_apply_viscosity.argtypes = [vp, f32]
        # 2-D double-precision path (include/psolver2d.h)
        L.ps2d_default_params.argtypes = [C.POINTER(Params2D)]
        L.ps2d_default_params.restype = None
        L.ps2d_create.argtypes = [i32, C.POINTER(Params2D), u64, C.POINTER(vp)]
        L.ps2d_destroy.argtypes = [vp]
        L.ps2d_create_fluid.argtypes = [vp, vp, vp, vp, u64, C.c_double]
        L.ps_save.argtypes = [vp, C.c_char_p]
        L.ps_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
        L.ps2d_save.argtypes = [vp, C.c_char_p]
        L.ps2d_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
        L.ps2d_build_scene.argtypes = [C.c_char_p, i32, u64, C.POINTER(vp)]
        L.ps2d_scene_name.argtypes = [C.c_char_p]
        L.ps2d_scene_name.restype = C.c_char_p
        L.ps2d_rand.argtypes = [vp, C.POINTER(i32)]
        L.ps2d_add_particles.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u64, C.POINTER(u64)]
        L.ps2d_add_distance_constraint.argtypes = [vp, u32, u32, C.c_double]
        L.ps2d_add_fluid_constraint.argtypes = [vp, vp, u64, C.c_double, C.POINTER(u32)]
        L.ps2d_add_gas_constraint.argtypes = [vp, vp, u64, C.c_double, i32, C.POINTER(u32)]
        L.ps2d_restore_rigid_body.argtypes = [vp, u32, u32, vp, vp, C.c_double, vp, C.c_double, C.c_double, C.POINTER(u32)]
        L.ps2d_create_rigid_body.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64, C.POINTER(u32)]
        L.ps2d_create_gas.argtypes = [vp, vp, vp, vp, u64, C.c_double, i32, C.POINTER(u32)]
        L.ps2d_create_smoke_emitter.argtypes = [vp, vp, C.c_double, u32, C.c_double]
        L.ps2d_set_forces.argtypes = [vp, vp]
        L.ps2d_mouse_pressed.argtypes = [vp, C.c_double, C.c_double]
        L.ps2d_build_scene_from.argtypes = [C.c_char_p, i32, u64, u32, u64, C.POINTER(vp)]
        L.ps2d_create_fluid_emitter.argtypes = [vp, vp, C.c_double, u32, C.c_double, C.c_double]
        L.ps2d_set_particle_timers.argtypes = [vp, vp]
        L.ps2d_get_particle_timers.argtypes = [vp, vp]
        L.ps2d_body_state.argtypes = [vp, u32, vp, C.POINTER(C.c_double)]
        L.ps2d_num_bodies.argtypes = [vp]
        L.ps2d_num_bodies.restype = u32
        L.ps2d_last_num_contact_constraints.argtypes = [vp]
        L.ps2d_last_num_contact_constraints.restype = u32
        L.ps2d_last_num_levels.argtypes = [vp]
        L.ps2d_last_num_levels.restype = u32
        L.ps2d_seed_rand.argtypes = [vp, u32, u64]
        L.ps2d_set_stabilization_iterations.argtypes = [vp, u32]
        L.ps2d_rand_calls.argtypes = [vp]
        L.ps2d_rand_calls.restype = u64
        L.ps2d_tick.argtypes = [vp, C.c_double]
        L.ps2d_num_particles.argtypes = [vp]
        L.ps2d_num_particles.restype = u64
        L.ps2d_last_num_boundary_constraints.argtypes = [vp]
        L.ps2d_last_num_boundary_constraints.restype = u32
        L.ps2d_launches_per_tick.argtypes = [vp]
        L.ps2d_launches_per_tick.restype = u32
        L.ps2d_download.argtypes = [vp, i32, vp]
        L.ps2d_kinetic_energy.argtypes = [vp, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _check(rc):
    if rc != PS_OK:
        raise PsError(rc, lib().ps_last_error().decode())


def default_params():
    p = Params()
    lib().ps_default_params(C.byref(p))
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def box_sdf(nx, ny, nz, radius=0.25):
    """SDF data of an nx x ny x nz lattice box of touching particles (x slowest, z fastest — np.meshgrid(..., indexing="ij") order), the way
    the reference CPU app's builders write it down (cpu/src/simulation.cpp:666-672): outward normal of the nearest face — the
    normalised sum of the nearest faces' normals on edges and corners — and depth = (layer + 1/2) * diameter * sqrt(number of such faces)"""
    out = np.zeros((nx * ny * nz, 4), np.float32)
    k = 0
    for a in range(nx):
        for b in range(ny):
            for c in range(nz):
                depth = [(a, (-1, 0, 0)), (nx - 1 - a, (1, 0, 0)), (b, (0, -1, 0)), (ny - 1 - b, (0, 1, 0)), (c, (0, 0, -1)), (nz - 1 - c, (0, 0, 1))]
                lo = min(d for d, _ in depth)
                g = np.sum([nrm for d, nrm in depth if d == lo], axis=0).astype(np.float64)
                faces = sum(1 for d, _ in depth if d == lo)
                if np.linalg.norm(g) < 1e-9:      # opposite faces equally near (a one- or two-particle-thick direction): no preferred side
                    g = np.array([0.0, 1.0, 0.0]); faces = 1
                out[k, :3] = g / np.linalg.norm(g)
                out[k, 3] = (lo + 0.5) * 2 * radius * np.sqrt(faces)
                k += 1
    return out


class Solver:
    """Thin handle on a PsCtx: raw SoA access + whole-step and per-stage entry points (include/psolver.h)."""

    def __init__(self, params=None, max_particles=1 << 20, device=0, _borrowed=None):
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)
            return
        p = params if params is not None else default_params()
        h = C.c_void_p()
        _check(lib().ps_create(device, C.byref(p), max_particles, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and self._owned:
            lib().ps_destroy(self._h)
        self._h = None

    __del__ = close

    # --- state ---
    @property
    def n(self):
        return int(lib().ps_num_particles(self._h))

    @property
    def num_cells(self):
        return int(lib().ps_num_cells(self._h))

    @property
    def params(self):
        p = Params()
        _check(lib().ps_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, p):
        _check(lib().ps_set_params(self._h, C.byref(p)))

    def append(self, pos4, vel4, inv_mass, rest_density, phase):
        pos4, vel4 = _arr(pos4, np.float32).reshape(-1, 4), _arr(vel4, np.float32).reshape(-1, 4)
        n = pos4.shape[0]
        w, ro, ph = _arr(inv_mass, np.float32).reshape(-1), _arr(rest_density, np.float32).reshape(-1), _arr(phase, np.int32).reshape(-1)
        assert vel4.shape[0] == n and w.size == n and ro.size == n and ph.size == n
        _check(lib().ps_append_particles(self._h, _ptr(pos4), _ptr(vel4), _ptr(w), _ptr(ro), _ptr(ph), n))

    def add_distance_constraints(self, idx_pairs, rest):
        idx, rest = _arr(idx_pairs, np.uint32).reshape(-1), _arr(rest, np.float32).reshape(-1)
        assert idx.size == 2 * rest.size
        _check(lib().ps_add_distance_constraints(self._h, _ptr(idx), _ptr(rest), rest.size))

    def add_point_constraints(self, idx, xyz):
        idx, xyz = _arr(idx, np.uint32).reshape(-1), _arr(xyz, np.float32).reshape(-1)
        assert xyz.size == 3 * idx.size
        _check(lib().ps_add_point_constraints(self._h, _ptr(idx), _ptr(xyz), idx.size))

    def distance_constraints(self):
        m = int(lib().ps_num_distance_constraints(self._h))
        idx, rest = np.zeros(2 * m, np.uint32), np.zeros(m, np.float32)
        _check(lib().ps_copy_distance_constraints(self._h, _ptr(idx), _ptr(rest)))
        return idx, rest

    def point_constraints(self):
        k = int(lib().ps_num_point_constraints(self._h))
        idx, xyz = np.zeros(k, np.uint32), np.zeros(3 * k, np.float32)
        _check(lib().ps_copy_point_constraints(self._h, _ptr(idx), _ptr(xyz)))
        return idx, xyz

    # --- stepping ---
    def step(self, dt):
        _check(lib().ps_step(self._h, dt))

    def step_streamed(self, dt, pos_in=None, vel_in=None, pos_out=None, vel_out=None):
        """ps_step_streamed: one step whose inputs / outputs are (pinned) host buffers given as raw addresses (or None); the transfers
        overlap the neighbouring calls' solver work.  io_wait(k) blocks until the outputs of the call made k calls ago have landed."""
        _check(lib().ps_step_streamed(self._h, dt, pos_in, vel_in, pos_out, vel_out))

    # --- the slab-decomposed step over NCCL behind the C ABI (csrc/ps_comm.cu) ---
    @staticmethod
    def comm_unique_id():
        """128 bytes from ncclGetUniqueId: rank 0 calls this and hands the bytes to the other ranks"""
        buf = C.create_string_buffer(128)
        _check(lib().ps_comm_get_unique_id(buf))
        return buf.raw

    def comm_init(self, id_bytes, rank, nranks):
        assert len(id_bytes) == 128
        _check(lib().ps_comm_init(self._h, C.c_char_p(bytes(id_bytes)), rank, nranks))

    def comm_set_slab(self, x_lo, x_hi, drift=0.25, exchange_lambda=True, halo_capacity=1 << 16, migrant_capacity=1 << 16):
        _check(lib().ps_comm_set_slab(self._h, x_lo, x_hi, drift, int(bool(exchange_lambda)), int(halo_capacity), int(migrant_capacity)))

    def comm_set_recut(self, cuts, every, x_min, x_max, bins=4096):
        """all nranks + 1 cut planes and the re-cut schedule (ps_comm_set_recut); same arguments on every rank"""
        a = np.ascontiguousarray(cuts, np.float32)
        _check(lib().ps_comm_set_recut(self._h, _ptr(a), int(every), x_min, x_max, int(bins)))

    def comm_cuts(self, nranks):
        a = np.zeros(nranks + 1, np.float32)
        k = C.c_uint32()
        _check(lib().ps_comm_get_cuts(self._h, _ptr(a), C.byref(k)))
        return a, int(k.value)

    def comm_step(self, dt):
        _check(lib().ps_comm_step(self._h, dt))

    def comm_stats(self):
        out = (C.c_uint64 * 4)()
        _check(lib().ps_comm_stats(self._h, out))
        return {"migrated_out": int(out[0]), "ghosts": int(out[1]), "bytes_sent": int(out[2]), "steps": int(out[3])}

    def comm_allreduce_sum(self, values):
        a = np.ascontiguousarray(values, np.float64).copy()
        _check(lib().ps_comm_allreduce_sum(self._h, _ptr(a), a.size))
        return a

    def comm_destroy(self):
        _check(lib().ps_comm_destroy(self._h))

    def io_begin(self, pos_in=None, vel_in=None):
        _check(lib().ps_io_begin(self._h, pos_in, vel_in))

    def io_prefetch(self, pos_in=None, vel_in=None):
        _check(lib().ps_io_prefetch(self._h, pos_in, vel_in))

    def io_end(self, pos_out=None, vel_out=None):
        _check(lib().ps_io_end(self._h, pos_out, vel_out))

    def io_wait(self, calls_back=0):
        _check(lib().ps_io_wait(self._h, calls_back))

    def sync(self):
        _check(lib().ps_sync(self._h))

    def last_step_ms(self):
        ms = C.c_float()
        _check(lib().ps_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    def timer_start(self):
        _check(lib().ps_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        _check(lib().ps_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def step_profiled(self, dt):
        """One eager step with an event after every stage -> ({stage: ms}, {stage: kernel launches})."""
        ms = np.zeros(NUM_STAGES, np.float32)
        ln = np.zeros(NUM_STAGES, np.uint32)
        _check(lib().ps_step_profiled(self._h, dt, _ptr(ms), _ptr(ln)))
        names = [lib().ps_stage_name(k).decode() for k in range(NUM_STAGES)]
        return dict(zip(names, ms.tolist())), dict(zip(names, ln.tolist()))

    def download_async(self, which, host_ptr, count_elems, offset_elems=0):
        _check(lib().ps_download_async(self._h, which, host_ptr, offset_elems, count_elems))

    def upload_async(self, which, host_ptr, count_elems, offset_elems=0):
        _check(lib().ps_upload_async(self._h, which, host_ptr, offset_elems, count_elems))

    @property
    def launches_per_step(self):
        return int(lib().ps_launches_per_step(self._h))

    def begin_step(self): _check(lib().ps_begin_step(self._h))
    def predict(self, dt): _check(lib().ps_predict(self._h, dt))
    def build_grid(self): _check(lib().ps_build_grid(self._h))
    def solve_contacts(self): _check(lib().ps_solve_contacts(self._h))
    def solve_fluid(self): _check(lib().ps_solve_fluid(self._h))
    def solve_fluid_lambda(self): _check(lib().ps_solve_fluid_lambda(self._h))
    def solve_fluid_delta(self): _check(lib().ps_solve_fluid_delta(self._h))
    def collide_world(self, iteration): _check(lib().ps_collide_world(self._h, iteration))
    def solve_distance(self): _check(lib().ps_solve_distance(self._h))
    def solve_point(self): _check(lib().ps_solve_point(self._h))
    def update_velocity(self, dt): _check(lib().ps_update_velocity(self._h, dt))

    def fluid_stats(self):
        """(mean |rho/rho0 - 1|, max |rho/rho0 - 1|, kinetic energy) of the current state"""
        a, b, k = C.c_double(), C.c_double(), C.c_double()
        _check(lib().ps_fluid_stats(self._h, C.byref(a), C.byref(b), C.byref(k)))
        return a.value, b.value, k.value

    # --- checkpoints ---
    def save(self, path): _check(lib().ps_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path, device=0):
        h = C.c_void_p()
        _check(lib().ps_load(os.fsencode(path), device, C.byref(h)))
        self = cls.__new__(cls)
        self._owned, self._h = True, h
        return self

    # --- not in the reference's GPU solver: 3-D shape matching, XSPH viscosity, vorticity confinement (psolver.h) ---
    def add_rigid_body(self, indices, stiffness=1.0):
        idx = _arr(indices, np.uint32).reshape(-1)
        b = C.c_uint32()
        _check(lib().ps_add_rigid_body(self._h, _ptr(idx), idx.size, stiffness, C.byref(b)))
        return int(b.value)

    @property
    def num_rigid_bodies(self):
        return int(lib().ps_num_rigid_bodies(self._h))

    def solve_shapes(self): _check(lib().ps_solve_shapes(self._h))

    def rigid_body_rotation(self, body):
        q = np.zeros(4, np.float32)
        _check(lib().ps_rigid_body_rotation(self._h, int(body), _ptr(q)))
        return q

    def set_rigid_body_sdf(self, body, sdf4):
        """per member (gx, gy, gz, depth): outward normal in the body's rest frame and depth below the surface; depth < 0 = none"""
        a = _arr(sdf4, np.float32).reshape(-1, 4)
        _check(lib().ps_set_rigid_body_sdf(self._h, int(body), _ptr(a)))

    def set_viscosity(self, xsph_c=0.0, vorticity_eps=0.0): _check(lib().ps_set_viscosity(self._h, xsph_c, vorticity_eps))
    def find_neighbors(self): _check(lib().ps_find_neighbors(self._h))
    def apply_viscosity(self, dt): _check(lib().ps_apply_viscosity(self._h, dt))

    # --- data ---
    def _count(self, which):
        if which in (ARR_CELL_START, ARR_CELL_END):
            return self.num_cells
        if which == ARR_CELL_BEGIN:
            return self.num_cells + 1
        if which == ARR_RANDS:
            return int(self.params.solver_iterations) * 6
        if which == ARR_NEIGHBOR_ROWS:
            return (self.n + 31) // 32 if self.params.neighbor_list_rows else 0
        return self.n * _ARR_WIDTH.get(which, 1)

    def download(self, which, out=None):
        cnt = self._count(which)
        a = out if out is not None else np.empty(cnt, dtype=_ARR_DTYPE[which])
        _check(lib().ps_download(self._h, which, _ptr(a), 0, cnt))
        w = _ARR_WIDTH.get(which, 1)
        return a.reshape(-1, w) if w > 1 else a

    def upload(self, which, data, offset_elems=0):
        a = _arr(data, _ARR_DTYPE[which]).reshape(-1)
        _check(lib().ps_upload(self._h, which, _ptr(a), offset_elems, a.size))

    def device_ptr(self, which):
        return lib().ps_device_ptr(self._h, which)

    @property
    def stream(self):
        return lib().ps_stream(self._h)

    def set_ghost_count(self, g):
        _check(lib().ps_set_ghost_count(self._h, g))

    # --- slab decomposition (device pointers in, counts out; see particlesolver_b200/slab.py) ---
    @property
    def n_owned(self):
        return int(lib().ps_num_owned(self._h))

    def slab_pack_halo(self, x_lo, x_hi, width, left_ptr, right_ptr, capacity):
        c = (C.c_uint32 * 2)()
        _check(lib().ps_slab_pack_halo(self._h, x_lo, x_hi, width, left_ptr, right_ptr, capacity, C.byref(c)))
        return int(c[0]), int(c[1])

    def slab_set_ghosts(self, left_ptr, n_left, right_ptr, n_right):
        _check(lib().ps_slab_set_ghosts(self._h, left_ptr, n_left, right_ptr, n_right))

    def slab_pack_migrants(self, x_lo, x_hi, left_ptr, right_ptr, capacity):
        c = (C.c_uint32 * 2)()
        _check(lib().ps_slab_pack_migrants(self._h, x_lo, x_hi, left_ptr, right_ptr, capacity, C.byref(c)))
        return int(c[0]), int(c[1])

    def slab_append_migrants(self, left_ptr, n_left, right_ptr, n_right):
        _check(lib().ps_slab_append_migrants(self._h, left_ptr, n_left, right_ptr, n_right))

    def slab_x_histogram(self, x_min, x_max, bins):
        h = np.zeros(int(bins), np.uint64)
        _check(lib().ps_slab_x_histogram(self._h, x_min, x_max, int(bins), _ptr(h)))
        return h.astype(np.int64)

    def slab_set_lambda_range(self, x_min, x_max):
        _check(lib().ps_slab_set_lambda_range(self._h, max(x_min, -3.0e38), min(x_max, 3.0e38)))

    def slab_pack_lambda(self, left_ptr, right_ptr, capacity):
        c = (C.c_uint32 * 2)()
        _check(lib().ps_slab_pack_lambda(self._h, left_ptr, right_ptr, capacity, C.byref(c)))
        return int(c[0]), int(c[1])

    def slab_set_ghost_lambda(self, left_ptr, n_left, right_ptr, n_right):
        _check(lib().ps_slab_set_ghost_lambda(self._h, left_ptr, n_left, right_ptr, n_right))

    def slab_set_lambda_sinks(self, left_ptr, right_ptr, capacity):
        """all-FLUID contexts: the lambda pass writes the halo members' lambda straight into these two message buffers (None, None: off)"""
        _check(lib().ps_slab_set_lambda_sinks(self._h, left_ptr, right_ptr, capacity))

    def download_owned(self, which):
        """The owned particles' part of a per-particle array (ghosts follow them)."""
        n = self.n_owned * _ARR_WIDTH.get(which, 1)
        a = np.empty(n, dtype=_ARR_DTYPE[which])
        _check(lib().ps_download(self._h, which, _ptr(a), 0, n))
        w = _ARR_WIDTH.get(which, 1)
        return a.reshape(-1, w) if w > 1 else a


class ParticleSystem:
    """The reference's host class (gpu/src/particlesystem.h:22-52), forwarded to the C++ implementation."""

    def __init__(self, particleRadius=0.25, gridSize=(64, 64, 64), maxParticles=15000, minBounds=(-50, 0, -50),
                 maxBounds=(50, 200, 50), iterations=5, _handle=None):
        L = lib()
        if _handle is None:
            mn, mx = (C.c_int * 3)(*minBounds), (C.c_int * 3)(*maxBounds)
            _handle = L.pshost_create(particleRadius, gridSize[0], gridSize[1], gridSize[2], maxParticles, mn, mx, iterations)
        self._h = C.c_void_p(_handle)
        self._raise()

    @classmethod
    def scene(cls, name, grid=64, max_particles=15000, iterations=5, side=100, seed=1):
        """One of the reference's demo scenes ("1".."9", particleapp.cpp:141-215) or the scaled configs "c2"/"c3"."""
        h = lib().pshost_build_scene(str(name).encode(), grid, max_particles, iterations, side, seed)
        if not h:
            raise ValueError(f"unknown scene {name!r}")
        return cls(_handle=h)

    def _raise(self):
        e = lib().pshost_error(self._h).decode()
        if e and not e.startswith("addParticleMultiple: batch dropped"):
            raise PsError(PS_ERR_CUDA, e)

    def close(self):
        if getattr(self, "_h", None):
            lib().pshost_destroy(self._h)
        self._h = None

    __del__ = close

    @property
    def solver(self):
        return Solver(_borrowed=lib().pshost_ctx(self._h))

    def getNumParticles(self):
        return int(lib().pshost_num_particles(self._h))

    def update(self, deltaTime):
        lib().pshost_update(self._h, deltaTime)
        self._raise()

    def addFluid(self, ll, ur, mass, density, color=(0, 0, 1)):
        lib().pshost_add_fluid(self._h, (C.c_int * 3)(*ll), (C.c_int * 3)(*ur), mass, density); self._raise()

    def addParticleGrid(self, ll, ur, mass, addJitter):
        lib().pshost_add_particle_grid(self._h, (C.c_int * 3)(*ll), (C.c_int * 3)(*ur), mass, int(addJitter)); self._raise()

    def addHorizCloth(self, ll, ur, spacing, dist, mass, holdEdges):
        lib().pshost_add_horiz_cloth(self._h, (C.c_int * 2)(*ll), (C.c_int * 2)(*ur), (C.c_float * 3)(*spacing), (C.c_float * 2)(*dist), mass,
                                     int(holdEdges)); self._raise()

    def addRope(self, start, spacing, dist, numLinks, mass, constrainStart):
        lib().pshost_add_rope(self._h, (C.c_float * 3)(*start), (C.c_float * 3)(*spacing), dist, numLinks, mass, int(constrainStart)); self._raise()

    def addRigidBox(self, ll, ur, mass, sdf=True, stiffness=1.0):
        """not in the reference: a lattice box as one shape-matched body with the box's SDF data; returns the body index"""
        b = lib().pshost_add_rigid_box(self._h, (C.c_int * 3)(*ll), (C.c_int * 3)(*ur), mass, int(bool(sdf)), stiffness)
        self._raise()
        return int(b)

    def addStaticSphere(self, ll, ur, spacing):
        lib().pshost_add_static_sphere(self._h, (C.c_int * 3)(*ll), (C.c_int * 3)(*ur), spacing); self._raise()

    def setParticleToAdd(self, pos, vel, mass):
        lib().pshost_set_particle_to_add(self._h, (C.c_float * 3)(*pos), (C.c_float * 3)(*vel), mass)

    def setFluidToAdd(self, pos, color, mass, density):
        lib().pshost_set_fluid_to_add(self._h, (C.c_float * 3)(*pos), (C.c_float * 3)(*color), mass, density)

    def makePointConstraint(self, index, point):
        lib().pshost_make_point_constraint(self._h, index, (C.c_float * 3)(*point)); self._raise()

    def makeDistanceConstraint(self, index, distance):
        lib().pshost_make_distance_constraint(self._h, index[0], index[1], distance); self._raise()

    def getPositions(self):
        a = np.empty((self.getNumParticles(), 4), np.float32)
        lib().pshost_get_positions(self._h, _ptr(a))
        return a

    def getVelocities(self):
        a = np.empty((self.getNumParticles(), 4), np.float32)
        lib().pshost_get_velocities(self._h, _ptr(a))
        return a


class Simulation2D:
    """The 2-D double-precision path: the reference CPU application's Simulation::tick (cpu/src/simulation.cpp:115-369) on
    the GPU with all its constraint groups; `createRigidBody`, `createFluid`, `createGas`, `createSmokeEmitter`, `tick`,
    `getNumParticles`, `getKineticEnergy` keep the reference's names (simulation.h:66-99); `addParticles`,
    `addDistanceConstraint`, `addFluidConstraint`, `addGasConstraint`, `restoreRigidBody` correspond to its Particle /
    DistanceConstraint / TotalFluidConstraint / GasConstraint / Body constructors."""
    SOLID, FLUID, GAS = 0, 1, 2

    def __init__(self, x_bounds=(-8.0, 8.0), y_bounds=(-8.0, 40.0), gravity=(0.0, -9.8), iterations=3, max_particles=1 << 16, device=0,
                 stabilization_iterations=0):
        p = Params2D()
        lib().ps2d_default_params(C.byref(p))
        p.x_bounds[:], p.y_bounds[:], p.gravity[:] = tuple(x_bounds), tuple(y_bounds), tuple(gravity)
        p.solver_iterations = iterations
        p.stabilization_iterations = stabilization_iterations   # 2 = the reference built with USE_STABILIZATION
        h = C.c_void_p()
        _check(lib().ps2d_create(device, C.byref(p), max_particles, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().ps2d_destroy(self._h)
        self._h = None

    __del__ = close

    def createFluid(self, positions, density, velocities=None, inv_mass=None):
        p = _arr(positions, np.float64).reshape(-1, 2)
        v = _arr(velocities if velocities is not None else np.zeros_like(p), np.float64).reshape(-1, 2)
        w = _arr(inv_mass if inv_mass is not None else np.ones(p.shape[0]), np.float64).reshape(-1)
        _check(lib().ps2d_create_fluid(self._h, _ptr(p), _ptr(v), _ptr(w), p.shape[0], density))

    @staticmethod
    def _opt(a, dtype, shape):
        return None if a is None else _arr(a, dtype).reshape(shape)

    def addParticles(self, positions, velocities=None, inv_mass=None, phase=None, bod=None, s_friction=None, k_friction=None):
        p = _arr(positions, np.float64).reshape(-1, 2)
        n = p.shape[0]
        v = self._opt(velocities, np.float64, (-1, 2))
        w = _arr(inv_mass if inv_mass is not None else np.ones(n), np.float64).reshape(-1)
        ph = _arr(phase if phase is not None else np.zeros(n), np.int32).reshape(-1)
        bd, sf, kf = self._opt(bod, np.int32, -1), self._opt(s_friction, np.float64, -1), self._opt(k_friction, np.float64, -1)
        first = C.c_uint64()
        _check(lib().ps2d_add_particles(self._h, _ptr(p), _ptr(v) if v is not None else None, _ptr(w), _ptr(ph), _ptr(bd) if bd is not None else None,
                                        _ptr(sf) if sf is not None else None, _ptr(kf) if kf is not None else None, n, C.byref(first)))
        return int(first.value)

    def addDistanceConstraint(self, i1, i2, d=-1.0):
        _check(lib().ps2d_add_distance_constraint(self._h, int(i1), int(i2), float(d)))

    def addFluidConstraint(self, indices, density):
        idx = _arr(indices, np.uint32).reshape(-1)
        k = C.c_uint32()
        _check(lib().ps2d_add_fluid_constraint(self._h, _ptr(idx), idx.shape[0], density, C.byref(k)))
        return int(k.value)

    def addGasConstraint(self, indices, density, open=False):
        idx = _arr(indices, np.uint32).reshape(-1)
        k = C.c_uint32()
        _check(lib().ps2d_add_gas_constraint(self._h, _ptr(idx), idx.shape[0], density, 1 if open else 0, C.byref(k)))
        return int(k.value)

    def restoreRigidBody(self, first, count, rs, sdf, inv_mass, center, angle, stiffness=1.0):
        rs, sdf, cen = _arr(rs, np.float64).reshape(-1, 2), _arr(sdf, np.float64).reshape(-1, 3), _arr(center, np.float64).reshape(2)
        assert rs.shape[0] == count and sdf.shape[0] == count
        b = C.c_uint32()
        _check(lib().ps2d_restore_rigid_body(self._h, int(first), int(count), _ptr(rs), _ptr(sdf), inv_mass, _ptr(cen), angle, stiffness, C.byref(b)))
        return int(b.value)

    def createRigidBody(self, positions, sdf, velocities=None, inv_mass=None, s_friction=None, k_friction=None):
        p = _arr(positions, np.float64).reshape(-1, 2)
        n = p.shape[0]
        v = self._opt(velocities, np.float64, (-1, 2))
        w = _arr(inv_mass if inv_mass is not None else np.ones(n), np.float64).reshape(-1)
        sd = _arr(sdf, np.float64).reshape(-1, 3)
        sf, kf = self._opt(s_friction, np.float64, -1), self._opt(k_friction, np.float64, -1)
        b = C.c_uint32()
        _check(lib().ps2d_create_rigid_body(self._h, _ptr(p), _ptr(v) if v is not None else None, _ptr(w), _ptr(sf) if sf is not None else None,
                                            _ptr(kf) if kf is not None else None, _ptr(sd), n, C.byref(b)))
        return int(b.value)

    def createGas(self, positions, density, open=False, velocities=None, inv_mass=None):
        p = _arr(positions, np.float64).reshape(-1, 2)
        v = self._opt(velocities, np.float64, (-1, 2))
        w = _arr(inv_mass if inv_mass is not None else np.ones(p.shape[0]), np.float64).reshape(-1)
        k = C.c_uint32()
        _check(lib().ps2d_create_gas(self._h, _ptr(p), _ptr(v) if v is not None else None, _ptr(w), p.shape[0], density, 1 if open else 0, C.byref(k)))
        return int(k.value)

    def createSmokeEmitter(self, posn, particles_per_sec, gas_index=None, timer=0.0):
        q = _arr(posn, np.float64).reshape(2)
        _check(lib().ps2d_create_smoke_emitter(self._h, _ptr(q), particles_per_sec, 0xFFFFFFFF if gas_index is None else int(gas_index), timer))

    def createFluidEmitter(self, posn, particles_per_sec, fluid_index, timer=0.0, total_timer=0.0):
        q = _arr(posn, np.float64).reshape(2)
        _check(lib().ps2d_create_fluid_emitter(self._h, _ptr(q), particles_per_sec, int(fluid_index), timer, total_timer))

    def setParticleTimers(self, t):
        t = _arr(t, np.float64).reshape(-1)
        assert t.shape[0] == self.getNumParticles()
        _check(lib().ps2d_set_particle_timers(self._h, _ptr(t)))

    def particleTimers(self):
        t = np.empty(self.getNumParticles(), np.float64)
        _check(lib().ps2d_get_particle_timers(self._h, _ptr(t)))
        return t

    def mousePressed(self, x, y):
        """Simulation::mousePressed: every particle's velocity gets an impulse of 7 towards the point"""
        _check(lib().ps2d_mouse_pressed(self._h, float(x), float(y)))

    def setForces(self, f):
        f = _arr(f, np.float64).reshape(-1, 2)
        assert f.shape[0] == self.getNumParticles()
        _check(lib().ps2d_set_forces(self._h, _ptr(f)))

    def bodyState(self, body):
        cen = np.zeros(2, np.float64)
        ang = C.c_double()
        _check(lib().ps2d_body_state(self._h, int(body), _ptr(cen), C.byref(ang)))
        return cen, ang.value

    def getNumBodies(self):
        return int(lib().ps2d_num_bodies(self._h))

    @property
    def num_contact_constraints(self):
        return int(lib().ps2d_last_num_contact_constraints(self._h))

    @property
    def num_levels(self):
        return int(lib().ps2d_last_num_levels(self._h))

    def _geti(self, which):
        a = np.empty(self.getNumParticles(), np.int32)
        _check(lib().ps2d_download(self._h, which, _ptr(a)))
        return a

    def inv_mass(self): return self._get(7, 1)
    def s_friction(self): return self._get(8, 1)
    def k_friction(self): return self._get(9, 1)
    def sdf_dist(self): return self._get(10, 1)
    def rs(self): return self._get(11, 2)
    def sdf_grad(self): return self._get(12, 2)
    def phases(self): return self._geti(13)
    def bods(self): return self._geti(14)
    def groups(self): return self._geti(15)

    def save(self, path): _check(lib().ps2d_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path, device=0):
        h = C.c_void_p()
        _check(lib().ps2d_load(os.fsencode(path), device, C.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        return self

    @classmethod
    def scene(cls, key, max_particles=0, device=0):
        """Simulation::init(type) for the CPU app's key-bound scenes (include/ps_scenes2d.h), jitter included"""
        h = C.c_void_p()
        _check(lib().ps2d_build_scene(str(key).encode(), device, max_particles, C.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        return self

    def forces(self): return self._get(4, 2)
    def tmass(self): return self._get(6, 1)

    def counts(self):
        a = np.empty(self.getNumParticles(), np.uint32)
        _check(lib().ps2d_download(self._h, 5, _ptr(a)))
        return a

    @classmethod
    def from_state(cls, scene, max_particles=None, device=0, stabilization_iterations=0):
        """Builds a simulation from a full restart state: the dict layout written by the reference driver
        (oracle/ref_cpu_driver.cpp dump_scene; tests/golden/ref_cpu_scenes.npz) — particles [px, py, vx, vy, imass, phase,
        bod, sFriction, kFriction(, fx, fy(, t))], bodies, the STANDARD constraint list in order, smoke / fluid emitters, rand() position."""
        P = np.array(scene["particles"], np.float64).reshape(-1, len(scene["particles"][0]) if scene["particles"] else 9)
        n = P.shape[0]
        sim = cls(scene["xbounds"], scene["ybounds"], scene["gravity"], max_particles=max_particles or max(1024, 2 * n), device=device,
                  stabilization_iterations=stabilization_iterations)
        sim.addParticles(P[:, 0:2], P[:, 2:4], P[:, 4], P[:, 5].astype(np.int32), P[:, 6].astype(np.int32), P[:, 7], P[:, 8])
        for b in scene["bodies"]:
            idx = np.array(b["particles"])
            assert np.array_equal(idx, np.arange(idx[0], idx[0] + idx.shape[0])), "body members are a contiguous index range"
            sim.restoreRigidBody(idx[0], idx.shape[0], b["rs"], b["sdf"], b["imass"], b["center"], b["angle"], b["stiffness"])
        for k, c in enumerate(scene["standard"]):
            if c["type"] == "fluid":
                assert sim.addFluidConstraint(c["ps"], c["p0"]) == k
            elif c["type"] == "gas":
                assert sim.addGasConstraint(c["ps"], c["p0"], bool(c["open"])) == k
            elif c["type"] == "distance":
                sim.addDistanceConstraint(c["i1"], c["i2"], c["d"])
            else:
                raise ValueError(c["type"])
        for e in scene.get("smoke_emitters", []):
            sim.createSmokeEmitter(e["posn"], e["rate"], e["standard_index"] if e["standard_index"] >= 0 else None, e.get("timer", 0.0))
        fe = scene.get("fluid_emitters", [])
        for e in (fe if isinstance(fe, list) else []):
            sim.createFluidEmitter(e["posn"], e["rate"], e["standard_index"], e.get("timer", 0.0), e.get("total_timer", 0.0))
        if P.shape[1] > 9:
            sim.setForces(P[:, 9:11])
        if P.shape[1] > 11:
            sim.setParticleTimers(P[:, 11])
        sim.seedRand(1, int(scene["rand_calls"]))
        return sim

    def setStabilizationIterations(self, iterations):
        _check(lib().ps2d_set_stabilization_iterations(self._h, int(iterations)))

    def seedRand(self, seed=1, skip=0):
        _check(lib().ps2d_seed_rand(self._h, seed, skip))

    @property
    def rand_calls(self):
        return int(lib().ps2d_rand_calls(self._h))

    def tick(self, seconds=0.01):
        _check(lib().ps2d_tick(self._h, seconds))

    def getNumParticles(self):
        return int(lib().ps2d_num_particles(self._h))

    def getKineticEnergy(self):
        e = C.c_double()
        _check(lib().ps2d_kinetic_energy(self._h, C.byref(e)))
        return e.value

    @property
    def num_boundary_constraints(self):
        return int(lib().ps2d_last_num_boundary_constraints(self._h))

    @property
    def launches_per_tick(self):
        return int(lib().ps2d_launches_per_tick(self._h))

    def _get(self, which, width):
        a = np.empty((self.getNumParticles(), width) if width > 1 else self.getNumParticles(), np.float64)
        _check(lib().ps2d_download(self._h, which, _ptr(a)))
        return a

    def positions(self): return self._get(0, 2)
    def velocities(self): return self._get(1, 2)
    def estimates(self): return self._get(2, 2)
    def lambdas(self): return self._get(3, 1)
