"""ctypes binding of oracle/libpsoracle.so — the CPU restatement of the reference's GPU step (oracle/gpu_step_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke().  Never imported by the particlesolver_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libpsoracle.so")


class OrParams(C.Structure):
    _fields_ = [("gravity", C.c_float * 3), ("radius", C.c_float), ("grid", C.c_uint32 * 3), ("origin", C.c_float * 3),
                ("cell", C.c_float * 3), ("min_b", C.c_int * 3), ("max_b", C.c_int * 3)]


_lib = None


def build():
    src = os.path.join(ORACLE_DIR, "gpu_step_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", ORACLE_DIR, "libpsoracle.so"], check=True, env=env, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        P = C.POINTER(OrParams)
        L.or_predict.argtypes = [vp, vp, vp, u32, f32, vp]
        L.or_calc_hash.argtypes = [vp, u32, P, vp, vp]
        L.or_sort.argtypes = [vp, vp, u32]
        L.or_reorder.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp, vp, vp, vp, vp]
        L.or_collide.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u32, P, vp]
        L.or_collide_ext.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u32, P, vp, vp, vp, vp]
        L.or_collide_ext.restype = None
        L.or_solve_fluids.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, P, vp, vp, vp]
        L.or_solve_fluids_stages.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, P, vp, vp, vp, C.c_int]
        L.or_solve_fluids_stages.restype = None
        L.or_collide_world.argtypes = [vp, vp, vp, u32, vp, P]
        L.or_solve_distance.argtypes = [vp, vp, vp, u32, vp, u32]
        L.or_solve_distance.restype = C.c_int
        L.or_solve_point.argtypes = [vp, vp, vp, u32]
        L.or_calc_velocity.argtypes = [vp, vp, vp, u32, f32]
        L.or_occurrences.argtypes = [vp, u32, vp, u32, vp, u32]
        L.or_state_new.argtypes = [u32, u32]
        L.or_state_new.restype = vp
        L.or_state_free.argtypes = [vp]
        L.or_fluid_stats.argtypes = [vp, vp, vp, vp, vp, vp, P, vp]
        L.or_set_omega.argtypes = [f32]
        L.or_set_omega.restype = None
        for f in ("or_state_hash", "or_state_index", "or_state_cell_start", "or_state_cell_end", "or_state_num_neighbors",
                  "or_state_lambda", "or_state_sorted_pos"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        for f in ("or_predict", "or_calc_hash", "or_sort", "or_reorder", "or_collide", "or_solve_fluids", "or_collide_world",
                  "or_solve_point", "or_calc_velocity", "or_occurrences", "or_state_free", "or_fluid_stats"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_params(radius=0.25, grid=(64, 64, 64), min_b=(-50, 0, -50), max_b=(50, 200, 50), gravity=(0.0, -9.8, 0.0), origin=(0, 0, 0),
                cell=None):
    p = OrParams()
    p.gravity[:] = gravity
    p.radius = radius
    p.grid[:] = grid
    p.origin[:] = origin
    c = cell if cell is not None else (np.float32(radius) * np.float32(2.0),) * 3
    p.cell[:] = [float(x) for x in c]
    p.min_b[:] = min_b
    p.max_b[:] = max_b
    return p


class OracleSystem:
    """Host-side state of one particle system, stepped stage by stage like ParticleSystem::update
    (reference gpu/src/particlesystem.cpp:144-246)."""

    def __init__(self, params, pos, vel, w, phase, ros, dist_idx=(), dist_rest=(), point_idx=(), point_xyz=(), iterations=5):
        self.p = params
        self.pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 4).copy()
        self.n = self.pos.shape[0]
        self.vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 4).copy()
        self.prev = np.zeros_like(self.pos)
        self.w = np.ascontiguousarray(w, np.float32).copy()
        self.phase = np.ascontiguousarray(phase, np.int32).copy()
        self.ros = np.ascontiguousarray(ros, np.float32).copy()
        self.dist_idx = np.ascontiguousarray(dist_idx, np.uint32).reshape(-1).copy()
        self.dist_rest = np.ascontiguousarray(dist_rest, np.float32).reshape(-1).copy()
        self.point_idx = np.ascontiguousarray(point_idx, np.uint32).reshape(-1).copy()
        self.point_xyz = np.ascontiguousarray(point_xyz, np.float32).reshape(-1).copy()
        self.iterations = iterations
        self.num_cells = int(params.grid[0]) * int(params.grid[1]) * int(params.grid[2])
        n, nc = self.n, self.num_cells
        self.hash = np.zeros(n, np.uint32)
        self.index = np.zeros(n, np.uint32)
        self.cell_start = np.zeros(nc, np.uint32)
        self.cell_end = np.zeros(nc, np.uint32)
        self.spos = np.zeros((n, 4), np.float32)
        self.sw = np.zeros(n, np.float32)
        self.sphase = np.zeros(n, np.int32)
        self.lam = np.zeros(n, np.float32)
        self.nn = np.zeros(n, np.uint32)
        self.occ = np.zeros(n, np.uint32)
        lib().or_occurrences(_p(self.occ), n, _p(self.dist_idx), self.dist_rest.size, _p(self.point_idx), self.point_idx.size)
        self.dist_nonprefix = 0
        self.omega = 1.0             # SOR factor on the averaged deltas (PsParams.omega); 1 == the reference
        self.self_collision = False  # True: the opt-in rule of PS_FLAG_SELF_COLLISION (not the reference's)
        self._adj = None
        self.sdf_world = None        # (n, 4) float32 by particle index: world-frame SDF of rigid-body particles (ps_set_rigid_body_sdf), depth < 0 = none

    def predict(self, dt):
        g = np.array(list(self.p.gravity), np.float32)
        lib().or_predict(_p(self.pos), _p(self.vel), _p(self.prev), self.n, min(dt, 0.05), _p(g))

    def calc_hash(self):
        lib().or_calc_hash(_p(self.pos), self.n, C.byref(self.p), _p(self.hash), _p(self.index))

    def sort(self):
        lib().or_sort(_p(self.hash), _p(self.index), self.n)

    def reorder(self):
        lib().or_reorder(_p(self.hash), _p(self.index), _p(self.pos), _p(self.w), _p(self.phase), self.n, self.num_cells,
                         _p(self.cell_start), _p(self.cell_end), _p(self.spos), _p(self.sw), _p(self.sphase))

    def build_grid(self):
        self.calc_hash(); self.sort(); self.reorder()

    def _adjacency(self):
        """CSR of the distance constraints by particle index (offsets n + 1, partners 2m)"""
        if self._adj is None:
            pairs = self.dist_idx.reshape(-1, 2).astype(np.int64)
            a = np.concatenate([pairs[:, 0], pairs[:, 1]]); b = np.concatenate([pairs[:, 1], pairs[:, 0]])
            order = np.argsort(a, kind="stable")
            off = np.zeros(self.n + 1, np.uint32)
            off[1:] = np.cumsum(np.bincount(a, minlength=self.n))
            self._adj = (off, np.ascontiguousarray(b[order], np.uint32))
        return self._adj

    def collide(self):
        lib().or_set_omega(self.omega)
        use_adj = self.self_collision and self.dist_rest.size
        if use_adj or self.sdf_world is not None:
            off, adj = self._adjacency() if use_adj else (None, None)
            sdf = np.ascontiguousarray(self.sdf_world, np.float32) if self.sdf_world is not None else None
            lib().or_collide_ext(_p(self.pos), _p(self.prev), _p(self.spos), _p(self.sw), _p(self.sphase), _p(self.index), _p(self.cell_start),
                                 _p(self.cell_end), self.n, C.byref(self.p), _p(self.nn), _p(off) if use_adj else None, _p(adj) if use_adj else None,
                                 _p(sdf) if sdf is not None else None)
            return
        lib().or_collide(_p(self.pos), _p(self.prev), _p(self.spos), _p(self.sw), _p(self.sphase), _p(self.index), _p(self.cell_start),
                         _p(self.cell_end), self.n, C.byref(self.p), _p(self.nn))

    def solve_fluids(self):
        lib().or_set_omega(self.omega)
        lib().or_solve_fluids(_p(self.spos), _p(self.sw), _p(self.sphase), _p(self.index), _p(self.cell_start), _p(self.cell_end),
                              _p(self.pos), self.n, C.byref(self.p), _p(self.ros), _p(self.lam), _p(self.nn))

    def solve_fluids_stage(self, stages):
        """1: lambdas only, 2: delta p only (from the lambdas in self.lam), 3: both"""
        lib().or_set_omega(self.omega)
        lib().or_solve_fluids_stages(_p(self.spos), _p(self.sw), _p(self.sphase), _p(self.index), _p(self.cell_start), _p(self.cell_end),
                                     _p(self.pos), self.n, C.byref(self.p), _p(self.ros), _p(self.lam), _p(self.nn), int(stages))

    def collide_world(self, rands6):
        r = np.ascontiguousarray(rands6, np.float32)
        lib().or_collide_world(_p(self.pos), _p(self.prev), _p(self.phase), self.n, _p(r), C.byref(self.p))

    def solve_distance(self):
        lib().or_set_omega(self.omega)
        self.dist_nonprefix |= lib().or_solve_distance(_p(self.pos), _p(self.dist_idx), _p(self.dist_rest), self.dist_rest.size, _p(self.occ), self.n)

    def solve_point(self):
        lib().or_solve_point(_p(self.pos), _p(self.point_idx), _p(self.point_xyz), self.point_idx.size)

    def calc_velocity(self, dt):
        lib().or_calc_velocity(_p(self.pos), _p(self.prev), _p(self.vel), self.n, min(dt, 0.05))

    def step(self, dt, rands):
        """rands: (iterations, 6) uniforms, the ones cuRAND handed the GPU implementation."""
        rands = np.ascontiguousarray(rands, np.float32).reshape(-1, 6)
        self.predict(dt)
        for it in range(self.iterations):
            self.build_grid()
            self.collide()
            self.solve_fluids()
            self.collide_world(rands[it])
            self.solve_distance()
            self.solve_point()
        self.calc_velocity(dt)

    def fluid_stats(self):
        """(mean |rho/rho0-1|, max |rho/rho0-1|, kinetic energy) on the current state."""
        L = lib()
        st = L.or_state_new(self.n, self.num_cells)
        out = np.zeros(3, np.float64)
        L.or_fluid_stats(st, _p(self.pos), _p(self.vel), _p(self.w), _p(self.phase), _p(self.ros), C.byref(self.p), _p(out))
        L.or_state_free(st)
        return tuple(out)
