"""CPU engine for particlesolver_b200.slab.SlabDomain built on the oracle (tests only): the same phase interface as
slab.CtxEngine, records as torch CPU uint8 tensors, so that the decomposition logic — cuts, halo selection, ghost
bookkeeping, migration, neighbour exchange over gloo — is exercised without a GPU."""
import numpy as np
import torch

import oracle_py as orc

HALO_DT = np.dtype([("pos", "<f4", 4), ("w", "<f4"), ("ros", "<f4"), ("phase", "<i4"), ("pad", "<u4")])
MIGR_DT = np.dtype([("pos", "<f4", 4), ("prev", "<f4", 4), ("vel", "<f4", 4), ("w", "<f4"), ("ros", "<f4"), ("phase", "<i4"), ("pad", "<u4")])
assert HALO_DT.itemsize == 32 and MIGR_DT.itemsize == 64


def _to_t(rec):
    return torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8).reshape(-1, rec.dtype.itemsize).copy())


def _from_t(t, dt):
    return np.ascontiguousarray(t.numpy()).reshape(-1).view(dt) if t.shape[0] else np.zeros(0, dt)


class OracleEngine:
    def __init__(self, params, pos, vel, w, phase, ros, iterations=5, rands=None):
        self.p = params
        self.pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 4).copy()
        self.vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 4).copy()
        self.prev = self.pos.copy()
        self.w = np.ascontiguousarray(w, np.float32).copy()
        self.phase = np.ascontiguousarray(phase, np.int32).copy()
        self.ros = np.ascontiguousarray(ros, np.float32).copy()
        self.n_owned = self.pos.shape[0]
        self.iterations = iterations
        self.rands = np.full((iterations, 6), 0.5, np.float32) if rands is None else rands
        self.o = None
        self.lambda_range = (-np.inf, np.inf)

    # ---- records ----
    def empty_records(self, count, record_bytes):
        return torch.empty((int(count), record_bytes), dtype=torch.uint8)

    def _drop_ghosts(self):
        n = self.n_owned
        self.pos, self.w, self.phase, self.ros = self.pos[:n], self.w[:n], self.phase[:n], self.ros[:n]

    def pack_halo(self, x_lo, x_hi, width):
        n = self.n_owned
        x = self.pos[:n, 0]
        out = []
        self._halo_sel = (x < np.float32(x_lo + width), x >= np.float32(x_hi - width))
        for sel in self._halo_sel:
            r = np.zeros(int(sel.sum()), HALO_DT)
            r["pos"], r["w"], r["ros"], r["phase"] = self.pos[:n][sel], self.w[:n][sel], self.ros[:n][sel], self.phase[:n][sel]
            out.append(_to_t(r))
        return tuple(out)

    def set_ghosts(self, from_left, from_right):
        self._drop_ghosts()
        for t in (from_left, from_right):
            r = _from_t(t, HALO_DT)
            self.pos = np.concatenate([self.pos, r["pos"]])
            self.w = np.concatenate([self.w, r["w"]])
            self.ros = np.concatenate([self.ros, r["ros"]])
            self.phase = np.concatenate([self.phase, r["phase"]])

    def pack_migrants(self, x_lo, x_hi):
        self._drop_ghosts()
        x = self.pos[:, 0]
        left, right = x < np.float32(x_lo), x >= np.float32(x_hi)
        out = []
        for sel in (left, right):
            r = np.zeros(int(sel.sum()), MIGR_DT)
            r["pos"], r["prev"], r["vel"] = self.pos[sel], self.prev[sel], self.vel[sel]
            r["w"], r["ros"], r["phase"] = self.w[sel], self.ros[sel], self.phase[sel]
            out.append(_to_t(r))
        keep = ~(left | right)
        self.pos, self.prev, self.vel = self.pos[keep], self.prev[keep], self.vel[keep]
        self.w, self.ros, self.phase = self.w[keep], self.ros[keep], self.phase[keep]
        self.n_owned = self.pos.shape[0]
        return tuple(out)

    def append_migrants(self, from_left, from_right):
        for t in (from_left, from_right):
            r = _from_t(t, MIGR_DT)
            self.pos = np.concatenate([self.pos, r["pos"]])
            self.prev = np.concatenate([self.prev, r["prev"]])
            self.vel = np.concatenate([self.vel, r["vel"]])
            self.w = np.concatenate([self.w, r["w"]])
            self.ros = np.concatenate([self.ros, r["ros"]])
            self.phase = np.concatenate([self.phase, r["phase"]])
        self.n_owned = self.pos.shape[0]

    def set_lambda_range(self, x_min, x_max):
        self.lambda_range = (x_min, x_max)

    def x_histogram(self, x_min, x_max, bins):
        x = self.pos[:self.n_owned, 0].astype(np.float64)
        b = np.clip(np.floor((x - x_min) * (bins / (x_max - x_min))).astype(np.int64), 0, bins - 1)
        return np.bincount(b, minlength=bins).astype(np.int64)

    # ---- stages: the oracle runs on owned + ghosts, ghosts are put back afterwards (they are never moved) ----
    def begin_step(self):
        pass

    def predict(self, dt):
        n = self.n_owned
        o = orc.OracleSystem(self.p, self.pos[:n], self.vel[:n], self.w[:n], self.phase[:n], self.ros[:n], iterations=self.iterations)
        o.predict(dt)
        self.prev = o.prev.copy()
        self.pos = np.concatenate([o.pos, self.pos[n:]])

    def _keep_ghosts(self, before):
        self.pos = self.o.pos.copy()
        self.pos[self.n_owned:] = before[self.n_owned:]
        self.o.pos[:] = self.pos

    def build_grid(self):
        n = self.pos.shape[0]
        vel = np.zeros((n, 4), np.float32); vel[:self.n_owned] = self.vel
        self.o = orc.OracleSystem(self.p, self.pos, vel, self.w, self.phase, self.ros, iterations=self.iterations)
        self.o.prev[:self.n_owned] = self.prev
        self.o.build_grid()

    def solve_contacts(self):
        before = self.pos.copy(); self.o.collide(); self._keep_ghosts(before)

    def solve_fluid(self):
        before = self.pos.copy(); self.o.solve_fluids(); self._keep_ghosts(before)

    # lambda exchange: K6, then the lambdas of the last halo pack's particles (record order) go out and the ghosts' come in, then K7
    def solve_fluid_lambda(self):
        self.o.solve_fluids_stage(1)
        lo, hi = self.lambda_range
        assert lo > hi, "exchange mode: no ghost computes its own lambda"
        slot_of = np.empty(self.o.n, np.int64); slot_of[self.o.index] = np.arange(self.o.n)
        self.o.lam[slot_of[self.n_owned:]] = np.nan  # a ghost lambda that is never received must show up in the result
        self._slot_of = slot_of

    def pack_lambda(self):
        lam_by_particle = self.o.lam[self._slot_of[:self.n_owned]]
        return tuple(torch.from_numpy(np.ascontiguousarray(lam_by_particle[sel]).view(np.uint8).reshape(-1, 4).copy()) for sel in self._halo_sel)

    def set_ghost_lambda(self, from_left, from_right):
        vals = np.concatenate([np.ascontiguousarray(t.numpy()).reshape(-1).view(np.float32) if t.shape[0] else np.zeros(0, np.float32)
                               for t in (from_left, from_right)])
        assert vals.size == self.o.n - self.n_owned
        self.o.lam[self._slot_of[self.n_owned:]] = vals

    def solve_fluid_delta(self):
        before = self.pos.copy(); self.o.solve_fluids_stage(2); self._keep_ghosts(before)

    def collide_world(self, it):
        before = self.pos.copy(); self.o.collide_world(self.rands[it]); self._keep_ghosts(before)

    def update_velocity(self, dt):
        n = self.n_owned
        o = orc.OracleSystem(self.p, self.pos[:n], self.vel[:n], self.w[:n], self.phase[:n], self.ros[:n], iterations=self.iterations)
        o.prev[:] = self.prev
        o.calc_velocity(dt)
        self.vel = o.vel.copy()

    def sync(self):
        pass
