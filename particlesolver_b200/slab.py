"""Spatial slab decomposition of one particle system over several GPUs (new work — the reference is
single-GPU).

One process per GPU.  Rank r owns the particles with x in [cuts[r], cuts[r+1]) and keeps, behind them in the same
arrays, ghost copies of its neighbours' particles within `halo` of the two faces.  Per step:

    begin      : predict (K1) on the owned particles
    migrate    : owned particles whose predicted x left the slab move to the neighbour with their full state
    5 x        : refresh the ghost halo (positions change every solver iteration, and the reference rebuilds its grid
                 every iteration too), then K2-K8 on owned + ghosts, writing owned particles only
    finish     : velocity update (K11)
    re-cut     : optionally, every k steps right after the predict, the cut planes move to the equal-count quantiles of the
                 particles' x (per-rank histograms, all-reduced); the migration that follows hands the particles over

Only the exchange itself happens here (torch.distributed send/recv between neighbouring ranks: NCCL over NVLink on
the GPU box, gloo in the CPU tests); selection, packing, compaction and unpacking are CUDA kernels behind the C ABI
(ps_slab_* in include/psolver.h).  There is no data-path collective besides that neighbour exchange.

Halo width.  K7 reads lambda_j of every neighbour j of an owned particle i, ghosts included.  Two ways to have them:
  * exchange_lambda=True (default): between K6 and K7 every rank sends the lambdas of the particles it
    put into its halo buffers (4 bytes per record, same order, so no counts travel) and K6 skips the ghosts altogether.
    halo = H + drift: just the neighbours of the owned particles.
  * exchange_lambda=False: ghost lambdas are computed locally.  A ghost within H (+ drift) of the face has its whole
    neighbourhood inside a 2H-wide halo, so its local lambda equals its owner's up to summation order:
    halo = 2H + 2*drift, lambda range = face +- (H + drift).  Twice the ghosts and K6 work for them, one message less.
`drift` bounds how far a particle moves inside the solver iterations of one step (migration runs once per step, right
after the predict, which is where velocities move particles).
"""
import math

import numpy as np

H = 2.0  # PBF support radius, reference gpu/src/cuda/integration_kernel.cuh:25
HALO_RECORD_BYTES = 32
LAMBDA_RECORD_BYTES = 4
MIGRANT_RECORD_BYTES = 64


def uniform_cuts(x_min, x_max, nranks):
    """nranks+1 cut planes: equal-width slabs over [x_min, x_max], outer faces at +-inf."""
    c = [x_min + (x_max - x_min) * r / nranks for r in range(nranks + 1)]
    c[0], c[-1] = -math.inf, math.inf
    return c


def quantile_cuts(x, nranks):
    """Equal-count cut planes from the x coordinates of the initial particle set."""
    q = np.quantile(np.asarray(x, np.float64), [r / nranks for r in range(1, nranks)]) if nranks > 1 else []
    return [-math.inf, *[float(v) for v in q], math.inf]


def balanced_cuts(hist, x_min, x_max, old_cuts, min_width, max_shift):
    """New cut planes from the global histogram of the particles' x (`hist[b]` = count in bin b of [x_min, x_max)): the
    equal-count quantiles, linearly interpolated inside a bin, then limited — a cut moves by at most `max_shift` per
    re-cut (particles change owner by hopping to the NEIGHBOURING rank, one hop per step, so a cut must not jump over a
    slab) and slabs stay at least `min_width` wide (a rank's ghosts all come from its two neighbours: the halo must fit
    inside a slab).  Pure function of its arguments: every rank computes the same planes from the all-reduced histogram."""
    nranks = len(old_cuts) - 1
    if nranks == 1:
        return [-math.inf, math.inf]
    hist = np.asarray(hist, np.float64)
    bins = hist.shape[0]
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    total = cum[-1]
    edges = x_min + (x_max - x_min) * np.arange(bins + 1) / bins
    new = []
    for r in range(1, nranks):
        target = total * r / nranks
        b = int(np.searchsorted(cum, target, side="right")) - 1
        b = min(max(b, 0), bins - 1)
        frac = (target - cum[b]) / hist[b] if hist[b] > 0 else 0.5
        new.append(float(edges[b] + frac * (edges[b + 1] - edges[b])))
    out = [-math.inf]
    for r in range(1, nranks):
        c = new[r - 1]
        old = old_cuts[r]
        if math.isfinite(old):
            c = min(max(c, old - max_shift), old + max_shift)
        if math.isfinite(out[-1]):
            c = max(c, out[-1] + min_width)
        out.append(c)
    out.append(math.inf)
    for r in range(nranks - 1, 1, -1):  # keep the minimum width from the right as well
        if out[r] - out[r - 1] < min_width:
            out[r - 1] = out[r] - min_width
    return out


class CtxEngine:
    """GPU engine: a libpsolver context + torch-owned device buffers for the records (so that torch.distributed can
    send them as they are)."""

    def __init__(self, solver, halo_capacity, migrant_capacity, device=None):
        import torch
        self.torch = torch
        self.sol = solver
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(solver.stream, device=self.device)  # NCCL ops are ordered against the solver's stream
        mk = lambda n, b: torch.empty((max(int(n), 1), b), dtype=torch.uint8, device=self.device)
        self.halo_cap, self.migr_cap = int(halo_capacity), int(migrant_capacity)
        self.halo_send = (mk(halo_capacity, HALO_RECORD_BYTES), mk(halo_capacity, HALO_RECORD_BYTES))
        self.migr_send = (mk(migrant_capacity, MIGRANT_RECORD_BYTES), mk(migrant_capacity, MIGRANT_RECORD_BYTES))
        self.lam_send = (mk(halo_capacity, LAMBDA_RECORD_BYTES), mk(halo_capacity, LAMBDA_RECORD_BYTES))

    # records in, records out (torch uint8 tensors [count, record_bytes] on the engine's device)
    def empty_records(self, count, record_bytes):
        return self.torch.empty((int(count), record_bytes), dtype=self.torch.uint8, device=self.device)

    def pack_halo(self, x_lo, x_hi, width):
        nl, nr = self.sol.slab_pack_halo(x_lo, x_hi, width, self.halo_send[0].data_ptr(), self.halo_send[1].data_ptr(), self.halo_cap)
        return self.halo_send[0][:nl], self.halo_send[1][:nr]

    def set_ghosts(self, from_left, from_right):
        self._keep = (from_left, from_right)  # the unpack kernel is asynchronous: keep the buffers alive
        self.sol.slab_set_ghosts(from_left.data_ptr() if from_left.shape[0] else None, from_left.shape[0],
                                 from_right.data_ptr() if from_right.shape[0] else None, from_right.shape[0])

    def pack_migrants(self, x_lo, x_hi):
        nl, nr = self.sol.slab_pack_migrants(x_lo, x_hi, self.migr_send[0].data_ptr(), self.migr_send[1].data_ptr(), self.migr_cap)
        return self.migr_send[0][:nl], self.migr_send[1][:nr]

    def append_migrants(self, from_left, from_right):
        self._keep_m = (from_left, from_right)
        self.sol.slab_append_migrants(from_left.data_ptr() if from_left.shape[0] else None, from_left.shape[0],
                                      from_right.data_ptr() if from_right.shape[0] else None, from_right.shape[0])

    def set_lambda_range(self, x_min, x_max):
        self.sol.slab_set_lambda_range(x_min, x_max)

    def pack_lambda(self):
        nl, nr = self.sol.slab_pack_lambda(self.lam_send[0].data_ptr(), self.lam_send[1].data_ptr(), self.halo_cap)
        return self.lam_send[0][:nl], self.lam_send[1][:nr]

    def set_ghost_lambda(self, from_left, from_right):
        self._keep_l = (from_left, from_right)
        self.sol.slab_set_ghost_lambda(from_left.data_ptr() if from_left.shape[0] else None, from_left.shape[0],
                                       from_right.data_ptr() if from_right.shape[0] else None, from_right.shape[0])

    def x_histogram(self, x_min, x_max, bins):
        return self.sol.slab_x_histogram(x_min, x_max, bins)

    # stages
    def begin_step(self): self.sol.begin_step()
    def predict(self, dt): self.sol.predict(dt)
    def build_grid(self): self.sol.build_grid()
    def solve_contacts(self): self.sol.solve_contacts()
    def solve_fluid(self): self.sol.solve_fluid()
    def solve_fluid_lambda(self): self.sol.solve_fluid_lambda()
    def solve_fluid_delta(self): self.sol.solve_fluid_delta()
    def collide_world(self, it): self.sol.collide_world(it)
    def update_velocity(self, dt): self.sol.update_velocity(dt)
    def sync(self): self.sol.sync()

    @property
    def n_owned(self): return self.sol.n_owned

    @property
    def iterations(self): return int(self.sol.params.solver_iterations)


class DistComm:
    """Neighbour exchange over torch.distributed (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.eng, self.group = torch, dist, engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.count_device = getattr(engine, "device", torch.device("cpu"))
        self.bytes_sent = 0

    def exchange(self, to_left, to_right, record_bytes, recv_counts=None):
        """Send `to_left` to rank-1 and `to_right` to rank+1; returns (from_left, from_right).  recv_counts = (from the
        left, from the right) when the receiver already knows them (the lambda exchange answers the halo exchange record
        by record): then nothing but the payload travels and the host never waits."""
        torch, dist = self.torch, self.dist
        r, w = self.rank, self.world
        if recv_counts is None:
            mine = torch.tensor([to_left.shape[0], to_right.shape[0]], dtype=torch.int64, device=self.count_device)
            allc = torch.empty((self.world, 2), dtype=torch.int64, device=self.count_device)
            dist.all_gather_into_tensor(allc, mine, group=self.group) if self.count_device.type == "cuda" else \
                dist.all_gather(list(allc.unbind(0)), mine, group=self.group)
            counts = allc.cpu().tolist()
            n_from_left = counts[r - 1][1] if r > 0 else 0
            n_from_right = counts[r + 1][0] if r < w - 1 else 0
        else:
            n_from_left, n_from_right = (int(recv_counts[0]) if r > 0 else 0), (int(recv_counts[1]) if r < w - 1 else 0)
        from_left = self.eng.empty_records(n_from_left, record_bytes)
        from_right = self.eng.empty_records(n_from_right, record_bytes)
        ops = []
        if r > 0 and to_left.shape[0]:
            ops.append(dist.P2POp(dist.isend, to_left, r - 1, self.group))
        if r < w - 1 and to_right.shape[0]:
            ops.append(dist.P2POp(dist.isend, to_right, r + 1, self.group))
        if n_from_left:
            ops.append(dist.P2POp(dist.irecv, from_left, r - 1, self.group))
        if n_from_right:
            ops.append(dist.P2POp(dist.irecv, from_right, r + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.bytes_sent += (to_left.shape[0] * (r > 0) + to_right.shape[0] * (r < w - 1)) * record_bytes
        return from_left, from_right

    def allreduce_sum(self, counts):
        """element-wise sum of an int64 numpy array over the ranks (the x-histograms of a re-cut)"""
        t = self.torch.as_tensor(np.ascontiguousarray(counts, np.int64), device=self.count_device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class SlabDomain:
    """Host logic of one rank's slab.  Backend-agnostic: `engine` is a CtxEngine (GPU) or, in the CPU tests, an engine
    over the oracle with the same methods."""

    def __init__(self, engine, rank, nranks, cuts, drift=0.25, comm=None, recut_every=0, recut_range=None, recut_bins=4096, exchange_lambda=True):
        """exchange_lambda: ghost lambdas come from their owners between K6 and K7 (narrow halo) instead of being computed
        locally (wide halo), see the module docstring.  recut_every > 0: every that many steps the cut planes are moved to the equal-count quantiles of the particles' x
        (a histogram over `recut_range` = (x_min, x_max) with `recut_bins` bins, summed over the ranks), so that slabs keep
        equal particle counts while the fluid flows — SURVEY §8e."""
        assert len(cuts) == nranks + 1 and all(cuts[k] < cuts[k + 1] for k in range(nranks))
        self.eng, self.rank, self.nranks, self.comm = engine, rank, nranks, comm
        self.exchange_lambda = bool(exchange_lambda)
        self.halo = (H + drift) if self.exchange_lambda else (2.0 * H + 2.0 * drift)
        self.lambda_ext = H + drift
        self.ghost_counts = (0, 0)
        self.stats = {"migrated_out": 0, "ghosts": 0, "recuts": 0}
        self.recut_every, self.recut_range, self.recut_bins = int(recut_every), recut_range, int(recut_bins)
        assert not self.recut_every or (recut_range is not None and recut_range[0] < recut_range[1])
        self.steps_done = 0
        self.set_cuts(cuts)

    def set_cuts(self, cuts):
        self.cuts = [float(c) for c in cuts]
        self.x_lo, self.x_hi = self.cuts[self.rank], self.cuts[self.rank + 1]
        if self.exchange_lambda:
            self.eng.set_lambda_range(1.0, -1.0)  # empty: no ghost computes a lambda, they all receive one
        else:
            self.eng.set_lambda_range(self.x_lo - self.lambda_ext, self.x_hi + self.lambda_ext)

    def x_histogram(self):
        return np.asarray(self.eng.x_histogram(self.recut_range[0], self.recut_range[1], self.recut_bins), np.int64)

    def recut_due(self):
        return self.recut_every > 0 and self.nranks > 1 and self.steps_done > 0 and self.steps_done % self.recut_every == 0

    def apply_recut(self, global_hist):
        """new cut planes from the summed histogram; takes effect with the migration that follows"""
        width = min((b - a for a, b in zip(self.cuts[1:-2], self.cuts[2:-1])), default=math.inf)
        max_shift = 0.5 * min(width, 4.0 * self.halo) if math.isfinite(width) else 2.0 * self.halo
        self.set_cuts(balanced_cuts(global_hist, self.recut_range[0], self.recut_range[1], self.cuts, min_width=self.halo + 0.5, max_shift=max_shift))
        self.stats["recuts"] += 1

    # ---- phases (a LocalCluster drives them in lock-step; step() strings them together over a communicator) ----
    def begin(self, dt):
        self.eng.begin_step()
        self.eng.predict(dt)

    def pack_migrants(self):
        l, r = self.eng.pack_migrants(self.x_lo, self.x_hi)
        self.stats["migrated_out"] += int(l.shape[0]) + int(r.shape[0])
        return l, r

    def apply_migrants(self, from_left, from_right):
        self.eng.append_migrants(from_left, from_right)

    def pack_halo(self):
        return self.eng.pack_halo(self.x_lo, self.x_hi, self.halo)

    def apply_halo(self, from_left, from_right):
        self.ghost_counts = (int(from_left.shape[0]), int(from_right.shape[0]))
        self.stats["ghosts"] = sum(self.ghost_counts)
        self.eng.set_ghosts(from_left, from_right)

    def solve(self, it):
        """one solver iteration without a lambda exchange (ghost lambdas computed locally)"""
        self.solve_first_half()
        self.solve_second_half(it)

    def solve_first_half(self):
        e = self.eng
        e.build_grid()
        e.solve_contacts()
        if self.exchange_lambda:
            e.solve_fluid_lambda()  # K6; K7 follows the exchange

    def pack_lambda(self):
        return self.eng.pack_lambda()

    def apply_lambda(self, from_left, from_right):
        self.eng.set_ghost_lambda(from_left, from_right)

    def solve_second_half(self, it):
        e = self.eng
        if self.exchange_lambda:
            e.solve_fluid_delta()
        else:
            e.solve_fluid()
        e.collide_world(it)

    def finish(self, dt):
        self.eng.update_velocity(dt)

    # ---- one whole step over the communicator ----
    def _on_stream(self):
        s = getattr(self.eng, "stream", None)
        if s is None:
            import contextlib
            return contextlib.nullcontext()
        return self.eng.torch.cuda.stream(s)

    def migrate(self):
        l, r = self.pack_migrants()
        with self._on_stream():
            fl, fr = self.comm.exchange(l, r, MIGRANT_RECORD_BYTES)
        self.apply_migrants(fl, fr)

    def refresh_halo(self):
        l, r = self.pack_halo()
        with self._on_stream():
            fl, fr = self.comm.exchange(l, r, HALO_RECORD_BYTES)
        self.apply_halo(fl, fr)

    def exchange_ghost_lambda(self):
        l, r = self.pack_lambda()
        with self._on_stream():
            fl, fr = self.comm.exchange(l, r, LAMBDA_RECORD_BYTES, recv_counts=self.ghost_counts)
        self.apply_lambda(fl, fr)

    def step(self, dt):
        self.begin(dt)
        if self.recut_due():
            self.apply_recut(self.comm.allreduce_sum(self.x_histogram()))
        self.migrate()
        for it in range(self.eng.iterations):
            self.refresh_halo()
            self.solve_first_half()
            if self.exchange_lambda:
                self.exchange_ghost_lambda()
            self.solve_second_half(it)
        self.finish(dt)
        self.steps_done += 1


class LocalCluster:
    """All slabs in one process, stepped in lock-step with in-process hand-over of the record buffers (tests; also a way
    to run several slabs on one GPU)."""

    def __init__(self, engines, cuts, drift=0.25, **options):
        n = len(engines)
        self.doms = [SlabDomain(e, r, n, cuts, drift, **options) for r, e in enumerate(engines)]

    def _hand_over(self, sends, apply, record_bytes):
        n = len(self.doms)
        for r, d in enumerate(self.doms):
            fl = sends[r - 1][1] if r > 0 else d.eng.empty_records(0, record_bytes)
            fr = sends[r + 1][0] if r < n - 1 else d.eng.empty_records(0, record_bytes)
            apply(d, fl, fr)
        for d in self.doms:
            d.eng.sync()  # the send buffers are reused by the next pack

    def step(self, dt):
        for d in self.doms:
            d.begin(dt)
        if self.doms[0].recut_due():
            total = sum(d.x_histogram() for d in self.doms)
            for d in self.doms:
                d.apply_recut(total)
        self._hand_over([d.pack_migrants() for d in self.doms], lambda d, a, b: d.apply_migrants(a, b), MIGRANT_RECORD_BYTES)
        for it in range(self.doms[0].eng.iterations):
            self._hand_over([d.pack_halo() for d in self.doms], lambda d, a, b: d.apply_halo(a, b), HALO_RECORD_BYTES)
            for d in self.doms:
                d.solve_first_half()
            if self.doms[0].exchange_lambda:
                sends = [d.pack_lambda() for d in self.doms]
                for d in self.doms:
                    d.eng.sync()  # every context has its own stream; the lambda pack is asynchronous (no counts to wait for)
                self._hand_over(sends, lambda d, a, b: d.apply_lambda(a, b), LAMBDA_RECORD_BYTES)
            for d in self.doms:
                d.solve_second_half(it)
        for d in self.doms:
            d.finish(dt)
            d.steps_done += 1


# ---------------------------------------------------------------- synthetic dam break ----------------------------------------------------------------
def _hash_uniform(idx, stream, seed=1234):
    """Counter-based uniforms in [0,1): a splitmix64 finaliser of (global lattice index, stream, seed), so that any rank
    count generates the identical particle set."""
    z = (idx.astype(np.uint64) * np.uint64(3) + np.uint64(stream)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
    z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return (z >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def dam_break_block(nx, ny, nz, ix0=0, ix1=None, origin=(0.3125, 0.3125, 0.3125), spacing=0.625, jitter=0.0025, rest_density=4.1):
    """Particles of lattice columns ix0 <= ix < ix1 of an nx x ny x nz fluid block (x fastest, then y, then z):
    spacing 2.5 r, jitter +-0.01 r from a counter-based hash of the GLOBAL lattice index, mass 1, rho0 = 4.1 (the
    lattice's own density, so the block starts near equilibrium).  Returns pos4, vel4, inv_mass, rest_density, phase."""
    ix1 = nx if ix1 is None else ix1
    ix = np.arange(ix0, ix1, dtype=np.int64)
    z, y, x = np.meshgrid(np.arange(nz, dtype=np.int64), np.arange(ny, dtype=np.int64), ix, indexing="ij")
    gid = ((z * ny + y) * nx + x).ravel()
    n = gid.size
    pos = np.ones((n, 4), np.float32)
    for c, lat in enumerate((x.ravel(), y.ravel(), z.ravel())):
        j = (_hash_uniform(gid, c) * 2.0 - 1.0) * jitter
        pos[:, c] = (origin[c] + lat * spacing + j).astype(np.float32)
    return pos, np.zeros((n, 4), np.float32), np.ones(n, np.float32), np.full(n, rest_density, np.float32), np.zeros(n, np.int32)
