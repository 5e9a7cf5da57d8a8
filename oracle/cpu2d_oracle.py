"""oracle/cpu2d_oracle.py — numpy restatement of one tick of the reference's 2-D CPU solver for all-fluid scenes
(config C1: CPU scene 6, two-fluid Rayleigh-Taylor).

TEST INFRASTRUCTURE ONLY: the parity checker of the CUDA 2-D path (particlesolver_b200 ps2d_*).  Nothing under
particlesolver_b200/ may import it.

Parity pin: tests/golden/ref_cpu_scene6.npz — states dumped by the reference's own UNMODIFIED CPU solver
(oracle/_ref/ref_cpu, built from /root/reference/cpu/src by oracle/Makefile; generator tests/golden/make_cpu_golden.py).
tests/test_cpu2d_oracle.py requires this file to reproduce them to 1e-12 after 1, 2, 3 and 10 ticks.

What is restated (double precision, the reference's operation order; sums over neighbours run sequentially in ascending
particle index, exactly like the reference's O(N^2) loops):
  Simulation::tick                      cpu/src/simulation.cpp:115-369   (ITERATIVE, no stabilization, 3 iterations)
  Particle::guess / confirmGuess        cpu/src/particle.h:56-65
  BoundaryConstraint::project           cpu/src/constraint/boundaryconstraint.cpp:14-93
  TotalFluidConstraint::project & co.   cpu/src/constraint/totalfluidconstraint.cpp:41-158
  frand()                               cpu/src/includes.h:25  (float-typed!)  over glibc rand() = random() TYPE_3
Only what an all-fluid scene exercises: no particle-particle contact constraints are generated when no particle is SOLID
(simulation.cpp:188-196), fluid wall friction is a no-op (sFriction = kFriction = 0, particle.h:50-51).
"""
import math

import numpy as np

PARTICLE_RAD = 0.25          # particle.h:6
EPSILON = 1e-4               # includes.h:34
H, H2, H6, H9 = 2.0, 4.0, 64.0, 512.0   # totalfluidconstraint.h:16-19
RELAXATION, K_P, E_P, DQ_P = 0.01, 0.1, 4, 0.2   # totalfluidconstraint.h:22-27
SOLVER_ITERATIONS = 3        # simulation.h:11
RAND_MAX = 2147483647


class GlibcRand:
    """glibc rand()/random(), TYPE_3 additive feedback generator (r[i] = r[i-31] + r[i-3]), as seeded by srand(seed)."""

    def __init__(self, seed=1, skip=0):
        r = [0] * 344
        r[0] = seed if seed else 1
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            r[i] = w + 2147483647 if w < 0 else w
        for i in range(31, 34):
            r[i] = r[i - 31]
        for i in range(34, 344):
            r[i] = (r[i - 31] + r[i - 3]) & 0xFFFFFFFF
        self.r = r[-31:]  # the last 31 words are the whole state
        self.calls = 0
        self.take(skip)

    def take(self, count):
        """the next `count` outputs of rand() as a uint32 array"""
        out = np.empty(count, np.int64)
        r = self.r
        for k in range(count):
            v = (r[-31] + r[-3]) & 0xFFFFFFFF
            r.append(v)
            del r[0]
            out[k] = v >> 1
        self.calls += count
        return out


def frand(raw):
    """includes.h:25 — `inline float frand() { return (double)rand() / (double)RAND_MAX; }`: the quotient is rounded to float"""
    return (raw.astype(np.float64) / float(RAND_MAX)).astype(np.float32).astype(np.float64)


def _poly6(r2):
    term2 = H2 - r2
    return np.where(r2 >= H2, 0.0, (315.0 / (64.0 * math.pi * H9)) * (term2 * term2 * term2))


def _seq_sum(m):
    """row sums accumulated strictly left to right (np.cumsum is sequential; adding an exact 0.0 changes nothing)"""
    return np.cumsum(m, axis=1)[:, -1] if m.shape[1] else np.zeros(m.shape[0])


class Cpu2dOracle:
    def __init__(self, p, v, imass, fluid, rho0, xbounds, ybounds, gravity=(0.0, -9.8), rand_skip=0, seed=1):
        self.p = np.array(p, np.float64).reshape(-1, 2).copy()
        self.v = np.array(v, np.float64).reshape(-1, 2).copy()
        self.ep = self.p.copy()
        self.imass = np.array(imass, np.float64).copy()
        self.fluid = np.array(fluid, np.int64).copy()      # index of the TotalFluidConstraint a particle belongs to
        self.rho0 = [float(x) for x in rho0]               # rest density of each fluid
        self.xb, self.yb, self.g = tuple(xbounds), tuple(ybounds), np.array(gravity, np.float64)
        self.rng = GlibcRand(seed, rand_skip)
        self.n = self.p.shape[0]
        self.last_num_boundary = 0

    # ---- BoundaryConstraint::project for every constraint of the list, one solver iteration ----
    def _project_boundaries(self, cons, counts):
        if not len(cons):
            return
        extra = frand(self.rng.take(len(cons))) * 0.003   # drawn before the early-out, one per constraint, in list order
        d = PARTICLE_RAD + extra
        for (i, value, is_x, greater), dk in zip(cons, d):  # independent: a constraint touches one coordinate of one particle
            c = 0 if is_x else 1
            if greater:
                if self.ep[i, c] >= value + PARTICLE_RAD:
                    continue
                self.ep[i, c] = value + dk
            else:
                if self.ep[i, c] <= value - PARTICLE_RAD:
                    continue
                self.ep[i, c] = value - dk
            # friction: sFriction = kFriction = 0 for fluids -> ep -= dpt * min(0 * d / |dpt|, 1) = ep (boundaryconstraint.cpp:79-92)

    # ---- TotalFluidConstraint::project ----
    def _project_fluid(self, f, counts):
        ps = np.nonzero(self.fluid == f)[0]
        p0 = self.rho0[f]
        ep = self.ep
        rx = ep[ps, 0][:, None] - ep[None, :, 0]
        ry = ep[ps, 1][:, None] - ep[None, :, 1]
        r2 = rx * rx + ry * ry
        self_mask = ps[:, None] == np.arange(self.n)[None, :]
        nb = (r2 < H2) & (self.imass[None, :] != 0.0) & ~self_mask          # neighbours j != i (totalfluidconstraint.cpp:55-76)
        rlen = np.sqrt(r2)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / rlen                                                    # glm::normalize = v * (1 / sqrt(dot))
            coef = (45.0 / (math.pi * H6))
            hm = H - rlen
            # spikyGrad = -normalize(r) * (45/(pi H6)) * (H - rlen) * (H - rlen), zero for rlen >= H or rlen == 0 (:129-135)
            ok = nb & (rlen < H) & (rlen != 0.0)
            sgx = np.where(ok, ((-(rx * inv)) * coef) * hm * hm, 0.0)
            sgy = np.where(ok, ((-(ry * inv)) * coef) * hm * hm, 0.0)
        # density: poly6(r2) / imass_j per neighbour, poly6(0) / imass_i for the particle itself, in index order (:66-81)
        terms = np.where(nb, _poly6(r2) / self.imass[None, :], 0.0)
        terms = np.where(self_mask, _poly6(0.0) / self.imass[None, :], terms)
        pi = _seq_sum(terms)
        # denominator: |grad_j|^2 with grad_j = -spikyGrad / p0 for every neighbour, then |sum_j spikyGrad / p0|^2 (:72-84,137-158)
        gx, gy = -sgx / p0, -sgy / p0
        denom = _seq_sum(np.where(nb, gx * gx + gy * gy, 0.0))
        ox, oy = _seq_sum(sgx) / p0, _seq_sum(sgy) / p0
        denom = denom + (ox * ox + oy * oy)
        lam_f = -((pi / p0) - 1.0) / (denom + RELAXATION)
        lam = np.zeros(self.n)           # lambdas[] is a per-constraint QHash: the other fluid's particles read 0 (:106)
        lam[ps] = lam_f
        # deltas (:95-111)
        base = _poly6(DQ_P * DQ_P * H * H)
        corr = -K_P * np.power(_poly6(rlen * rlen) / base, float(E_P))
        s = (lam_f[:, None] + lam[None, :]) + corr
        dx = _seq_sum(np.where(nb, s * sgx, 0.0)) / p0
        dy = _seq_sum(np.where(nb, s * sgy, 0.0)) / p0
        div = (nb.sum(axis=1) + 1).astype(np.float64) + counts[ps]            # neighbours incl. self + boundary count (:113-115)
        self.ep[ps, 0] += dx / div
        self.ep[ps, 1] += dy / div
        return lam_f

    def tick(self, dt):
        n = self.n
        # (1)-(4) forces, prediction (simulation.cpp:139-161; imass == 0 particles do not move, particle.h:56-58)
        self.v = self.v + dt * self.g
        self.ep = np.where((self.imass == 0.0)[:, None], self.p, self.p + dt * self.v)
        # (8) boundary constraints, generated once per tick from the predicted positions, particle order, x before y (:202-224)
        cons, counts = [], np.zeros(n)
        for i in range(n):
            if self.ep[i, 0] < self.xb[0] + PARTICLE_RAD:
                cons.append((i, self.xb[0], True, True))
            elif self.ep[i, 0] > self.xb[1] - PARTICLE_RAD:
                cons.append((i, self.xb[1], True, False))
            if self.ep[i, 1] < self.yb[0] + PARTICLE_RAD:
                cons.append((i, self.yb[0], False, True))
            elif self.ep[i, 1] > self.yb[1] - PARTICLE_RAD:
                cons.append((i, self.yb[1], False, False))
        for c in cons:
            counts[c[0]] += 1           # BoundaryConstraint::updateCounts; TotalFluidConstraint::updateCounts is empty
        self.last_num_boundary = len(cons)
        # (16)-(21) solver iterations: CONTACT group (the boundary constraints), then STANDARD (fluid 0, fluid 1, ...)
        for _ in range(SOLVER_ITERATIONS):
            self._project_boundaries(cons, counts)
            for f in range(len(self.rho0)):
                self._project_fluid(f, counts)
        # (23)-(27) velocities, sleeping (particle.h:60-65)
        self.v = (self.ep - self.p) / dt
        d = self.ep - self.p
        still = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < EPSILON
        self.v[still] = 0.0
        self.p = np.where(still[:, None], self.p, self.ep)

    def kinetic_energy(self):
        """Simulation::getKineticEnergy (simulation.cpp:1293-1303): sum of .5 * |v|^2 / imass over movable particles"""
        m = self.imass != 0.0
        return float(np.cumsum(0.5 * (self.v[m, 0] * self.v[m, 0] + self.v[m, 1] * self.v[m, 1]) / self.imass[m])[-1])
