"""oracle/cpu2d_full_oracle.py — plain-Python restatement of one tick of the reference's 2-D CPU solver with EVERY
constraint group it has: contact, rigid (SDF) contact with friction, wall constraints with friction
and jitter, distance, shape matching, fluid density (with solid coupling) and gas (buoyancy, open-boundary drag,
pseudo-vorticity force), plus the smoke emitter's particle injection.

TEST INFRASTRUCTURE ONLY: the parity checker of the general CUDA 2-D path (ps2d_* in include/psolver2d.h).  Nothing under
particlesolver_b200/ may import it.

Parity pin: tests/golden/ref_cpu_scenes.npz — scene descriptions and states dumped by the reference's own UNMODIFIED CPU
solver (oracle/_ref/ref_cpu, built from /root/reference/cpu/src by oracle/Makefile; generator
tests/golden/make_cpu_scenes_golden.py) for all its key-bound scenes.  tests/test_cpu2d_full_oracle.py requires this
file to reproduce them.

Sequential on purpose: the reference is Gauss-Seidel over its constraint lists, and the operation order below is the
reference's, expression by expression (double precision, glm's component-wise vector arithmetic, the same libm).
  Simulation::tick                     cpu/src/simulation.cpp:115-369   (ITERATIVE, USE_STABILIZATION off, 3 iterations)
  Particle                             cpu/src/particle.h:21-79
  Body::updateCOM / computeRs          cpu/src/solver/particle.cpp:15-73
  BoundaryConstraint::project          cpu/src/constraint/boundaryconstraint.cpp:14-93
  ContactConstraint::project           cpu/src/constraint/contactconstraint.cpp:13-42
  RigidContactConstraint::project      cpu/src/constraint/rigidcontactconstraint.cpp:13-96
  DistanceConstraint::project          cpu/src/constraint/distanceconstraint.cpp:20-40
  TotalShapeConstraint::project/guess  cpu/src/constraint/totalshapeconstraint.cpp:14-24,80-87
  TotalFluidConstraint::project & co.  cpu/src/constraint/totalfluidconstraint.cpp:41-158
  GasConstraint::project & co.         cpu/src/constraint/gasconstraint.cpp:30-116,...
  OpenSmokeEmitter::tick               cpu/src/opensmokeemitter.cpp:17-29  (particle injection only; the emitter's own
                                       tracer particles are display-only and never read by the solver)
  FluidEmitter::tick                   cpu/src/fluidemitter.cpp:13-79  (VOLCANO scene: emission, freezing into solids)
The stabilization pass (the reference's compile-time option USE_STABILIZATION, off in its build) is restated behind
`stabilization_iterations` and pinned by tests/golden/ref_cpu_scenes_stab.npz (oracle/_ref/ref_cpu_stab).  Not restated: the matrix
solver (dead under ITERATIVE).
"""
import math

import numpy as np

from cpu2d_oracle import GlibcRand, RAND_MAX

PARTICLE_RAD, PARTICLE_DIAM = .25, .5      # particle.h:6-7
EPSILON = .0001                            # includes.h:34
SOLID, FLUID, GAS = 0, 1, 2                # particle.h:10-15
SOLVER_ITERATIONS = 3                      # simulation.h:11
ALPHA = -.2                                # simulation.h:21
H, H2, H6, H9 = 2., 4., 64., 512.
RELAXATION = .01
FLUID_C = dict(K_P=.1, E_P=4, DQ_P=.2, S_SOLID=0.)    # totalfluidconstraint.h:22-30
GAS_C = dict(K_P=.2, E_P=4, DQ_P=.25, S_SOLID=.5)     # gasconstraint.h:12-20
M_PI = math.pi


def poly6(r2):
    if r2 >= H2:
        return 0.
    t = H2 - r2
    return (315. / (64. * M_PI * H9)) * (t * t * t)


def spiky_grad(rx, ry, rlen2):
    """-normalize(r) * (45 / (pi H^6)) * (H - x) * (H - x); the argument named rlen2 is a LENGTH at every call site
    but one (the gas vorticity term passes dot(r, r), gasconstraint.cpp:100)."""
    if rlen2 >= H or rlen2 == 0:
        return 0., 0.
    inv = 1. / math.sqrt(rx * rx + ry * ry)
    c = 45. / (M_PI * H6)
    hm = H - rlen2
    return ((-(rx * inv)) * c) * hm * hm, ((-(ry * inv)) * c) * hm * hm


class Cpu2dFullOracle:
    def __init__(self, scene, rand_seed=1, stabilization_iterations=0):
        """scene: the dict of scene.json written by oracle/_ref/ref_cpu --dump (see oracle/ref_cpu_driver.cpp);
        stabilization_iterations: 0 = the reference as built, 2 = built with USE_STABILIZATION (STABILIZATION_ITERATIONS, simulation.h:18)"""
        self.stabilization_iterations = int(stabilization_iterations)
        P = scene["particles"]
        self.p = [[q[0], q[1]] for q in P]
        self.v = [[q[2], q[3]] for q in P]
        self.ep = [[0., 0.] for _ in P]
        self.f = [[q[9], q[10]] if len(q) > 9 else [0., 0.] for q in P]
        self.imass = [q[4] for q in P]
        self.tmass = list(self.imass)
        self.ph = [int(q[5]) for q in P]
        self.bod = [int(q[6]) for q in P]
        self.sfric = [q[7] for q in P]
        self.kfric = [q[8] for q in P]
        self.t = [q[11] if len(q) > 11 else 4. for q in P]   # Particle::t (particle.h:38), only read by the FluidEmitter
        self.xb, self.yb, self.gravity = list(scene["xbounds"]), list(scene["ybounds"]), list(scene["gravity"])
        self.bodies = []
        for b in scene["bodies"]:
            self.bodies.append(dict(particles=list(b["particles"]), rs={i: tuple(r) for i, r in zip(b["particles"], b["rs"])},
                                    sdf={i: tuple(s) for i, s in zip(b["particles"], b["sdf"])}, imass=b["imass"], center=list(b["center"]),
                                    angle=b["angle"], stiffness=b["stiffness"]))
        self.standard = []
        for c in scene["standard"]:
            c = dict(c)
            if c["type"] in ("fluid", "gas"):
                c["ps"] = list(c["ps"])
            self.standard.append(c)
        self.emitters = [dict(posn=list(e["posn"]), rate=e["rate"], timer=e.get("timer", 0.), gas=(self.standard[e["standard_index"]] if e["standard_index"] >= 0 else None))
                         for e in scene.get("smoke_emitters", [])]
        fe = scene.get("fluid_emitters", [])
        self.fluid_emitters = [dict(posn=list(e["posn"]), rate=e["rate"], timer=e.get("timer", 0.), total_timer=e.get("total_timer", 0.),
                                    fluid=self.standard[e["standard_index"]]) for e in (fe if isinstance(fe, list) else [])]
        self.rng = GlibcRand(rand_seed, int(scene["rand_calls"]))
        self.num_contacts = 0

    # ------------------------------------------------------------------ helpers
    @property
    def n(self):
        return len(self.p)

    def frand(self):
        return float(np.float32(float(self.rng.take(1)[0]) / float(RAND_MAX)))  # includes.h:25: float-typed

    def sdf_data(self, i):
        """Particle::getSDFData (solver/particle.cpp:3-13): the body's SDF sample rotated by the body angle"""
        if self.ph[i] != SOLID or self.bod[i] < 0:
            return 0., 0., -1.
        b = self.bodies[self.bod[i]]
        gx, gy, d = b["sdf"][i]
        c, s = math.cos(b["angle"]), math.sin(b["angle"])
        return gx * c - gy * s, gx * s + gy * c, d

    # ------------------------------------------------------------------ constraints
    def project_boundary(self, c, counts, stabile=False):
        """stabile: the STABILIZATION copy of the constraint (boundaryconstraint.cpp:32-35,74-76): moves p along with ep, no friction;
        the jitter draw and the validity test (on ep) are the same"""
        i, value, is_x, greater = c[1], c[2], c[3], c[4]
        ep, p = self.ep[i], self.p[i]
        extra = self.frand() * .003 if self.ph[i] in (FLUID, GAS) else 0
        d = PARTICLE_RAD + extra
        if greater:
            if is_x:
                if ep[0] >= value + PARTICLE_RAD:
                    return
                ep[0] = value + d
                if stabile:
                    p[0] = value + d
                n = (1., 0.)
            else:
                if ep[1] >= value + PARTICLE_RAD:
                    return
                ep[1] = value + d
                if stabile:
                    p[1] = value + d
                n = (0., 1.)
        else:
            if is_x:
                if ep[0] <= value - PARTICLE_RAD:
                    return
                ep[0] = value - d
                if stabile:
                    p[0] = value - d
                n = (-1., 0.)
            else:
                if ep[1] <= value - PARTICLE_RAD:
                    return
                ep[1] = value - d
                if stabile:
                    p[1] = value - d
                n = (0., -1.)
        if stabile:
            return
        cnt = float(counts[i])
        dpx, dpy = (ep[0] - p[0]) / cnt, (ep[1] - p[1]) / cnt
        dn = dpx * n[0] + dpy * n[1]
        tx, ty = dpx - dn * n[0], dpy - dn * n[1]
        ldpt = math.sqrt(tx * tx + ty * ty)
        if ldpt < EPSILON:
            return
        if ldpt < math.sqrt(self.sfric[i]) * d:
            ep[0] -= tx
            ep[1] -= ty
        else:
            m = min(math.sqrt(self.kfric[i]) * d / ldpt, 1.)
            ep[0] -= tx * m
            ep[1] -= ty * m

    def project_contact(self, c, counts):
        i1, i2 = c[1], c[2]
        if self.tmass[i1] == 0. and self.tmass[i2] == 0.:
            return
        e1, e2 = self.ep[i1], self.ep[i2]
        dx, dy = e1[0] - e2[0], e1[1] - e2[1]
        wsum = self.tmass[i1] + self.tmass[i2]
        dist = math.sqrt(dx * dx + dy * dy)
        mag = dist - PARTICLE_DIAM
        if mag > 0:
            return
        scale = mag / wsum
        sd = scale / dist
        dpx, dpy = sd * dx, sd * dy
        c1, c2 = float(counts[i1]), float(counts[i2])
        t1, t2 = self.tmass[i1], self.tmass[i2]
        e1[0] += ((-t1) * dpx) / c1
        e1[1] += ((-t1) * dpy) / c1
        e2[0] += (t2 * dpx) / c2
        e2[1] += (t2 * dpy) / c2

    def project_rigid_contact(self, c, counts, stabile=False):
        """stabile: the STABILIZATION copy (rigidcontactconstraint.cpp:15,35,61-67,81-95): geometry from p (getP(true)), the
        correction goes to p, friction to p and ep"""
        i1, i2 = c[1], c[2]
        q1, q2 = self.ep[i1], self.ep[i2]          # always ep: friction
        e1, e2 = (self.p[i1], self.p[i2]) if stabile else (q1, q2)   # getP(stabile): geometry and correction
        g1x, g1y, d1 = self.sdf_data(i1)
        g2x, g2y, d2 = self.sdf_data(i2)
        if d1 < 0 or d2 < 0:
            x, y = e2[0] - e1[0], e2[1] - e1[1]
            ln = math.sqrt(x * x + y * y)
            d = PARTICLE_DIAM - ln
            if d < EPSILON:
                return
            nx, ny = x / ln, y / ln
        else:
            if d1 < d2:
                d, nx, ny = d1, g1x, g1y
            else:
                d, nx, ny = d2, -g2x, -g2y
            if d < PARTICLE_DIAM + EPSILON:
                # initBoundary (:13-27)
                x, y = e1[0] - e2[0], e1[1] - e2[1]
                ln = math.sqrt(x * x + y * y)
                d = PARTICLE_DIAM - ln
                if d < EPSILON:
                    return
                if ln > EPSILON:
                    x, y = x / ln, y / ln
                else:
                    x, y = 0., 1.
                dp = x * nx + y * ny
                if dp < 0:
                    nx, ny = x - (2.0 * dp) * nx, y - (2.0 * dp) * ny
                else:
                    nx, ny = x, y
        t1, t2 = self.tmass[i1], self.tmass[i2]
        wsum = t1 + t2
        s = (1.0 / wsum) * d
        dpx, dpy = s * nx, s * ny
        c1, c2 = float(counts[i1]), float(counts[i2])
        e1[0] += ((-t1) * dpx) / c1
        e1[1] += ((-t1) * dpy) / c1
        e2[0] += (t2 * dpx) / c2
        e2[1] += (t2 * dpy) / c2
        # friction (:68-95)
        inv = 1. / math.sqrt(nx * nx + ny * ny)
        nfx, nfy = nx * inv, ny * inv
        p1, p2 = self.p[i1], self.p[i2]
        fx, fy = (q1[0] - p1[0]) - (q2[0] - p2[0]), (q1[1] - p1[1]) - (q2[1] - p2[1])
        dn = fx * nfx + fy * nfy
        tx, ty = fx - dn * nfx, fy - dn * nfy
        ldpt = math.sqrt(tx * tx + ty * ty)
        if ldpt < EPSILON:
            return
        sfric = math.sqrt(self.sfric[i1] * self.sfric[i2])
        kfric = math.sqrt(self.kfric[i1] * self.kfric[i2])
        if not (ldpt < sfric * d):
            m = min(kfric * d / ldpt, 1.)
            tx, ty = tx * m, ty * m
        if stabile:
            p1[0] -= (tx * t1) / wsum
            p1[1] -= (ty * t1) / wsum
            p2[0] += (tx * t2) / wsum
            p2[1] += (ty * t2) / wsum
        q1[0] -= (tx * t1) / wsum
        q1[1] -= (ty * t1) / wsum
        q2[0] += (tx * t2) / wsum
        q2[1] += (ty * t2) / wsum

    def project_distance(self, c, counts):
        i1, i2, d = c["i1"], c["i2"], c["d"]
        if self.imass[i1] == 0. and self.imass[i2] == 0.:
            return
        e1, e2 = self.ep[i1], self.ep[i2]
        dx, dy = e1[0] - e2[0], e1[1] - e2[1]
        wsum = self.imass[i1] + self.imass[i2]
        dist = math.sqrt(dx * dx + dy * dy)
        mag = dist - d
        scale = mag / wsum
        sd = scale / dist
        dpx, dpy = sd * dx, sd * dy
        c1, c2 = float(counts[i1]), float(counts[i2])
        w1, w2 = self.imass[i1], self.imass[i2]
        e1[0] += ((-w1) * dpx) / c1
        e1[1] += ((-w1) * dpy) / c1
        e2[0] += (w2 * dpx) / c2
        e2[1] += (w2 * dpy) / c2

    def update_com(self, b):
        """Body::updateCOM(estimates, true), solver/particle.cpp:15-57"""
        tx = ty = 0.
        for i in b["particles"]:
            tx += self.ep[i][0] / self.imass[i]
            ty += self.ep[i][1] / self.imass[i]
        b["center"] = [tx * b["imass"], ty * b["imass"]]
        angle, prev = 0.0, 0.0
        for k, i in enumerate(b["particles"]):
            qx, qy = b["rs"][i]
            if qx * qx + qy * qy == 0:
                continue
            rx, ry = self.ep[i][0] - b["center"][0], self.ep[i][1] - b["center"][1]
            cos = rx * qx + ry * qy
            sin = ry * qx - rx * qy
            nxt = math.atan2(sin, cos)
            if k > 0:
                if prev - nxt >= M_PI:
                    nxt += 2 * M_PI
            else:
                if nxt < 0:
                    nxt += 2 * M_PI
            prev = nxt
            nxt /= self.imass[i]
            angle += nxt
        b["angle"] = angle * b["imass"]

    def project_shape(self, b):
        self.update_com(b)
        c, s = math.cos(b["angle"]), math.sin(b["angle"])
        for i in b["particles"]:
            qx, qy = b["rs"][i]
            gx, gy = (c * qx - s * qy) + b["center"][0], (s * qx + c * qy) + b["center"][1]
            e = self.ep[i]
            e[0] += (gx - e[0]) * b["stiffness"]
            e[1] += (gy - e[1]) * b["stiffness"]

    def project_fluid_or_gas(self, c, counts):
        gas = c["type"] == "gas"
        K = GAS_C if gas else FLUID_C
        p0, ps = c["p0"], c["ps"]
        n = self.n
        ep, imass, ph = self.ep, self.imass, self.ph
        lambdas = {}
        neighbors = []
        for i in ps:
            nb = []
            pi = denom = 0.
            ex, ey = ep[i]
            for j in range(n):
                if j != i:
                    if imass[j] == 0:
                        continue
                    rx, ry = ex - ep[j][0], ey - ep[j][1]
                    r2 = rx * rx + ry * ry
                    if r2 < H2:
                        nb.append(j)
                        incr = poly6(r2) / imass[j]
                        if ph[j] == SOLID:
                            incr *= K["S_SOLID"]
                        pi += incr
                        sx, sy = spiky_grad(rx, ry, math.sqrt(r2))
                        gx, gy = (-sx) / p0, (-sy) / p0
                        denom += gx * gx + gy * gy
                else:
                    nb.append(j)
                    pi += poly6(0) / imass[i]
            ox = oy = 0.
            for j in nb:
                rx, ry = ex - ep[j][0], ey - ep[j][1]
                sx, sy = spiky_grad(rx, ry, math.sqrt(rx * rx + ry * ry))
                w = K["S_SOLID"] if ph[j] == SOLID else 1.
                ox += w * sx
                oy += w * sy
            ox, oy = ox / p0, oy / p0
            denom += ox * ox + oy * oy
            p_rat = pi / p0
            if gas and c["open"]:
                s = (1. - p_rat)
                self.f[i][0] += (self.v[i][0] * s) * -50.
                self.f[i][1] += (self.v[i][1] * s) * -50.
            lambdas[i] = -(p_rat - 1.) / (denom + RELAXATION)
            neighbors.append(nb)
        c["lambdas"] = lambdas   # TotalFluidConstraint::lambdas survives the call (the FluidEmitter reads it)
        base6 = poly6(K["DQ_P"] * K["DQ_P"] * H * H)
        deltas = []
        for k, i in enumerate(ps):
            dx = dy = 0.
            fvx = fvy = 0.
            ex, ey = ep[i]
            li = lambdas[i]
            for j in neighbors[k]:
                if i == j:
                    continue
                rx, ry = ex - ep[j][0], ey - ep[j][1]
                rlen = math.sqrt(rx * rx + ry * ry)
                sx, sy = spiky_grad(rx, ry, rlen)
                corr = -K["K_P"] * math.pow(poly6(rlen * rlen) / base6, K["E_P"])
                s = li + lambdas.get(j, 0.) + corr
                dx += s * sx
                dy += s * sy
                if gas:  # pseudo-vorticity force (gasconstraint.cpp:99-103)
                    r2 = rx * rx + ry * ry
                    gx, gy = spiky_grad(rx, ry, r2)
                    wx, wy = gx * self.v[j][0], gy * self.v[j][1]
                    L = math.sqrt(wx * wx + wy * wy)
                    cx, cy = 0. * 0. - ry * L, L * rx - 0. * 0.
                    p6 = poly6(r2)
                    fvx += cx * p6
                    fvy += cy * p6
            deltas.append((dx / p0, dy / p0))
            if gas:
                self.f[i][0] += fvx
                self.f[i][1] += fvy
        for k, i in enumerate(ps):
            div = float(len(neighbors[k])) + counts[i]
            ep[i][0] += deltas[k][0] / div
            ep[i][1] += deltas[k][1] / div

    # ------------------------------------------------------------------ the tick
    def tick(self, dt):
        n = self.n
        p, v, ep, f = self.p, self.v, self.ep, self.f
        counts = [0] * n
        for i in range(n):
            gx, gy = self.gravity
            if self.ph[i] == GAS:
                gx, gy = gx * ALPHA, gy * ALPHA
            v[i][0] = v[i][0] + dt * gx + dt * f[i][0]
            v[i][1] = v[i][1] + dt * gy + dt * f[i][1]
            f[i][0] = f[i][1] = 0.
            if self.imass[i] == 0.:
                ep[i][0], ep[i][1] = p[i][0], p[i][1]
            else:
                ep[i][0], ep[i][1] = p[i][0] + dt * v[i][0], p[i][1] + dt * v[i][1]
            self.tmass[i] = 1. / ((1. / self.imass[i]) * math.exp(-p[i][1])) if self.imass[i] != 0.0 else 0.0
        contacts = []
        thr = PARTICLE_DIAM - EPSILON
        epa = np.array(ep)
        for i in range(n):
            # candidates by numpy, the decision in the reference's own arithmetic
            d2 = (epa[i + 1:, 0] - epa[i, 0]) ** 2 + (epa[i + 1:, 1] - epa[i, 1]) ** 2
            for j in (np.nonzero(d2 < .26)[0] + i + 1):
                j = int(j)
                if self.imass[i] == 0 and self.imass[j] == 0:
                    continue
                if self.ph[i] == SOLID and self.ph[j] == SOLID and self.bod[i] == self.bod[j] and self.bod[i] != -1:
                    continue
                x, y = ep[j][0] - ep[i][0], ep[j][1] - ep[i][1]
                if math.sqrt(x * x + y * y) < thr:
                    if self.ph[i] == SOLID and self.ph[j] == SOLID:
                        contacts.append(("rigid", i, j))
                    elif self.ph[i] == SOLID or self.ph[j] == SOLID:
                        contacts.append(("contact", i, j))
            if ep[i][0] < self.xb[0] + PARTICLE_RAD:
                contacts.append(("boundary", i, self.xb[0], True, True))
            elif ep[i][0] > self.xb[1] - PARTICLE_RAD:
                contacts.append(("boundary", i, self.xb[1], True, False))
            if ep[i][1] < self.yb[0] + PARTICLE_RAD:
                contacts.append(("boundary", i, self.yb[0], False, True))
            elif ep[i][1] > self.yb[1] - PARTICLE_RAD:
                contacts.append(("boundary", i, self.yb[1], False, False))
        self.num_contacts = len(contacts)
        self.last_contacts = contacts
        for c in contacts:
            counts[c[1]] += 1
            if c[0] != "boundary":
                counts[c[2]] += 1
        for c in self.standard:
            if c["type"] == "distance":
                counts[c["i1"]] += 1
                counts[c["i2"]] += 1
        for b in self.bodies:
            for i in b["particles"]:
                counts[i] += 1
        self.counts = counts
        # the reference's compile-time option USE_STABILIZATION (simulation.h:17, off in its build): STABILIZATION_ITERATIONS passes
        # over the stabile copies of the rigid contacts and wall constraints, in the CONTACT list's order, before the solver
        # iterations (simulation.cpp:190-192,204-223,249-271); the counts do not include them (:236-238)
        for _ in range(self.stabilization_iterations):
            for c in contacts:
                if c[0] == "boundary":
                    self.project_boundary(c, counts, True)
                elif c[0] == "rigid":
                    self.project_rigid_contact(c, counts, True)
        for _ in range(SOLVER_ITERATIONS):
            for c in contacts:
                if c[0] == "boundary":
                    self.project_boundary(c, counts)
                elif c[0] == "rigid":
                    self.project_rigid_contact(c, counts)
                else:
                    self.project_contact(c, counts)
            for c in self.standard:
                if c["type"] == "distance":
                    self.project_distance(c, counts)
                else:
                    self.project_fluid_or_gas(c, counts)
            for b in self.bodies:
                self.project_shape(b)
        for i in range(n):
            dx, dy = ep[i][0] - p[i][0], ep[i][1] - p[i][1]
            v[i][0], v[i][1] = dx / dt, dy / dt
            if math.sqrt(dx * dx + dy * dy) < EPSILON:
                v[i][0] = v[i][1] = 0.
                continue
            p[i][0], p[i][1] = ep[i][0], ep[i][1]
        for e in self.emitters:  # OpenSmokeEmitter::tick: one GAS particle of mass 1 per 1/rate seconds, appended to the gas
            e["timer"] += dt
            while e["timer"] >= 1. / e["rate"]:
                e["timer"] -= 1. / e["rate"]
                if e["gas"] is not None:
                    e["gas"]["ps"].append(len(self.p))
                    self.add_particle(e["posn"], 1., GAS)

        for e in self.fluid_emitters:  # FluidEmitter::tick, cpu/src/fluidemitter.cpp:13-79
            fs = e["fluid"]
            ps, lambdas = fs["ps"], fs.get("lambdas", {})
            for i in range(len(ps) - 1, -1, -1):
                k = ps[i]
                if math.sqrt(v[k][0] * v[k][0] + v[k][1] * v[k][1]) < .06 and p[k][1] <= 5:
                    # [sic] the hash is keyed by particle index but read with the loop index i; a missing key reads 0
                    if lambdas.get(i, 0.) <= 0:
                        self.t[k] -= 1
                        if self.t[k] <= 0:  # the particle freezes into an immovable solid and leaves the fluid
                            self.t[k] = 0
                            self.imass[k] = 0
                            self.ph[k] = SOLID
                            ep[k][0], ep[k][1] = p[k][0], p[k][1]
                            v[k][0] = v[k][1] = 0.
                            f[k][0] = f[k][1] = 0.
                            del ps[i]
                    else:
                        self.t[k] += dt
                        if self.t[k] > 3:
                            self.t[k] = 3
            e["timer"] += dt
            e["total_timer"] += dt
            while e["total_timer"] < 5 and e["timer"] >= 1. / e["rate"]:
                e["timer"] -= 1. / e["rate"]
                fs["ps"].append(len(self.p))
                self.add_particle(e["posn"], 1., FLUID)
                self.v[-1] = [self.frand(), 1.]

    def mouse_pressed(self, x, y):
        """Simulation::mousePressed (simulation.cpp:1305-1314): v += 7 * normalize(point - p) for every particle"""
        for i in range(self.n):
            dx, dy = x - self.p[i][0], y - self.p[i][1]
            inv = 1. / math.sqrt(dx * dx + dy * dy)
            self.v[i][0] += 7. * (dx * inv)
            self.v[i][1] += 7. * (dy * inv)

    def add_particle(self, pos, mass, phase):
        """Particle(pos, mass, phase) (particle.h:31-52) appended to the particle list"""
        self.p.append([pos[0], pos[1]])
        self.v.append([0., 0.])
        self.ep.append([0., 0.])
        self.f.append([0., 0.])
        im = -mass if mass <= 0 else 1. / mass
        self.imass.append(im)
        self.tmass.append(im)
        self.ph.append(phase)
        self.bod.append(-1)
        self.sfric.append(0.)
        self.kfric.append(0.)
        self.t.append(4.)

    def positions(self):
        return np.array(self.p)

    def velocities(self):
        return np.array(self.v)

    def kinetic_energy(self):
        e = 0.
        for i in range(self.n):
            if self.imass[i] != 0.:
                e += .5 * (self.v[i][0] * self.v[i][0] + self.v[i][1] * self.v[i][1]) / self.imass[i]
        return e
