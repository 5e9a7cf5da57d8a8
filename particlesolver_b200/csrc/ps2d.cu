// particlesolver_b200/csrc/ps2d.cu — the 2-D double-precision path (include/psolver2d.h): one Simulation::tick of the
// reference's CPU application on the GPU, every constraint group of its ITERATIVE solver.
// Compiled WITHOUT -use_fast_math and with -fmad=false: the reference is plain x86-64 double arithmetic without
// contraction, and a tick is compared with it to 1e-9.
//
// How a sequential (Gauss-Seidel) constraint list runs on the GPU without changing its result: LEVEL SCHEDULING.
// Two constraints commute unless they share a particle, so only the per-particle order of updates matters.  Each
// particle's constraints, in list order, form its sequence; level(c) = 1 + max over c's particles of the level of the
// constraint before c in that particle's sequence (0 if none).  Constraints of one level touch disjoint particles: they are
// projected in parallel, and levels are executed in ascending order.  For the per-tick CONTACT list (pairs found by
// an all-pairs test, then the walls, particle by particle — simulation.cpp:165-225) a particle's sequence is simply its
// contact partners in ascending index followed by its walls, so the levels are the least fixpoint of a local rule and are
// found by parallel relaxation (k2d_contact_levels).  Distance constraints are static: their schedule is built once on
// the host when the list changes.  Fluid / gas constraints are Jacobi inside and run as whole kernels at their place in
// the STANDARD list; their all-pairs neighbour loops (totalfluidconstraint.cpp:52-76) run one warp per particle.
//
// Two ways to issue a tick, same device functions (d_*) and therefore the same bits: the LAUNCH SEQUENCE (issue_tick: 11-29
// kernels, replayed as a CUDA graph) and, for scenes of up to kFusedTickMaxN particles, ONE KERNEL on one thread-block cluster
// (k2d_tick_fused) — see the comment above it for when each wins.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/psolver.h"
#include "../../include/psolver2d.h"

void ps_set_error(const char *fmt, ...);

namespace {
typedef uint32_t u32;
constexpr int kBlock = 128;
constexpr int kSerialBlock = 1024;                   // the level-scheduled kernels run as one CTA (scenes of 10^1..10^4 particles)
constexpr int kMaxC = PS2D_MAX_CONTACTS;             // pair slots per particle
constexpr int kEntries = kMaxC + 2;                  // + at most one x wall and one y wall
constexpr u32 kRigidBit = 0x80000000u;
constexpr double kRad = 0.25, kDiam = 0.5;           // PARTICLE_RAD / PARTICLE_DIAM, cpu/src/particle.h:6-7
constexpr double kEps = 1e-4;                        // EPSILON, cpu/src/includes.h:34
constexpr double kAlpha = -.2;                       // ALPHA, simulation.h:21
constexpr double kH = 2., kH2 = 4., kH6 = 64., kH9 = 512.;  // totalfluidconstraint.h:16-19, gasconstraint.h:4-7
constexpr double kRelax = .01;                       // RELAXATION
constexpr double kPi = 3.14159265358979323846;

struct FluidConsts {  // totalfluidconstraint.h:22-30 / gasconstraint.h:12-20
    double k_p, dq_p, s_solid;
    int gas, open;
};

__device__ __forceinline__ double poly6(double r2) {  // totalfluidconstraint.cpp:121-127
    if (r2 >= kH2) return 0.;
    const double term2 = kH2 - r2;
    return (315. / (64. * kPi * kH9)) * (term2 * term2 * term2);
}
// spikyGrad(r, x) = -normalize(r) * (45 / (pi H^6)) * (H - x)^2, zero for x >= H and x == 0 (:129-135); x is the length
// of r at every call site but the gas vorticity term, which passes dot(r, r) (gasconstraint.cpp:100);
// glm::normalize(v) = v * (1 / sqrt(dot(v, v)))
__device__ __forceinline__ double2 spiky_grad(double rx, double ry, double x) {
    if (x >= kH || x == 0.) return make_double2(0., 0.);
    const double inv = 1. / sqrt(rx * rx + ry * ry);
    const double c = 45. / (kPi * kH6), hm = kH - x;
    return make_double2((-(rx * inv) * c) * hm * hm, (-(ry * inv) * c) * hm * hm);
}

// (1)-(4): v += dt g (gas: ALPHA g) + dt f; f = 0; ep = p + dt v (fixed particles stay); tmass = height-scaled inverse
// mass; simulation.cpp:139-161, particle.h:56-58,67-73
__device__ __forceinline__ void d_predict(u32 i, double2 *v, double2 *ep, double2 *f, double *tmass, const double2 *p, const double *imass, const int *phase,
                                          double dt, double gx, double gy) {
    if (phase[i] == PS2D_PHASE_GAS) { gx = gx * kAlpha; gy = gy * kAlpha; }
    double2 vi = v[i];
    const double2 fi = f[i];
    vi.x = vi.x + dt * gx + dt * fi.x;
    vi.y = vi.y + dt * gy + dt * fi.y;
    v[i] = vi;
    f[i] = make_double2(0., 0.);
    const double2 pi = p[i];
    const double im = imass[i];
    ep[i] = im == 0. ? pi : make_double2(pi.x + dt * vi.x, pi.y + dt * vi.y);
    tmass[i] = im != 0. ? 1. / ((1. / im) * exp(-pi.y)) : 0.;
}
__global__ void k2d_predict(double2 *__restrict__ v, double2 *__restrict__ ep, double2 *__restrict__ f, double *__restrict__ tmass,
                            const double2 *__restrict__ p, const double *__restrict__ imass, const int *__restrict__ phase, u32 n, double dt, double gx,
                            double gy) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) d_predict(i, v, ep, f, tmass, p, imass, phase, dt, gx, gy);
}

// (6)-(8): the CONTACT list of a tick, stored per particle.  nb[i*kMaxC + r], r < cnt[i]: i's contact partners in
// ascending index (bit 31: both SOLID -> RigidContactConstraint, else ContactConstraint); flags: walls violated by the
// predicted position, at most one per axis, x before y (bit0 x-low, bit1 x-high, bit2 y-low, bit3 y-high).  The
// reference's list is [pairs (i, j > i) by j, wall x, wall y] for i = 0, 1, ...; particle i's own sequence in that list
// is therefore [nb ascending, wall x, wall y].  counts[i] = constraints on i in all groups (Constraint::updateCounts);
// draws[i] = wall constraints of i that draw jitter (fluid / gas particles only, boundaryconstraint.cpp:19).
__device__ __forceinline__ void d_find_contacts(u32 i, u32 lane, const double2 *ep, const double *imass, const int *phase, const int *bod,
                                                const u32 *static_counts, u32 n, double x0, double x1, double y0, double y1, u32 *nb, u32 *cnt, u32 *flags,
                                                u32 *counts, u32 *draws, u32 *overflow, int any_solid) {
    // one WARP per particle: lane l tests partners l, l + 32, ...; a ballot per 32 candidates keeps the partner list in
    // ascending index (the serial form, one thread walking all n candidates, was 65 % of a tick at n = 1452)
    const double2 e = ep[i];
    const double im = imass[i];
    const int ph = phase[i], bd = bod[i];
    u32 c = 0;
    if (any_solid) {  // no SOLID particle: no particle-particle contact constraint (simulation.cpp:188-196)
        for (u32 base = 0; base < n; base += 32) {
            const u32 j = base + lane;
            bool hit = false, solid2 = false;
            if (j < n && j != i) {
                const int phj = phase[j];
                solid2 = ph == PS2D_PHASE_SOLID && phj == PS2D_PHASE_SOLID;
                if ((solid2 || ph == PS2D_PHASE_SOLID || phj == PS2D_PHASE_SOLID) && !(im == 0. && imass[j] == 0.) && !(solid2 && bd == bod[j] && bd != -1)) {
                    const double2 q = ep[j];
                    const double dx = q.x - e.x, dy = q.y - e.y;
                    // the reference's test is sqrt(d2) < DIAM - EPS; the square root is only taken where d2 is within 1e-9 of the
                    // threshold's square (sqrt is monotone and correctly rounded, so the answer is the same everywhere else)
                    constexpr double kT = kDiam - kEps, kT2 = kT * kT;
                    const double d2 = dx * dx + dy * dy;
                    hit = d2 < kT2 * (1. - 1e-9) ? true : (d2 > kT2 * (1. + 1e-9) ? false : sqrt(d2) < kT);
                }
            }
            const u32 m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const u32 at = c + __popc(m & ((1u << lane) - 1u));
                if (at < (u32)kMaxC) nb[(size_t)i * kMaxC + at] = j | (solid2 ? kRigidBit : 0u);
            }
            c += __popc(m);
        }
    }
    if (lane != 0) return;
    if (c > (u32)kMaxC) { atomicMax(overflow, c); c = kMaxC; }
    u32 fl = 0;
    if (e.x < x0 + kRad) fl |= 1u; else if (e.x > x1 - kRad) fl |= 2u;
    if (e.y < y0 + kRad) fl |= 4u; else if (e.y > y1 - kRad) fl |= 8u;
    cnt[i] = c;
    flags[i] = fl;
    counts[i] = static_counts[i] + c + __popc(fl);
    draws[i] = (ph == PS2D_PHASE_FLUID || ph == PS2D_PHASE_GAS) ? __popc(fl) : 0u;
}

__global__ void __launch_bounds__(kBlock) k2d_find_contacts(const double2 *__restrict__ ep, const double *__restrict__ imass, const int *__restrict__ phase,
                                                            const int *__restrict__ bod, const u32 *__restrict__ static_counts, u32 n, double x0, double x1,
                                                            double y0, double y1, u32 *__restrict__ nb, u32 *__restrict__ cnt, u32 *__restrict__ flags,
                                                            u32 *__restrict__ counts, u32 *__restrict__ draws, u32 *__restrict__ overflow, int any_solid) {
    const u32 i = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (i >= n) return;  // warp-uniform
    d_find_contacts(i, threadIdx.x & 31, ep, imass, phase, bod, static_counts, n, x0, x1, y0, y1, nb, cnt, flags, counts, draws, overflow, any_solid);
}

// rank[i] = sum of counts before i (the position of i's first jitter draw in the rand() stream of an iteration);
// total -> *num.  One CTA, sequential over chunks.
__device__ void d_scan_counts(const u32 *counts, u32 *rank, u32 n, u32 *num) {  // one CTA (a multiple of 32 threads, at most 1024)
    __shared__ u32 sm[32];
    __shared__ u32 carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += blockDim.x) {
        const u32 i = base + threadIdx.x;
        const u32 c = i < n ? counts[i] : 0u;
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm[wid] = incl;
        __syncthreads();
        u32 before = carry, tot = 0;
        for (int w = 0; w < nw; w++) {
            if (w < wid) before += sm[w];
            tot += sm[w];
        }
        if (i < n) rank[i] = before + incl - c;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *num = carry;
}
__global__ void __launch_bounds__(1024) k2d_scan_counts(const u32 *__restrict__ counts, u32 *__restrict__ rank, u32 n, u32 *__restrict__ num) {
    d_scan_counts(counts, rank, n, num);
}

// Level schedule of the CONTACT list (see the file header).  Entry (i, r) is the r-th constraint of particle i's
// sequence; a pair constraint appears in both particles' sequences, q = its position in the partner's.  The rule
//     lvl(i, r) = 1 + max(lvl(i, r - 1), lvl(j, q - 1))      (pair with j; walls: the first term only)
// is monotone, so relaxing it from 0 in any order converges to its least fixpoint — the levels of the sequential list.
// info[0] = number of levels, info[1] = number of constraints in the list.
__device__ void d_contact_levels(const u32 *nb, const u32 *cnt, const u32 *flags, u32 n, unsigned char *nbq, u32 *lvl, u32 *info) {  // one CTA
    __shared__ u32 s_max, s_pairs, s_walls;
    if (threadIdx.x == 0) { s_max = 0; s_pairs = 0; s_walls = 0; }
    __syncthreads();
    u32 pairs = 0, walls = 0;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const u32 c = cnt[i], len = c + __popc(flags[i]);
        for (u32 r = 0; r < c; r++) {
            const u32 j = nb[(size_t)i * kMaxC + r] & ~kRigidBit, cj = cnt[j];
            u32 q = 0;
            while (q < cj && (nb[(size_t)j * kMaxC + q] & ~kRigidBit) != i) q++;
            nbq[(size_t)i * kMaxC + r] = (unsigned char)q;  // found by symmetry of the contact test
        }
        for (u32 r = 0; r < len; r++) lvl[(size_t)i * kEntries + r] = 0;
        pairs += c;
        walls += len - c;
    }
    atomicAdd(&s_pairs, pairs);
    atomicAdd(&s_walls, walls);
    __syncthreads();
    for (;;) {
        int changed = 0;
        for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
            const u32 c = cnt[i], len = c + __popc(flags[i]);
            u32 prev = 0;
            for (u32 r = 0; r < len; r++) {
                u32 other = 0;
                if (r < c) {
                    const u32 j = nb[(size_t)i * kMaxC + r] & ~kRigidBit, q = nbq[(size_t)i * kMaxC + r];
                    if (q > 0) other = *((volatile u32 *)&lvl[(size_t)j * kEntries + q - 1]);
                }
                const u32 v = 1 + max(prev, other);
                if (v != lvl[(size_t)i * kEntries + r]) { lvl[(size_t)i * kEntries + r] = v; changed = 1; }
                prev = v;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    u32 m = 0;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const u32 len = cnt[i] + __popc(flags[i]);
        if (len) m = max(m, lvl[(size_t)i * kEntries + len - 1]);
    }
    atomicMax(&s_max, m);
    __syncthreads();
    if (threadIdx.x == 0) { info[0] = s_max; info[1] = s_pairs / 2 + s_walls; }
}

__global__ void __launch_bounds__(kSerialBlock) k2d_contact_levels(const u32 *__restrict__ nb, const u32 *__restrict__ cnt, const u32 *__restrict__ flags,
                                                                   u32 n, unsigned char *__restrict__ nbq, u32 *__restrict__ lvl, u32 *__restrict__ info) {
    d_contact_levels(nb, cnt, flags, n, nbq, lvl, info);
}

struct Particles2D {  // device views used by the projection kernels
    double2 *ep;
    double2 *p;  // written by the stabilization pass only
    const double *tmass, *sfric, *kfric;
    const int *phase, *bod;
    const u32 *counts;
    const double2 *sdf_grad;
    const double *sdf_dist;
    const double *body_angle;
};

// Particle::getSDFData (solver/particle.cpp:3-13): the body's SDF sample rotated by the body's current angle
__device__ __forceinline__ double3 sdf_data(const Particles2D &P, u32 i) {
    const int b = P.bod[i];
    if (P.phase[i] != PS2D_PHASE_SOLID || b < 0) return make_double3(0., 0., -1.);
    const double2 g = P.sdf_grad[i];
    const double a = P.body_angle[b];
    const double c = cos(a), s = sin(a);
    return make_double3(g.x * c - g.y * s, g.x * s + g.y * c, P.sdf_dist[i]);
}

// ContactConstraint::project (contactconstraint.cpp:13-42)
__device__ void project_contact(const Particles2D &P, u32 i1, u32 i2) {
    const double t1 = P.tmass[i1], t2 = P.tmass[i2];
    if (t1 == 0. && t2 == 0.) return;
    double2 e1 = P.ep[i1], e2 = P.ep[i2];
    const double dx = e1.x - e2.x, dy = e1.y - e2.y;
    const double wsum = t1 + t2, dist = sqrt(dx * dx + dy * dy), mag = dist - kDiam;
    if (mag > 0.) return;
    const double sd = (mag / wsum) / dist;
    const double dpx = sd * dx, dpy = sd * dy;
    const double c1 = (double)P.counts[i1], c2 = (double)P.counts[i2];
    e1.x += ((-t1) * dpx) / c1; e1.y += ((-t1) * dpy) / c1;
    e2.x += (t2 * dpx) / c2;    e2.y += (t2 * dpy) / c2;
    P.ep[i1] = e1; P.ep[i2] = e2;
}

// RigidContactConstraint::project (rigidcontactconstraint.cpp:29-96), stabile == false
// stabile: the STABILIZATION copy of the constraint (the reference's option USE_STABILIZATION; :15,35,61-67,81-95): geometry from
// p (getP(true)) and the correction goes to p; friction reads ep - p and moves both
__device__ void project_rigid_contact(const Particles2D &P, u32 i1, u32 i2, bool stabile) {
    double2 q1 = P.ep[i1], q2 = P.ep[i2];          // ep
    double2 p1 = P.p[i1], p2 = P.p[i2];
    double2 e1 = stabile ? p1 : q1, e2 = stabile ? p2 : q2;   // getP(stabile)
    const double3 s1 = sdf_data(P, i1), s2 = sdf_data(P, i2);
    double d, nx, ny;
    if (s1.z < 0. || s2.z < 0.) {
        const double x = e2.x - e1.x, y = e2.y - e1.y;
        const double len = sqrt(x * x + y * y);
        d = kDiam - len;
        if (d < kEps) return;
        nx = x / len; ny = y / len;
    } else {
        if (s1.z < s2.z) { d = s1.z; nx = s1.x; ny = s1.y; }
        else { d = s2.z; nx = -s2.x; ny = -s2.y; }
        if (d < kDiam + kEps) {  // initBoundary (:13-27)
            double x = e1.x - e2.x, y = e1.y - e2.y;
            const double len = sqrt(x * x + y * y);
            d = kDiam - len;
            if (d < kEps) return;
            if (len > kEps) { x = x / len; y = y / len; } else { x = 0.; y = 1.; }
            const double dp = x * nx + y * ny;
            if (dp < 0.) { nx = x - (2.0 * dp) * nx; ny = y - (2.0 * dp) * ny; }
            else { nx = x; ny = y; }
        }
    }
    const double t1 = P.tmass[i1], t2 = P.tmass[i2];
    const double wsum = t1 + t2;
    const double s = (1.0 / wsum) * d;
    const double dpx = s * nx, dpy = s * ny;
    const double c1 = (double)P.counts[i1], c2 = (double)P.counts[i2];
    e1.x += ((-t1) * dpx) / c1; e1.y += ((-t1) * dpy) / c1;
    e2.x += (t2 * dpx) / c2;    e2.y += (t2 * dpy) / c2;
    if (stabile) { p1 = e1; p2 = e2; } else { q1 = e1; q2 = e2; }
    // friction (:68-95)
    const double inv = 1. / sqrt(nx * nx + ny * ny);
    const double nfx = nx * inv, nfy = ny * inv;
    const double fx = (q1.x - p1.x) - (q2.x - p2.x), fy = (q1.y - p1.y) - (q2.y - p2.y);
    const double dn = fx * nfx + fy * nfy;
    double tx = fx - dn * nfx, ty = fy - dn * nfy;
    const double ldpt = sqrt(tx * tx + ty * ty);
    if (!(ldpt < kEps)) {
        const double sfric = sqrt(P.sfric[i1] * P.sfric[i2]), kfric = sqrt(P.kfric[i1] * P.kfric[i2]);
        if (!(ldpt < sfric * d)) {
            const double m = fmin(kfric * d / ldpt, 1.);
            tx = tx * m; ty = ty * m;
        }
        if (stabile) {
            p1.x -= (tx * t1) / wsum; p1.y -= (ty * t1) / wsum;
            p2.x += (tx * t2) / wsum; p2.y += (ty * t2) / wsum;
        }
        q1.x -= (tx * t1) / wsum; q1.y -= (ty * t1) / wsum;
        q2.x += (tx * t2) / wsum; q2.y += (ty * t2) / wsum;
    }
    P.ep[i1] = q1; P.ep[i2] = q2;
    if (stabile) { P.p[i1] = p1; P.p[i2] = p2; }
}

// BoundaryConstraint::project (boundaryconstraint.cpp:14-93).  raw < 0: no jitter draw (solids).  stabile: the STABILIZATION copy
// (:32-35,74-76): p follows ep, no friction; the validity test reads ep either way
__device__ void project_boundary(const Particles2D &P, u32 i, double value, bool is_x, bool greater, int raw, bool stabile) {
    double2 e = P.ep[i];
    const double extra = raw >= 0 ? (double)(float)((double)raw / 2147483647.0) * .003 : 0.;  // frand() is float-typed, includes.h:25
    const double d = kRad + extra;
    double nx, ny;
    if (greater) {
        if (is_x) { if (e.x >= value + kRad) return; e.x = value + d; nx = 1.; ny = 0.; }
        else { if (e.y >= value + kRad) return; e.y = value + d; nx = 0.; ny = 1.; }
    } else {
        if (is_x) { if (e.x <= value - kRad) return; e.x = value - d; nx = -1.; ny = 0.; }
        else { if (e.y <= value - kRad) return; e.y = value - d; nx = 0.; ny = -1.; }
    }
    if (stabile) {
        double2 q = P.p[i];
        if (is_x) q.x = e.x; else q.y = e.y;
        P.p[i] = q;
        P.ep[i] = e;
        return;
    }
    // friction: walls have a coefficient of friction of 1 (:72-92)
    const double2 p = P.p[i];
    const double cn = (double)P.counts[i];
    const double dpx = (e.x - p.x) / cn, dpy = (e.y - p.y) / cn;
    const double dn = dpx * nx + dpy * ny;
    const double tx = dpx - dn * nx, ty = dpy - dn * ny;
    const double ldpt = sqrt(tx * tx + ty * ty);
    if (!(ldpt < kEps)) {
        if (ldpt < sqrt(P.sfric[i]) * d) { e.x -= tx; e.y -= ty; }
        else {
            const double m = fmin(sqrt(P.kfric[i]) * d / ldpt, 1.);
            e.x -= tx * m; e.y -= ty * m;
        }
    }
    P.ep[i] = e;
}

// One solver iteration over the CONTACT list, level by level.  Thread t owns particles t, t + 1024, ...; cur[i] walks
// particle i's sequence (levels strictly increase along it).  A pair constraint is projected by its lower-index owner
// (it is at the same level in both sequences); the partner just steps over it.
__device__ void d_contact_project(const Particles2D &P, const u32 *nb, const u32 *cnt, const u32 *flags, const u32 *lvl, const u32 *rank, const u32 *info, u32 *cur,
                                  u32 n, const int *raw, u32 window_base, u32 iteration, double x0, double x1, double y0, double y1, bool stabile) {
    // stabile: one pass over the STABILIZATION list = the rigid contacts and wall constraints of the CONTACT list, in its order
    // (simulation.cpp:190-192,204-223), so the same level schedule holds with the plain contacts stepped over.
    // info[-1] = jittered wall constraints of this tick (k2d_scan_counts): iteration t draws window[base + t * num + position]
    const u32 levels = info[0], draw_base = window_base + iteration * info[-1];
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) cur[i] = 0;
    for (u32 l = 1; l <= levels; l++) {
        __syncthreads();
        for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
            const u32 c = cnt[i], fl = flags[i], len = c + __popc(fl), r = cur[i];
            if (r >= len || lvl[(size_t)i * kEntries + r] != l) continue;
            cur[i] = r + 1;
            if (r < c) {
                const u32 w = nb[(size_t)i * kMaxC + r], j = w & ~kRigidBit;
                if (i < j) {
                    if (w & kRigidBit) project_rigid_contact(P, i, j, stabile);
                    else if (!stabile) project_contact(P, i, j);
                }
            } else {
                const bool second = r > c, is_x = !second && (fl & 3u);
                const int ph = P.phase[i];
                int draw = -1;
                if (ph == PS2D_PHASE_FLUID || ph == PS2D_PHASE_GAS) draw = raw[draw_base + rank[i] + (second ? 1u : 0u)];
                if (is_x) project_boundary(P, i, (fl & 1u) ? x0 : x1, true, (fl & 1u) != 0, draw, stabile);
                else project_boundary(P, i, (fl & 4u) ? y0 : y1, false, (fl & 4u) != 0, draw, stabile);
            }
        }
    }
}

__global__ void __launch_bounds__(kSerialBlock) k2d_contact_project(Particles2D P, const u32 *__restrict__ nb, const u32 *__restrict__ cnt,
                                                                    const u32 *__restrict__ flags, const u32 *__restrict__ lvl, const u32 *__restrict__ rank,
                                                                    const u32 *__restrict__ info, u32 *__restrict__ cur, u32 n, const int *__restrict__ raw,
                                                                    const u32 *__restrict__ window_base_ptr, u32 iteration, double x0, double x1, double y0, double y1, bool stabile) {
    // the draws' base is a device word, so that the tick's launch sequence can be replayed as a graph
    d_contact_project(P, nb, cnt, flags, lvl, rank, info, cur, n, raw, *window_base_ptr, iteration, x0, x1, y0, y1, stabile);
}

// DistanceConstraint::project (distanceconstraint.cpp:20-40) for one run of consecutive distance constraints of the
// STANDARD list, stored level by level (level_off[l] .. level_off[l + 1]); the schedule is built on the host when the
// list changes (constraints are static).
// x / d for d = 2^k is x * 2^-k: both are the correctly rounded value of the same real number, bit for bit (a double-precision
// divide is a ~25-instruction dependent sequence on the GPU; the rope's links divide by w1 + w2 = 2 and by constraint counts of 1 or 2)
__device__ __forceinline__ bool pow2_reciprocal(double d, double *inv) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    const unsigned ex = (unsigned)(b >> 52);  // sign 0 and a zero mantissa, or the test below fails
    if ((b & 0x800fffffffffffffull) != 0ull || ex < 2u || ex > 2044u) return false;
    *inv = __longlong_as_double((long long)((unsigned long long)(2046u - ex) << 52));
    return true;
}
// Everything a distance constraint needs that does not depend on the positions, prepared once per tick by all threads (inverse
// masses are constant, the constraint counts are fixed once the tick's contacts are known): the level loop is a dependent chain
// on ONE warp, whose every instruction — not only the double-precision ones — costs issue latency.  A divisor is stored as its
// reciprocal when it is a power of two (flag bit set), else as itself.
struct __align__(16) DistSlot {  // 64 bytes
    u32 i1, i2, flags, pad;      // flags: 1 live, 2 / 4 / 8: wsum / count 1 / count 2 stored as reciprocals
    double rest, nw1, w2, wsum, c1, c2;  // nw1 = -w1
};
__device__ __forceinline__ DistSlot dist_make_slot(u32 i1, u32 i2, double rest, const double *imass, const u32 *counts) {
    DistSlot r;
    r.i1 = i1; r.i2 = i2; r.flags = 0u; r.pad = 0u; r.rest = rest;
    const double w1 = imass[i1], w2 = imass[i2];
    r.nw1 = -w1; r.w2 = w2;
    r.wsum = w1 + w2; r.c1 = (double)counts[i1]; r.c2 = (double)counts[i2];
    if (!(w1 == 0. && w2 == 0.)) r.flags |= 1u;
    double inv;
    if (pow2_reciprocal(r.wsum, &inv)) { r.wsum = inv; r.flags |= 2u; }
    if (pow2_reciprocal(r.c1, &inv)) { r.c1 = inv; r.flags |= 4u; }
    if (pow2_reciprocal(r.c2, &inv)) { r.c2 = inv; r.flags |= 8u; }
    return r;
}
// DistanceConstraint::project (distanceconstraint.cpp:20-40)
__device__ __forceinline__ void dist_project(double2 *ep, const DistSlot &r) {
    if (!(r.flags & 1u)) return;
    double2 e1 = ep[r.i1], e2 = ep[r.i2];
    const double dx = e1.x - e2.x, dy = e1.y - e2.y;
    const double dist = sqrt(dx * dx + dy * dy), mag = dist - r.rest;
    if ((r.flags & 14u) == 14u) {  // the straight line: one sqrt and one divide
        const double sd = (mag * r.wsum) / dist;
        const double dpx = sd * dx, dpy = sd * dy;
        e1.x += (r.nw1 * dpx) * r.c1; e1.y += (r.nw1 * dpy) * r.c1;
        e2.x += (r.w2 * dpx) * r.c2;  e2.y += (r.w2 * dpy) * r.c2;
    } else {
        const double sd = ((r.flags & 2u) ? mag * r.wsum : mag / r.wsum) / dist;
        const double dpx = sd * dx, dpy = sd * dy;
        if (r.flags & 4u) { e1.x += (r.nw1 * dpx) * r.c1; e1.y += (r.nw1 * dpy) * r.c1; }
        else { e1.x += (r.nw1 * dpx) / r.c1; e1.y += (r.nw1 * dpy) / r.c1; }
        if (r.flags & 8u) { e2.x += (r.w2 * dpx) * r.c2; e2.y += (r.w2 * dpy) * r.c2; }
        else { e2.x += (r.w2 * dpx) / r.c2; e2.y += (r.w2 * dpy) / r.c2; }
    }
    ep[r.i1] = e1; ep[r.i2] = e2;
}
// slots[k - rec_base] for the records k of a run (or of all runs), by all threads of the caller
__device__ __forceinline__ void d_distance_prepare(DistSlot *slots, u32 first, u32 count, const double *imass, const u32 *counts, const u32 *i1s, const u32 *i2s,
                                                   const double *rest, u32 tid, u32 nthreads) {
    for (u32 k = tid; k < count; k += nthreads) slots[k] = dist_make_slot(i1s[first + k], i2s[first + k], rest[first + k], imass, counts);
}
// The levels of one run, level_off[l] .. level_off[l + 1] (absolute record numbers; slots[0] is record rec_base).
// kWarp: the caller is one warp and the levels are at most 32 wide — a warp barrier per level instead of a CTA barrier.
template <bool kWarp>
__device__ __forceinline__ void d_distance_levels(double2 *ep, const DistSlot *slots, u32 rec_base, const u32 *level_off, u32 levels, u32 tid, u32 nthreads) {
    if (!levels) return;
    u32 b = level_off[0], e = level_off[1];
    u32 e_next = levels > 1 ? level_off[2] : e;
    for (u32 l = 0; l < levels; l++) {
        const u32 e_after = l + 3 <= levels ? level_off[l + 3] : e_next;  // offsets two levels ahead: never waited for
        for (u32 k = b + tid; k < e; k += nthreads) dist_project(ep, slots[k - rec_base]);
        if (kWarp) __syncwarp(); else __syncthreads();
        b = e; e = e_next; e_next = e_after;
    }
}
// runs too long for shared memory: straight from global memory, slots made on the fly
__global__ void __launch_bounds__(kSerialBlock) k2d_distance_run(double2 *__restrict__ ep, const double *__restrict__ imass, const u32 *__restrict__ counts,
                                                                 const u32 *__restrict__ i1s, const u32 *__restrict__ i2s, const double *__restrict__ rest,
                                                                 const u32 *__restrict__ level_off, u32 levels) {
    for (u32 l = 0; l < levels; l++) {
        if (l) __syncthreads();
        const u32 b = level_off[l], e = level_off[l + 1];
        for (u32 k = b + threadIdx.x; k < e; k += kSerialBlock) dist_project(ep, dist_make_slot(i1s[k], i2s[k], rest[k], imass, counts));
    }
}

// The same out of shared memory: the predicted positions of all n particles and the run's prepared slots.  A rope of L links is L
// levels of one constraint each — a dependent chain whose every step used to wait for two global round trips (record, then
// positions: 1.3 us per level on B200, 65 % of a tick of the gas-rope scene); from shared memory, with the position-independent
// part of every constraint prepared up front, a level costs two 16-byte loads, the sqrt / divide chain and a warp barrier.
// Same expressions in the same order per constraint: bit-identical to k2d_distance_run.
template <int kThreads>  // 32: the levels are at most a warp wide (ropes), one warp and warp barriers; else a CTA of up to kSerialBlock threads
__global__ void __launch_bounds__(kThreads) k2d_distance_run_staged(double2 *__restrict__ ep, const double *__restrict__ imass, const u32 *__restrict__ counts,
                                                                        const u32 *__restrict__ i1s, const u32 *__restrict__ i2s, const double *__restrict__ rest,
                                                                        const u32 *__restrict__ level_off, u32 levels, u32 n, u32 first, u32 count) {
    extern __shared__ __align__(16) unsigned char stage2d[];
    DistSlot *slots = reinterpret_cast<DistSlot *>(stage2d);                 // count
    double2 *s_ep = reinterpret_cast<double2 *>(slots + count);              // n
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) s_ep[i] = ep[i];
    d_distance_prepare(slots, first, count, imass, counts, i1s, i2s, rest, threadIdx.x, blockDim.x);
    __syncthreads();
    if (kThreads == 32) d_distance_levels<true>(s_ep, slots, first, level_off, levels, threadIdx.x, 32u);
    else d_distance_levels<false>(s_ep, slots, first, level_off, levels, threadIdx.x, blockDim.x);
    __syncthreads();
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) ep[i] = s_ep[i];
}
static inline size_t distance_stage_bytes(u32 n, u32 count) { return (size_t)n * 16 + (size_t)count * sizeof(DistSlot) + 16; }
constexpr size_t kDistanceStageMax = 200 * 1024;

// TotalShapeConstraint::project for every body (totalshapeconstraint.cpp:14-24): Body::updateCOM (centre of mass and the
// mass-weighted mean angle, with the reference's sequential unwrapping of consecutive angles, solver/particle.cpp:15-57),
// then every member moves to its rotated rest position.  Bodies own disjoint particles, so they run in parallel; one
// thread per body keeps the reference's summation order (bodies are a dozen particles).
__device__ void d_shape(u32 b, double2 *ep, const double *imass, const double2 *rs, const u32 *b_first, const u32 *b_count, const double *b_imass, const double *b_stiff,
                        double2 *b_center, double *b_angle) {
    const u32 first = b_first[b], count = b_count[b];
    const double bim = b_imass[b];
    double tx = 0., ty = 0.;
    for (u32 k = 0; k < count; k++) {
        const double2 e = ep[first + k];
        const double im = imass[first + k];
        tx += e.x / im; ty += e.y / im;
    }
    const double cx = tx * bim, cy = ty * bim;
    double angle = 0., prev = 0.;
    for (u32 k = 0; k < count; k++) {
        const double2 q = rs[first + k];
        if (q.x * q.x + q.y * q.y == 0.) continue;
        const double2 e = ep[first + k];
        const double rx = e.x - cx, ry = e.y - cy;
        const double co = rx * q.x + ry * q.y, si = ry * q.x - rx * q.y;
        double next = atan2(si, co);
        if (k > 0) { if (prev - next >= kPi) next += 2 * kPi; }
        else { if (next < 0.) next += 2 * kPi; }
        prev = next;
        next /= imass[first + k];
        angle += next;
    }
    angle *= bim;
    b_center[b] = make_double2(cx, cy);
    b_angle[b] = angle;
    const double c = cos(angle), s = sin(angle), stiff = b_stiff[b];
    for (u32 k = 0; k < count; k++) {
        const double2 q = rs[first + k];
        const double gx = (c * q.x - s * q.y) + cx, gy = (s * q.x + c * q.y) + cy;
        double2 e = ep[first + k];
        e.x += (gx - e.x) * stiff; e.y += (gy - e.y) * stiff;
        ep[first + k] = e;
    }
}

__global__ void __launch_bounds__(kBlock) k2d_shape(double2 *__restrict__ ep, const double *__restrict__ imass, const double2 *__restrict__ rs,
                                                    const u32 *__restrict__ b_first, const u32 *__restrict__ b_count, const double *__restrict__ b_imass,
                                                    const double *__restrict__ b_stiff, double2 *__restrict__ b_center, double *__restrict__ b_angle, u32 nb) {
    const u32 b = blockIdx.x * kBlock + threadIdx.x;
    if (b < nb) d_shape(b, ep, imass, rs, b_first, b_count, b_imass, b_stiff, b_center, b_angle);
}

// TotalFluidConstraint / GasConstraint::project, first loop (totalfluidconstraint.cpp:45-93, gasconstraint.cpp:33-85): lambda
// of every particle of STANDARD constraint `op`, 0 for everybody else (the constraint's lambdas is a QHash cleared per
// call: any other particle reads 0, :106).  SOLID neighbours count S_SOLID-fold, immovable ones not at all.
// One WARP per particle, in two kinds of trips over the reference's all-pairs loop: a CHEAP one per 32 candidates (distance test
// only; a ballot compacts the hits, in ascending index, into a 64-entry ring in shared memory) and a HEAVY one per 32 neighbours
// found (one neighbour per lane: the 2 sqrt + 4 divides of the kernel terms) — a particle has some 25 neighbours among hundreds of
// candidates, and with the terms evaluated inside the candidate trip every trip paid for them (14 heavy trips per particle at
// N = 432; now 1).  Lane partial sums meet in a butterfly of xor-shuffles: deterministic; the summation order differs from the
// reference's sequential one, i.e. the result agrees to rounding, ~1e-16 relative, instead of bit for bit (measured in
// tests/test_gpu_2d_full.py).  The serial form of this loop (one thread per particle) cost 115 + 177 us per constraint at N = 432.
constexpr u32 kRing = 64;  // entries per warp: fewer than 32 wait when up to 32 arrive
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// the candidates of particle i within H (and i itself when with_self), 32 at a time in ascending index, handed to visit(j)
// one per lane; returns how many there were.  ring: this warp's kRing words of shared memory.  keep != nullptr: the first kNbrKeep
// of them are also written there, for the second loop of the constraint (the reference reuses its neighbour lists the same way, :96).
constexpr u32 kNbrKeep = 128;  // list entries kept per particle (2-D fluids at rest have ~25-50 neighbours within H; more: the second loop searches again)
template <class Visit>
__device__ __forceinline__ u32 for_each_neighbor_2d(u32 i, u32 lane, u32 *ring, const double2 &pi, const double2 *ep, const double *imass, u32 n, bool with_self,
                                                    u32 *keep, Visit visit) {
    const u32 lt = (1u << lane) - 1u;
    u32 head = 0, cnt = 0, total = 0, kept = 0;
    for (u32 base = 0; base < n; base += 32) {
        const u32 j = base + lane;
        bool hit = false;
        if (j < n) {
            if (j == i) hit = with_self;
            else if (imass[j] != 0.) {  // fixed particles are ignored
                const double2 pj = ep[j];
                const double rx = pi.x - pj.x, ry = pi.y - pj.y;
                hit = rx * rx + ry * ry < kH2;
            }
        }
        const u32 m = __ballot_sync(0xffffffffu, hit);
        if (!m) continue;
        if (hit) ring[(head + cnt + __popc(m & lt)) & (kRing - 1)] = j;
        cnt += __popc(m); total += __popc(m);
        __syncwarp();
        if (cnt >= 32u) {
            const u32 jj = ring[(head + lane) & (kRing - 1)];
            if (keep && kept + lane < kNbrKeep) keep[kept + lane] = jj;
            kept += 32u;
            visit(jj);
            head = (head + 32u) & (kRing - 1); cnt -= 32u;
            __syncwarp();
        }
    }
    if (lane < cnt) {
        const u32 jj = ring[(head + lane) & (kRing - 1)];
        if (keep && kept + lane < kNbrKeep) keep[kept + lane] = jj;
        visit(jj);
    }
    __syncwarp();
    return total;
}
__device__ void d_fluid_lambda(u32 i, u32 lane, u32 *ring, const double2 *ep, const double *imass, const int *phase, const int *group, u32 n, int op, double p0,
                               const FluidConsts &K, double *lambda, u32 *nbcount, u32 *nbr_keep, const double2 *v, double2 *f) {
    if (group[i] != op) { if (lane == 0) lambda[i] = 0.; return; }
    const double2 pi = ep[i];
    double rho = 0., denom = 0., ox = 0., oy = 0.;
    const u32 nbc = for_each_neighbor_2d(i, lane, ring, pi, ep, imass, n, true, nbr_keep ? nbr_keep + (size_t)i * kNbrKeep : nullptr, [&](u32 j) {
        const double im = imass[j];
        if (j == i) {  // the particle itself (:78-81)
            rho += poly6(0.) / im;
            return;
        }
        const double2 pj = ep[j];
        const double rx = pi.x - pj.x, ry = pi.y - pj.y;
        const double r2 = rx * rx + ry * ry;
        double incr = poly6(r2) / im;
        const bool solid = phase[j] == PS2D_PHASE_SOLID;
        if (solid) incr *= K.s_solid;
        rho += incr;
        const double2 sg = spiky_grad(rx, ry, sqrt(r2));
        const double gx = -sg.x / p0, gy = -sg.y / p0;  // grad(k, j) = -spikyGrad / p0 (:137-144)
        denom += gx * gx + gy * gy;
        const double w = solid ? K.s_solid : 1.;       // grad(k, i) = sum_j w_j spikyGrad / p0 (:146-157)
        ox += w * sg.x; oy += w * sg.y;
    });
    rho = warp_sum_d(rho); denom = warp_sum_d(denom); ox = warp_sum_d(ox); oy = warp_sum_d(oy);
    if (lane != 0) return;
    ox = ox / p0; oy = oy / p0;
    denom += ox * ox + oy * oy;
    const double p_rat = rho / p0;
    if (K.gas && K.open) {  // open-boundary drag, gasconstraint.cpp:80
        const double2 vi = v[i];
        double2 fi = f[i];
        const double s = 1. - p_rat;
        fi.x += (vi.x * s) * -50.; fi.y += (vi.y * s) * -50.;
        f[i] = fi;
    }
    lambda[i] = -(p_rat - 1.) / (denom + kRelax);
    nbcount[i] = nbc;
}

__global__ void __launch_bounds__(kBlock) k2d_fluid_lambda(const double2 *__restrict__ ep, const double *__restrict__ imass, const int *__restrict__ phase,
                                                           const int *__restrict__ group, u32 n, int op, double p0, FluidConsts K, double *__restrict__ lambda,
                                                           u32 *__restrict__ nbcount, u32 *__restrict__ nbr_keep, const double2 *__restrict__ v,
                                                           double2 *__restrict__ f) {
    __shared__ u32 rings[kBlock / 32][kRing];
    const u32 i = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (i >= n) return;  // warp-uniform
    d_fluid_lambda(i, threadIdx.x & 31, rings[threadIdx.x >> 5], ep, imass, phase, group, n, op, p0, K, lambda, nbcount, nbr_keep, v, f);
}

// second loop (:95-111): delta_i = sum_j (lambda_i + lambda_j + s_corr) spikyGrad / p0, divided by (#neighbours incl. self +
// constraint count) (:113-115).  Written to `delta`, applied by k2d_fluid_apply: all deltas of a constraint come from the
// same ep.  Gas: the pseudo-vorticity force of gasconstraint.cpp:99-107 goes to the force accumulator.  One warp per particle,
// the same two kinds of trips.
__device__ void d_fluid_delta(u32 i, u32 lane, u32 *ring, const double2 *ep, const double *imass, u32 n, double p0, const FluidConsts &K, const double *lambda,
                              const u32 *nbcount, const u32 *nbr_keep, const u32 *counts, double2 *delta, const double2 *v, double2 *f) {
    const double2 pi = ep[i];
    const double li = __ldcg(lambda + i);  // (.cg: in the fused tick another CTA of the cluster may have written it)
    const double base6 = poly6(K.dq_p * K.dq_p * kH * kH);
    double dx = 0., dy = 0., fvx = 0., fvy = 0.;
    auto visit = [&](u32 j) {
        const double2 pj = ep[j];
        const double rx = pi.x - pj.x, ry = pi.y - pj.y;
        const double r2 = rx * rx + ry * ry;
        const double rlen = sqrt(r2);
        const double2 sg = spiky_grad(rx, ry, rlen);
        const double q = poly6(rlen * rlen) / base6, q2 = q * q;
        const double corr = -K.k_p * (q2 * q2);  // pow(q, E_P), E_P = 4, as two squarings (within 1 ulp of libm's pow)
        const double s = (li + __ldcg(lambda + j)) + corr;
        dx += s * sg.x; dy += s * sg.y;
        if (K.gas) {
            const double2 g = spiky_grad(rx, ry, r2);  // [sic] the squared length as the length
            const double2 vj = v[j];
            const double wx = g.x * vj.x, wy = g.y * vj.y;
            const double L = sqrt(wx * wx + wy * wy);
            const double cx = 0. * 0. - ry * L, cy = L * rx - 0. * 0.;  // cross((0,0,L), (rx,ry,0))
            const double p6 = poly6(r2);
            fvx += cx * p6; fvy += cy * p6;
        }
    };
    const u32 nbc = __ldcg(nbcount + i);  // neighbours the first loop found, the particle itself included
    if (nbr_keep && nbc <= kNbrKeep) {  // the first loop's list is complete: no second search (positions have not moved since)
        const u32 *list = nbr_keep + (size_t)i * kNbrKeep;
        for (u32 k = lane; k < nbc; k += 32) {
            const u32 j = __ldcg(list + k);
            if (j != i) visit(j);
        }
    } else {
        for_each_neighbor_2d(i, lane, ring, pi, ep, imass, n, false, nullptr, visit);
    }
    dx = warp_sum_d(dx); dy = warp_sum_d(dy);
    if (K.gas) { fvx = warp_sum_d(fvx); fvy = warp_sum_d(fvy); }
    if (lane != 0) return;
    const double div = (double)nbc + (double)counts[i];
    delta[i] = make_double2((dx / p0) / div, (dy / p0) / div);
    if (K.gas) {
        double2 fi = f[i];
        fi.x += fvx; fi.y += fvy;
        f[i] = fi;
    }
}
__global__ void __launch_bounds__(kBlock) k2d_fluid_delta(const double2 *__restrict__ ep, const double *__restrict__ imass, const int *__restrict__ group, u32 n,
                                                          int op, double p0, FluidConsts K, const double *__restrict__ lambda, const u32 *__restrict__ nbcount,
                                                          const u32 *nbr_keep, const u32 *__restrict__ counts, double2 *__restrict__ delta,
                                                          const double2 *__restrict__ v, double2 *__restrict__ f) {
    __shared__ u32 rings[kBlock / 32][kRing];
    const u32 i = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (i >= n || group[i] != op) return;  // warp-uniform
    d_fluid_delta(i, threadIdx.x & 31, rings[threadIdx.x >> 5], ep, imass, n, p0, K, lambda, nbcount, nbr_keep, counts, delta, v, f);
}
__device__ __forceinline__ void d_fluid_apply(u32 i, double2 *ep, const double2 *delta) {
    double2 e = ep[i];
    const double2 d = __ldcg(delta + i);
    e.x += d.x; e.y += d.y;
    ep[i] = e;
}
__global__ void k2d_fluid_apply(double2 *__restrict__ ep, const double2 *__restrict__ delta, const int *__restrict__ group, u32 n, int op) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n || group[i] != op) return;
    d_fluid_apply(i, ep, delta);
}

// (23)-(27): v = (ep - p) / dt; particles that moved less than EPSILON sleep (particle.h:60-65)
__device__ __forceinline__ void d_finish(u32 i, double2 *p, double2 *v, const double2 *ep, double dt) {
    const double2 pi = p[i], e = ep[i];
    const double dx = e.x - pi.x, dy = e.y - pi.y;
    if (sqrt(dx * dx + dy * dy) < kEps) { v[i] = make_double2(0., 0.); return; }
    v[i] = make_double2(dx / dt, dy / dt);
    p[i] = e;
}
__global__ void k2d_finish(double2 *__restrict__ p, double2 *__restrict__ v, const double2 *__restrict__ ep, u32 n, double dt) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) d_finish(i, p, v, ep, dt);
}

// Simulation::mousePressed (simulation.cpp:1305-1314): every particle gets a velocity impulse of 7 towards the point
__global__ void k2d_impulse(double2 *__restrict__ v, const double2 *__restrict__ p, u32 n, double px, double py) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const double2 q = p[i];
    const double dx = px - q.x, dy = py - q.y;
    const double inv = 1. / sqrt(dx * dx + dy * dy);  // glm::normalize
    double2 vi = v[i];
    vi.x += 7. * (dx * inv); vi.y += 7. * (dy * inv);
    v[i] = vi;
}

enum StdKind { STD_FLUID, STD_GAS, STD_DISTANCE };

// ---- the whole tick as ONE kernel, for scenes of up to a few hundred particles ----------------------------------------
// A tick of such a scene is 8-29 dependent launches whose kernels run for a few microseconds each: even replayed as a graph, the
// gaps between dependent nodes (2-3 us each) are a large part of the tick.  k2d_tick_fused runs the same device functions in the same
// order inside one thread-block cluster (see the kernel), with a barrier where the launch boundaries were, the predicted positions and
// the contact cursors in shared memory for the whole tick, and the STANDARD list as a device-side op table.  Same functions, same
// arithmetic: bit-identical to the launch sequence (tests/test_gpu_2d_fused.py).  It stops paying where the all-pairs loops of the
// fluid / contact search want more SMs than a cluster has: above kFusedTickMaxN particles the launch sequence (148 SMs) is used.
struct FusedOp {  // one entry of the STANDARD list, distance constraints as whole runs
    u32 kind;      // STD_FLUID / STD_GAS / STD_DISTANCE
    u32 open;      // gas: open boundary
    u32 keep;      // fluid with an emitter: lambda of the last iteration is kept for the emitter's host logic
    u32 index;     // position in the STANDARD list (= the value of group[] of its members)
    double p0;
    u32 level_first, levels, width, pad;  // distance run
};
struct FusedTick {
    double2 *v, *ep, *f, *p, *delta;
    double *tmass, *lambda, *lambda_keep;
    const double *imass, *sfric, *kfric, *sdf_dist;
    const double2 *sdf_grad, *rs;
    const int *phase, *bod, *group, *raw;
    const u32 *static_counts;
    u32 *nb, *cnt, *flags, *counts, *draws, *rank, *lvl, *nbcount, *nbr_keep, *scalars, *scalars_out;
    unsigned char *nbq;
    const u32 *b_first, *b_count;
    const double *b_imass, *b_stiff;
    double2 *b_center;
    double *b_angle;
    const u32 *dc_i1, *dc_i2, *dc_level_off;
    const double *dc_rest;
    const FusedOp *ops;
    u32 n, nops, nbodies, ndist, window_base, stabilization_iterations, solver_iterations;
    int any_solid;
    unsigned long long *prof;  // PS2D_FUSED_PROFILE: cycles per phase, accumulated over the ticks (thread 0's clock)
    double dt, gx, gy, x0, x1, y0, y1;
};
// out of line: the chains' registers are then allocated for the chain alone, not for the whole tick
__device__ __noinline__ void fused_distance_warp(double2 *ep, const DistSlot *slots, const u32 *level_off, u32 levels, u32 lane) {
    d_distance_levels<true>(ep, slots, 0u, level_off, levels, lane, 32u);
}
__device__ __noinline__ void fused_distance_cta(double2 *ep, const DistSlot *slots, const u32 *level_off, u32 levels) {
    d_distance_levels<false>(ep, slots, 0u, level_off, levels, threadIdx.x, blockDim.x);
}
constexpr u32 kFusedTickMaxN = 800;   // measured break-even between 765 and 931 particles (16-CTA cluster): profiles/r2zz_2d_fused_tick.txt
constexpr u32 kFusedCluster = 16;   // CTAs of the one cluster wanted (8 is the portable maximum; 16 is asked for and halved until the device accepts)
constexpr u32 kFusedClusterMax = 16;
constexpr u32 kFusedBlock = 512;      // registers: 128 per thread, the serial chains of a tick must not spill
constexpr u32 kFusedTickCapN = 2048;
constexpr size_t kFusedStageMax = 160 * 1024;  // shared memory of the fused tick: 64 B per distance constraint + 36 B per particle

// Launched as ONE thread-block cluster of up to kFusedCluster CTAs: CTA 0 runs the tick; the others join it for the all-pairs loops (contact
// search, fluid lambda and delta), which are double-precision throughput on one SM otherwise (2 sqrt + 4 divides per pair in range).
// They read the predicted positions from the copy CTA 0 publishes in global memory and meet CTA 0 at cluster barriers (release /
// acquire at cluster scope: what one side wrote to global memory before the barrier the other side reads after it).
__global__ void __launch_bounds__(kFusedBlock, 1) k2d_tick_fused(const FusedTick A) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char fused_stage[];
    DistSlot *s_slots = reinterpret_cast<DistSlot *>(fused_stage);           // ndist: every distance constraint of the STANDARD list
    double2 *s_ep = reinterpret_cast<double2 *>(s_slots + A.ndist);          // n
    __shared__ u32 s_ring[kFusedBlock / 32][kRing];                          // the warps' neighbour rings (d_fluid_lambda / d_fluid_delta)
    const u32 rank = cluster.block_rank(), nranks = cluster.num_blocks();
    const bool lead = rank == 0;
    // every CTA keeps its own copies, in shared memory, of what the all-pairs loops read: inverse masses, phases and groups (constant
    // during a tick) and the predicted positions — CTA 0's are the tick's state, the others refresh theirs from the copy CTA 0
    // publishes in global memory before each cluster barrier that opens an all-pairs phase (one coalesced burst per phase instead of
    // dependent L2 misses per trip: cluster.sync flushes L1.  Reading CTA 0's shared memory directly — distributed shared memory —
    // makes its one shared-memory port serve n x n x 16 bytes per phase: measured 77 k cycles per lambda pass at 432 particles)
    double *c_imass = reinterpret_cast<double *>(s_ep + A.n);                // n (before s_cur: 8-byte aligned)
    u32 *s_cur = reinterpret_cast<u32 *>(c_imass + A.n);                     // n: the contact cursors (CTA 0)
    int *c_phase = reinterpret_cast<int *>(s_cur + A.n);                     // n
    int *c_group = c_phase + A.n;                                            // n
    const double2 *all_ep = s_ep;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = kFusedBlock / 32, n = A.n;
    const u32 cwarp = rank * nwarps + warp, cwarps = nranks * nwarps;        // this warp among the cluster's
    long long t_mark = A.prof ? clock64() : 0;
#define PS2D_PHASE(k) do { if (A.prof && lead && tid == 0) { const long long t_ = clock64(); A.prof[k] += (unsigned long long)(t_ - t_mark); t_mark = t_; } } while (0)
    for (u32 i = tid; i < n; i += kFusedBlock) { c_imass[i] = A.imass[i]; c_phase[i] = A.phase[i]; c_group[i] = A.group[i]; }
    if (lead)
        for (u32 i = tid; i < n; i += kFusedBlock) {
            d_predict(i, A.v, s_ep, A.f, A.tmass, A.p, A.imass, A.phase, A.dt, A.gx, A.gy);
            A.ep[i] = s_ep[i];
        }
    cluster.sync();
    if (!lead)
        for (u32 i = tid; i < n; i += kFusedBlock) s_ep[i] = A.ep[i];
    __syncthreads();
    for (u32 i = cwarp; i < n; i += cwarps)
        d_find_contacts(i, lane, all_ep, c_imass, c_phase, A.bod, A.static_counts, n, A.x0, A.x1, A.y0, A.y1, A.nb, A.cnt, A.flags, A.counts, A.draws, A.scalars + 3,
                        A.any_solid);
    cluster.sync();
    const Particles2D V{s_ep, A.p, A.tmass, A.sfric, A.kfric, A.phase, A.bod, A.counts, A.sdf_grad, A.sdf_dist, A.b_angle};
    if (lead) {
        d_scan_counts(A.draws, A.rank, n, A.scalars);
        __syncthreads();
        d_contact_levels(A.nb, A.cnt, A.flags, n, A.nbq, A.lvl, A.scalars + 1);
        d_distance_prepare(s_slots, 0u, A.ndist, A.imass, A.counts, A.dc_i1, A.dc_i2, A.dc_rest, tid, kFusedBlock);  // counts are final since the contact search
        __syncthreads();
        PS2D_PHASE(0);
        for (u32 st = 0; st < A.stabilization_iterations; st++) {
            d_contact_project(V, A.nb, A.cnt, A.flags, A.lvl, A.rank, A.scalars + 1, s_cur, n, A.raw, A.window_base, st, A.x0, A.x1, A.y0, A.y1, true);
            __syncthreads();
        }
    }
    for (u32 it = 0; it < A.solver_iterations; it++) {
        if (lead) {
            d_contact_project(V, A.nb, A.cnt, A.flags, A.lvl, A.rank, A.scalars + 1, s_cur, n, A.raw, A.window_base, A.stabilization_iterations + it, A.x0, A.x1,
                              A.y0, A.y1, false);
            __syncthreads();
            PS2D_PHASE(1);
        }
        for (u32 o = 0; o < A.nops; o++) {
            const FusedOp op = A.ops[o];
            if (op.kind == STD_DISTANCE) {
                if (lead) {
                    const u32 *lo = A.dc_level_off + op.level_first;
                    if (op.width <= 32u) {
                        if (warp == 0) fused_distance_warp(s_ep, s_slots, lo, op.levels, lane);
                    } else {
                        fused_distance_cta(s_ep, s_slots, lo, op.levels);
                    }
                    __syncthreads();
                    PS2D_PHASE(2);
                }
                continue;
            }
            const FluidConsts K = op.kind == STD_GAS ? FluidConsts{.2, .25, .5, 1, (int)op.open} : FluidConsts{.1, .2, 0., 0, 0};
            if (lead)
                for (u32 i = tid; i < n; i += kFusedBlock) A.ep[i] = s_ep[i];
            cluster.sync();  // CTA 0's positions are those this constraint sees
            if (!lead) {
                for (u32 i = tid; i < n; i += kFusedBlock) s_ep[i] = A.ep[i];
                __syncthreads();
            }
            for (u32 i = cwarp; i < n; i += cwarps) d_fluid_lambda(i, lane, s_ring[warp], all_ep, c_imass, c_phase, c_group, n, (int)op.index, op.p0, K, A.lambda, A.nbcount, A.nbr_keep, A.v, A.f);
            cluster.sync();  // every lambda is written
            PS2D_PHASE(3);
            if (lead && op.keep && it + 1 == A.solver_iterations)
                for (u32 i = tid; i < n; i += kFusedBlock) A.lambda_keep[i] = __ldcg(A.lambda + i);
            for (u32 i = cwarp; i < n; i += cwarps)
                if (c_group[i] == (int)op.index) d_fluid_delta(i, lane, s_ring[warp], all_ep, c_imass, n, op.p0, K, A.lambda, A.nbcount, A.nbr_keep, A.counts, A.delta, A.v, A.f);
            cluster.sync();  // every delta is written, nobody reads the positions any more
            if (lead) {
                for (u32 i = tid; i < n; i += kFusedBlock)
                    if (c_group[i] == (int)op.index) d_fluid_apply(i, s_ep, A.delta);
                __syncthreads();
                PS2D_PHASE(4);
            }
        }
        if (lead && A.nbodies) {
            for (u32 b = tid; b < A.nbodies; b += kFusedBlock) d_shape(b, s_ep, A.imass, A.rs, A.b_first, A.b_count, A.b_imass, A.b_stiff, A.b_center, A.b_angle);
            __syncthreads();
            PS2D_PHASE(5);
        }
    }
    if (!lead) return;
    for (u32 i = tid; i < n; i += kFusedBlock) {
        A.ep[i] = s_ep[i];
        d_finish(i, A.p, A.v, s_ep, A.dt);
    }
    if (tid == 0) {  // the tick's four scalars, straight into the host's (mapped, pinned) words
        for (int k = 0; k < 4; k++) A.scalars_out[k] = A.scalars[k];  // visible to the host when the kernel has ended (ps2d_tick waits for the stream)
    }
    PS2D_PHASE(6);
#undef PS2D_PHASE
}

// New particles arrive as one record each in a pinned staging buffer and are scattered to the arrays by one small kernel: an emitter
// that adds a particle every few ticks costs one asynchronous copy and one launch, not sixteen synchronous uploads.
struct AppendRec { double px, py, vx, vy, im, sf, kf; int phase, bod, group, pad; };
__global__ void k2d_append(const AppendRec *__restrict__ rec, u32 at, u32 n, double2 *p, double2 *ep, double2 *v, double2 *f, double2 *rs, double2 *sdf_grad,
                           double *sdf_dist, double *imass, double *tmass, double *sfric, double *kfric, double *lambda, int *phase, int *bod, int *group,
                           u32 *static_counts) {
    const u32 k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= n) return;
    const AppendRec r = rec[k];
    const u32 i = at + k;
    p[i] = ep[i] = make_double2(r.px, r.py);
    v[i] = make_double2(r.vx, r.vy);
    f[i] = rs[i] = sdf_grad[i] = make_double2(0., 0.);
    sdf_dist[i] = -1.;
    imass[i] = tmass[i] = r.im;
    sfric[i] = r.sf; kfric[i] = r.kf;
    lambda[i] = 0.;
    phase[i] = r.phase; bod[i] = r.bod; group[i] = r.group;
    static_counts[i] = 0u;
}

// glibc rand() = random(), TYPE_3: r[i] = r[i-31] + r[i-3], output r[i] >> 1 (after 310 discarded words)
struct GlibcRand {
    std::vector<uint32_t> r;
    uint64_t calls = 0;
    void seed(uint32_t s) {
        r.assign(344, 0);
        r[0] = s ? s : 1;
        for (int i = 1; i < 31; i++) {
            const long hi = (long)r[i - 1] / 127773, lo = (long)r[i - 1] % 127773;
            long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            r[i] = (uint32_t)w;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
        r.erase(r.begin(), r.end() - 31);
        calls = 0;
    }
    int next() {
        const uint32_t v = r[0] + r[28];
        r.erase(r.begin());
        r.push_back(v);
        calls++;
        return (int)(v >> 1);
    }
};

struct StdOp {  // one entry of m_globalConstraints[STANDARD]
    StdKind kind;
    double p0 = 0.;  // fluid / gas rest density
    int open = 0;
    u32 i1 = 0, i2 = 0;
    double d = 0.;
};
struct DistanceRun {  // consecutive distance constraints [begin, end) of the STANDARD list, as uploaded level by level
    size_t begin, end;
    u32 dev_first, dev_level_first, levels;
    u32 width;  // constraints in the widest level
};
struct Emitter { double x, y, rate, timer; u32 standard_index; };
struct FluidEmitterRec { double x, y, rate, timer, total_timer; u32 standard_index; };
}  // namespace

struct Ps2dCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Ps2dParams params{};
    uint64_t cap = 0;
    u32 n = 0;
    double2 *p = nullptr, *v = nullptr, *ep = nullptr, *f = nullptr, *delta = nullptr, *rs = nullptr, *sdf_grad = nullptr;
    double *imass = nullptr, *tmass = nullptr, *sfric = nullptr, *kfric = nullptr, *lambda = nullptr, *sdf_dist = nullptr;
    int *phase = nullptr, *bod = nullptr, *group = nullptr, *raw = nullptr;
    u32 *static_counts = nullptr, *flags = nullptr, *counts = nullptr, *draws = nullptr, *rank = nullptr, *nbcount = nullptr, *nb = nullptr, *cnt = nullptr,
        *lvl = nullptr, *cur = nullptr, *scalars = nullptr, *scalars_host = nullptr;  // scalars: [0] draws, [1] levels, [2] constraints, [3] overflow
    // the tick's launch sequence as a CUDA graph (a tick of the reference's scenes is 11-29 dependent launches of a few
    // microseconds each: issued eagerly the host is the bottleneck).  Re-captured when anything its nodes bake in changes.
    u32 *window_base_dev = nullptr;
    cudaGraphExec_t tick_graph = nullptr;
    struct TickKey { uint64_t n; double dt; uint64_t standard_version; Ps2dParams params; const void *bodies, *raw, *lambda_keep; uint64_t misc; } tick_key{};
    uint64_t standard_version = 0;
    u32 tick_key_age = 0;  // ticks the current key has been seen
    unsigned char *nbq = nullptr;
    // rigid bodies
    u32 nbodies = 0, bodies_cap = 0;
    u32 *b_first = nullptr, *b_count = nullptr;
    double *b_imass = nullptr, *b_stiff = nullptr, *b_angle = nullptr;
    double2 *b_center = nullptr;
    // distance constraints, level-sorted per run
    u32 *dc_i1 = nullptr, *dc_i2 = nullptr, *dc_level_off = nullptr;
    double *dc_rest = nullptr;
    size_t dc_cap = 0, dc_level_cap = 0;
    std::vector<DistanceRun> runs;
    FusedOp *fused_ops = nullptr;      // the STANDARD list for k2d_tick_fused (rebuild_standard)
    size_t fused_ops_cap = 0;
    u32 fused_nops = 0;
    u32 dc_total = 0;                  // distance constraints in the STANDARD list
    u32 *scalars_out = nullptr;        // scalars_host as the device sees it (mapped pinned memory), or null
    u32 *nbr_keep = nullptr;           // kNbrKeep neighbour indices per particle, first loop -> second loop of a fluid / gas constraint (null: max_particles too large)
    bool last_tick_fused = false;
    AppendRec *append_host = nullptr, *append_dev = nullptr;  // staging of ps2d_add_particles
    size_t append_cap = 0;
    cudaEvent_t append_done = nullptr;
    bool append_pending = false;
    unsigned long long *fused_prof = nullptr;  // PS2D_FUSED_PROFILE=1: per-phase cycles of k2d_tick_fused, printed by ps2d_destroy
    uint64_t fused_ticks = 0;
    bool standard_dirty = true;
    std::vector<StdOp> standard;
    std::vector<u32> h_static_counts;
    std::vector<double> h_imass;  // host mirror: constraint constructors validate against it
    std::vector<int> h_phase;
    std::vector<Emitter> emitters;
    std::vector<FluidEmitterRec> fluid_emitters;
    std::vector<int> h_group;   // host mirror of `group` (the FluidEmitter walks its fluid's member list)
    std::vector<double> h_t;    // Particle::t (particle.h:38): freeze countdown, only the FluidEmitter reads it
    double *lambda_keep = nullptr;  // lambda of the fluid emitters' constraints after the last solver iteration
    // look-ahead window of the rand() stream on the device: raw[k] = draw number win_pos + k of the stream (counted like
    // rng.calls).  A tick indexes it with its own (device-side) constraint count, so the host learns how many draws were
    // consumed only at the end of the tick and no round trip is needed in the middle of it.
    size_t raw_cap = 0;
    uint64_t win_pos = 0, win_len = 0;
    std::vector<int> h_raw;
    GlibcRand rng;
    int any_solid = 0, any_jitter = 0;
    u32 last_num_boundary = 0, last_contacts = 0, last_levels = 0, launches = 0;
};

#define CU2(x)                                                                                        \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            ps_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return PS_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

extern "C" void ps2d_default_params(Ps2dParams *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->x_bounds[0] = -8.; p->x_bounds[1] = 8.;    // scene 6, cpu/src/simulation.cpp:898-899
    p->y_bounds[0] = -8.; p->y_bounds[1] = 40.;
    p->gravity[0] = 0.; p->gravity[1] = -9.8;
    p->solver_iterations = 3;                     // simulation.h:11
    p->stabilization_iterations = 0;              // USE_STABILIZATION is commented out in the reference's build (simulation.h:17)
}

template <class T>
static bool dev_alloc(T **ptr, size_t count) { return cudaMalloc((void **)ptr, std::max<size_t>(count, 1) * sizeof(T)) == cudaSuccess; }

extern "C" int ps2d_create(int device, const Ps2dParams *params, uint64_t max_particles, Ps2dCtx **out) {
    if (!params || !out || !max_particles) { ps_set_error("ps2d_create: bad argument"); return PS_ERR_INVALID; }
    *out = nullptr;
    if (!(params->x_bounds[0] < params->x_bounds[1]) || !(params->y_bounds[0] < params->y_bounds[1]) || params->solver_iterations > 64 || params->stabilization_iterations > 64) {
        ps_set_error("ps2d_create: bad bounds or iteration count"); return PS_ERR_INVALID;
    }
    if (max_particles > (1u << 28)) { ps_set_error("ps2d_create: max_particles too large"); return PS_ERR_INVALID; }
    int ndev = 0;
    CU2(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { ps_set_error("ps2d_create: device %d of %d", device, ndev); return PS_ERR_INVALID; }
    cudaDeviceProp prop;
    CU2(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { ps_set_error("ps2d_create: device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor); return PS_ERR_CUDA; }
    CU2(cudaSetDevice(device));
    Ps2dCtx *c = new Ps2dCtx();
    c->device = device; c->params = *params; c->cap = max_particles;
    c->rng.seed(1);
    const size_t n = max_particles;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && dev_alloc(&c->p, n) && dev_alloc(&c->v, n) && dev_alloc(&c->ep, n) && dev_alloc(&c->f, n) && dev_alloc(&c->delta, n) && dev_alloc(&c->rs, n) &&
         dev_alloc(&c->sdf_grad, n) && dev_alloc(&c->imass, n) && dev_alloc(&c->tmass, n) && dev_alloc(&c->sfric, n) && dev_alloc(&c->kfric, n) &&
         dev_alloc(&c->lambda, n) && dev_alloc(&c->sdf_dist, n) && dev_alloc(&c->phase, n) && dev_alloc(&c->bod, n) && dev_alloc(&c->group, n) &&
         dev_alloc(&c->static_counts, n) && dev_alloc(&c->flags, n) && dev_alloc(&c->counts, n) && dev_alloc(&c->draws, n) && dev_alloc(&c->rank, n) &&
         dev_alloc(&c->nbcount, n) && dev_alloc(&c->nb, n * kMaxC) && dev_alloc(&c->cnt, n) && dev_alloc(&c->lvl, n * kEntries) && dev_alloc(&c->cur, n) &&
         dev_alloc(&c->nbq, n * kMaxC) && dev_alloc(&c->scalars, 8) && cudaMallocHost((void **)&c->scalars_host, 32) == cudaSuccess;
    if (ok) c->window_base_dev = c->scalars + 4;
    if (ok && n <= (1u << 20) && !dev_alloc(&c->nbr_keep, n * kNbrKeep)) { c->nbr_keep = nullptr; cudaGetLastError(); }  // optional: 512 B per particle
    if (ok && getenv("PS2D_FUSED_PROFILE")) ok = dev_alloc(&c->fused_prof, 8) && cudaMemset(c->fused_prof, 0, 64) == cudaSuccess;
    if (ok && cudaHostGetDevicePointer((void **)&c->scalars_out, c->scalars_host, 0) != cudaSuccess) { c->scalars_out = nullptr; cudaGetLastError(); }
    if (ok) ok = cudaMemsetAsync(c->f, 0, n * 16, c->stream) == cudaSuccess && cudaMemsetAsync(c->lambda, 0, n * 8, c->stream) == cudaSuccess &&
                 cudaMemsetAsync(c->scalars, 0, 16, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (!ok) {
        ps_set_error("ps2d_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        ps2d_destroy(c);
        return PS_ERR_CUDA;
    }
    *out = c;
    return PS_OK;
}

extern "C" int ps2d_destroy(Ps2dCtx *c) {
    if (!c) return PS_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->append_host) cudaFreeHost(c->append_host);
    if (c->append_dev) cudaFree(c->append_dev);
    if (c->append_done) cudaEventDestroy(c->append_done);
    if (c->fused_prof && c->fused_ticks) {
        unsigned long long h[8] = {};
        cudaMemcpy(h, c->fused_prof, sizeof h, cudaMemcpyDeviceToHost);
        const double t = (double)c->fused_ticks;
        fprintf(stderr, "ps2d fused tick, cycles per tick over %.0f ticks: set-up %.0f, contacts %.0f, distance runs %.0f, fluid lambda %.0f, fluid delta+apply %.0f, shape %.0f, finish %.0f\n",
                t, h[0] / t, h[1] / t, h[2] / t, h[3] / t, h[4] / t, h[5] / t, h[6] / t);
    }
    if (c->fused_prof) cudaFree(c->fused_prof);
    void *ptrs[] = {c->p, c->v, c->ep, c->f, c->delta, c->rs, c->sdf_grad, c->imass, c->tmass, c->sfric, c->kfric, c->lambda, c->sdf_dist, c->phase, c->bod,
                    c->group, c->raw, c->static_counts, c->flags, c->counts, c->draws, c->rank, c->nbcount, c->nb, c->cnt, c->lvl, c->cur, c->nbq, c->scalars,
                    c->b_first, c->b_count, c->b_imass, c->b_stiff, c->b_angle, c->b_center, c->dc_i1, c->dc_i2, c->dc_level_off, c->dc_rest, c->lambda_keep, c->fused_ops, c->nbr_keep};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (c->scalars_host) cudaFreeHost(c->scalars_host);
    if (c->tick_graph) cudaGraphExecDestroy(c->tick_graph);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PS_OK;
}

template <class T>
static cudaError_t upload(Ps2dCtx *c, T *dst, const T *src, size_t count) {
    if (!count) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(c->stream);  // callers pass temporaries
}

static int append_particles(Ps2dCtx *c, const double *p2, const double *v2, const double *inv_mass, const int32_t *phase, const int32_t *bod,
                            const double *s_friction, const double *k_friction, const int32_t *group, uint64_t n, uint64_t *first) {
    if (!c || !p2 || !inv_mass || !phase) { ps_set_error("ps2d_add_particles: null argument"); return PS_ERR_INVALID; }
    if (c->n + n > c->cap) { ps_set_error("ps2d_add_particles: %llu + %llu exceeds max_particles", (unsigned long long)c->n, (unsigned long long)n); return PS_ERR_CAPACITY; }
    for (uint64_t k = 0; k < n; k++) {
        if (phase[k] < PS2D_PHASE_SOLID || phase[k] > PS2D_PHASE_GAS) { ps_set_error("ps2d_add_particles: unknown phase %d", phase[k]); return PS_ERR_INVALID; }
        if (!(inv_mass[k] >= 0.)) { ps_set_error("ps2d_add_particles: negative inverse mass"); return PS_ERR_INVALID; }
    }
    const u32 at = c->n;
    if (first) *first = at;
    if (!n) return PS_OK;
    CU2(cudaSetDevice(c->device));
    if (!c->append_done) CU2(cudaEventCreateWithFlags(&c->append_done, cudaEventDisableTiming));
    if (c->append_pending) { CU2(cudaEventSynchronize(c->append_done)); c->append_pending = false; }  // the staging buffer is free again
    if (n > c->append_cap) {
        CU2(cudaStreamSynchronize(c->stream));
        if (c->append_host) cudaFreeHost(c->append_host);
        if (c->append_dev) cudaFree(c->append_dev);
        c->append_host = nullptr; c->append_dev = nullptr;
        c->append_cap = std::max<size_t>(n, 64);
        if (cudaMallocHost((void **)&c->append_host, c->append_cap * sizeof(AppendRec)) != cudaSuccess || !dev_alloc(&c->append_dev, c->append_cap)) {
            c->append_cap = 0;
            ps_set_error("ps2d_add_particles: staging allocation failed"); return PS_ERR_CUDA;
        }
    }
    for (uint64_t k = 0; k < n; k++) {
        AppendRec &r = c->append_host[k];
        r.px = p2[2 * k]; r.py = p2[2 * k + 1];
        r.vx = v2 ? v2[2 * k] : 0.; r.vy = v2 ? v2[2 * k + 1] : 0.;
        r.im = inv_mass[k];
        r.sf = s_friction ? s_friction[k] : 0.; r.kf = k_friction ? k_friction[k] : 0.;
        r.phase = phase[k]; r.bod = bod ? bod[k] : -1; r.group = group ? group[k] : -1; r.pad = 0;
    }
    CU2(cudaMemcpyAsync(c->append_dev, c->append_host, n * sizeof(AppendRec), cudaMemcpyHostToDevice, c->stream));
    k2d_append<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, c->stream>>>(c->append_dev, at, (u32)n, c->p, c->ep, c->v, c->f, c->rs, c->sdf_grad, c->sdf_dist, c->imass,
                                                                                 c->tmass, c->sfric, c->kfric, c->lambda, c->phase, c->bod, c->group, c->static_counts);
    CU2(cudaGetLastError());
    CU2(cudaEventRecord(c->append_done, c->stream));
    c->append_pending = true;
    for (uint64_t k = 0; k < n; k++) {
        c->h_imass.push_back(inv_mass[k]);
        c->h_phase.push_back(phase[k]);
        c->h_static_counts.push_back(0);
        c->h_group.push_back(group ? group[k] : -1);
        c->h_t.push_back(4.);
        if (phase[k] == PS2D_PHASE_SOLID) c->any_solid = 1; else c->any_jitter = 1;
    }
    c->n += (u32)n;
    return PS_OK;
}
extern "C" int ps2d_add_particles(Ps2dCtx *c, const double *p2, const double *v2, const double *inv_mass, const int32_t *phase, const int32_t *bod,
                                  const double *s_friction, const double *k_friction, uint64_t n, uint64_t *first) {
    return append_particles(c, p2, v2, inv_mass, phase, bod, s_friction, k_friction, nullptr, n, first);
}

static int bump_static_counts(Ps2dCtx *c, u32 i) {
    c->h_static_counts[i]++;
    CU2(upload(c, c->static_counts + i, &c->h_static_counts[i], 1));
    return PS_OK;
}

extern "C" int ps2d_add_distance_constraint(Ps2dCtx *c, uint32_t i1, uint32_t i2, double d) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (i1 >= c->n || i2 >= c->n || i1 == i2) { ps_set_error("ps2d_add_distance_constraint: bad particle indices %u, %u", i1, i2); return PS_ERR_INVALID; }
    CU2(cudaSetDevice(c->device));
    if (d < 0.) {  // DistanceConstraint(first, second, particles): d = length(p1 - p2)
        double a[2], b[2];
        CU2(cudaStreamSynchronize(c->stream));  // particles are appended on the context's stream
        CU2(cudaMemcpy(a, c->p + i1, 16, cudaMemcpyDeviceToHost));
        CU2(cudaMemcpy(b, c->p + i2, 16, cudaMemcpyDeviceToHost));
        const double x = a[0] - b[0], y = a[1] - b[1];
        d = std::sqrt(x * x + y * y);
    }
    StdOp op;
    op.kind = STD_DISTANCE; op.i1 = i1; op.i2 = i2; op.d = d;
    c->standard.push_back(op);
    c->standard_dirty = true;
    int r = bump_static_counts(c, i1);
    if (r == PS_OK) r = bump_static_counts(c, i2);
    return r;
}

static int add_group(Ps2dCtx *c, const uint32_t *indices, uint64_t n, double density, int want_phase, int open, uint32_t *standard_index, const char *who) {
    if (!c || (!indices && n)) { ps_set_error("%s: null argument", who); return PS_ERR_INVALID; }
    if (!(density > 0.)) { ps_set_error("%s: density must be positive", who); return PS_ERR_INVALID; }
    for (uint64_t k = 0; k < n; k++) {
        if (indices[k] >= c->n) { ps_set_error("%s: particle index %u out of range", who, indices[k]); return PS_ERR_INVALID; }
        if (c->h_phase[indices[k]] != want_phase) { ps_set_error("%s: particle %u has phase %d", who, indices[k], c->h_phase[indices[k]]); return PS_ERR_INVALID; }
        if (c->h_imass[indices[k]] == 0.) { ps_set_error("A fluid cannot have a point of infinite mass."); return PS_ERR_INVALID; }  // simulation.cpp:419-422,441-444
    }
    CU2(cudaSetDevice(c->device));
    const int id = (int)c->standard.size();
    for (uint64_t k = 0; k < n; k++) { CU2(upload(c, c->group + indices[k], &id, 1)); c->h_group[indices[k]] = id; }
    StdOp op;
    op.kind = want_phase == PS2D_PHASE_GAS ? STD_GAS : STD_FLUID; op.p0 = density; op.open = open;
    c->standard.push_back(op);
    c->standard_dirty = true;
    if (standard_index) *standard_index = (u32)id;
    return PS_OK;
}
extern "C" int ps2d_add_fluid_constraint(Ps2dCtx *c, const uint32_t *indices, uint64_t n, double density, uint32_t *standard_index) {
    return add_group(c, indices, n, density, PS2D_PHASE_FLUID, 0, standard_index, "ps2d_add_fluid_constraint");
}
extern "C" int ps2d_add_gas_constraint(Ps2dCtx *c, const uint32_t *indices, uint64_t n, double density, int open, uint32_t *standard_index) {
    return add_group(c, indices, n, density, PS2D_PHASE_GAS, open, standard_index, "ps2d_add_gas_constraint");
}

static int grow_bodies(Ps2dCtx *c) {
    if (c->nbodies < c->bodies_cap) return PS_OK;
    const u32 cap = std::max<u32>(64, c->bodies_cap * 2);
    u32 *bf = nullptr, *bc = nullptr;
    double *bi = nullptr, *bs = nullptr, *ba = nullptr;
    double2 *bcen = nullptr;
    if (!dev_alloc(&bf, cap) || !dev_alloc(&bc, cap) || !dev_alloc(&bi, cap) || !dev_alloc(&bs, cap) || !dev_alloc(&ba, cap) || !dev_alloc(&bcen, cap)) {
        ps_set_error("ps2d: body table allocation failed"); return PS_ERR_CUDA;
    }
    if (c->nbodies) {
        CU2(cudaMemcpy(bf, c->b_first, c->nbodies * 4, cudaMemcpyDeviceToDevice)); CU2(cudaMemcpy(bc, c->b_count, c->nbodies * 4, cudaMemcpyDeviceToDevice));
        CU2(cudaMemcpy(bi, c->b_imass, c->nbodies * 8, cudaMemcpyDeviceToDevice)); CU2(cudaMemcpy(bs, c->b_stiff, c->nbodies * 8, cudaMemcpyDeviceToDevice));
        CU2(cudaMemcpy(ba, c->b_angle, c->nbodies * 8, cudaMemcpyDeviceToDevice)); CU2(cudaMemcpy(bcen, c->b_center, c->nbodies * 16, cudaMemcpyDeviceToDevice));
    }
    void *old[] = {c->b_first, c->b_count, c->b_imass, c->b_stiff, c->b_angle, c->b_center};
    for (void *q : old) if (q) cudaFree(q);
    c->b_first = bf; c->b_count = bc; c->b_imass = bi; c->b_stiff = bs; c->b_angle = ba; c->b_center = bcen; c->bodies_cap = cap;
    return PS_OK;
}

extern "C" int ps2d_restore_rigid_body(Ps2dCtx *c, uint32_t first, uint32_t n, const double *rs2, const double *sdf3, double inv_mass, const double *center2,
                                       double angle, double stiffness, uint32_t *body_index) {
    if (!c || !rs2 || !sdf3 || !center2) { ps_set_error("ps2d_restore_rigid_body: null argument"); return PS_ERR_INVALID; }
    if (n <= 1) { ps_set_error("Rigid bodies must be at least 2 points."); return PS_ERR_INVALID; }  // simulation.cpp:371-374
    if ((uint64_t)first + n > c->n) { ps_set_error("ps2d_restore_rigid_body: particles [%u, %u) out of range", first, first + n); return PS_ERR_INVALID; }
    for (u32 k = 0; k < n; k++) {
        if (c->h_phase[first + k] != PS2D_PHASE_SOLID) { ps_set_error("ps2d_restore_rigid_body: particle %u is not SOLID", first + k); return PS_ERR_INVALID; }
        if (c->h_imass[first + k] == 0.) { ps_set_error("A rigid body cannot have a point of infinite mass."); return PS_ERR_INVALID; }  // simulation.cpp:386-389
    }
    CU2(cudaSetDevice(c->device));
    int r = grow_bodies(c);
    if (r != PS_OK) return r;
    const u32 b = c->nbodies;
    std::vector<double> grad(2 * (size_t)n), dist(n);
    std::vector<int> bods(n, (int)b);
    for (u32 k = 0; k < n; k++) { grad[2 * k] = sdf3[3 * k]; grad[2 * k + 1] = sdf3[3 * k + 1]; dist[k] = sdf3[3 * k + 2]; }
    CU2(upload(c, (double *)(c->rs + first), rs2, 2 * (size_t)n));
    CU2(upload(c, (double *)(c->sdf_grad + first), grad.data(), 2 * (size_t)n));
    CU2(upload(c, c->sdf_dist + first, dist.data(), n));
    CU2(upload(c, c->bod + first, bods.data(), n));
    CU2(upload(c, c->b_first + b, &first, 1));
    CU2(upload(c, c->b_count + b, &n, 1));
    CU2(upload(c, c->b_imass + b, &inv_mass, 1));
    CU2(upload(c, c->b_stiff + b, &stiffness, 1));
    CU2(upload(c, c->b_angle + b, &angle, 1));
    CU2(upload(c, (double *)(c->b_center + b), center2, 2));
    for (u32 k = 0; k < n; k++) c->h_static_counts[first + k]++;  // TotalShapeConstraint::updateCounts
    CU2(upload(c, c->static_counts + first, &c->h_static_counts[first], n));
    c->nbodies++;
    if (body_index) *body_index = b;
    return PS_OK;
}

extern "C" int ps2d_create_rigid_body(Ps2dCtx *c, const double *p2, const double *v2, const double *inv_mass, const double *s_friction,
                                      const double *k_friction, const double *sdf3, uint64_t n, uint32_t *body_index) {
    if (!c || !p2 || !inv_mass || !sdf3) { ps_set_error("ps2d_create_rigid_body: null argument"); return PS_ERR_INVALID; }
    if (n <= 1) { ps_set_error("Rigid bodies must be at least 2 points."); return PS_ERR_INVALID; }
    double total_mass = 0.;
    for (uint64_t k = 0; k < n; k++) {
        if (inv_mass[k] == 0.) { ps_set_error("A rigid body cannot have a point of infinite mass."); return PS_ERR_INVALID; }
        total_mass += 1.0 / inv_mass[k];
    }
    // body->imass = 1 / totalMass; updateCOM(particles, false): centre from p (the angle loop sees no r vectors yet: 0);
    // computeRs: r_i = p_i - centre, imass recomputed over the particles with r != 0 (simulation.cpp:398-402, particle.cpp:59-73)
    double bim = 1.0 / total_mass, tx = 0., ty = 0.;
    for (uint64_t k = 0; k < n; k++) { tx += p2[2 * k] / inv_mass[k]; ty += p2[2 * k + 1] / inv_mass[k]; }
    const double center[2] = {tx * bim, ty * bim};
    std::vector<double> rs(2 * n);
    double m = 0.;
    for (uint64_t k = 0; k < n; k++) {
        const double rx = p2[2 * k] - center[0], ry = p2[2 * k + 1] - center[1];
        rs[2 * k] = rx; rs[2 * k + 1] = ry;
        if (rx * rx + ry * ry != 0.) m += 1.0 / inv_mass[k];
    }
    bim = 1.0 / m;
    std::vector<int32_t> ph(n, PS2D_PHASE_SOLID);
    uint64_t first = 0;
    int r = ps2d_add_particles(c, p2, v2, inv_mass, ph.data(), nullptr, s_friction, k_friction, n, &first);
    if (r != PS_OK) return r;
    return ps2d_restore_rigid_body(c, (u32)first, (u32)n, rs.data(), sdf3, bim, center, 0., 1., body_index);
}

extern "C" int ps2d_create_fluid(Ps2dCtx *c, const double *p2, const double *v2, const double *inv_mass, uint64_t n, double density) {
    if (!c || !p2 || !v2 || !inv_mass) { ps_set_error("ps2d_create_fluid: null argument"); return PS_ERR_INVALID; }
    if (!(density > 0.)) { ps_set_error("ps2d_create_fluid: density must be positive"); return PS_ERR_INVALID; }
    for (uint64_t k = 0; k < n; k++)
        if (inv_mass[k] == 0.0) { ps_set_error("A fluid cannot have a point of infinite mass."); return PS_ERR_INVALID; }  // simulation.cpp:441-444
    std::vector<int32_t> ph(n, PS2D_PHASE_FLUID);
    uint64_t first = 0;
    int r = ps2d_add_particles(c, p2, v2, inv_mass, ph.data(), nullptr, nullptr, nullptr, n, &first);
    if (r != PS_OK) return r;
    std::vector<u32> idx(n);
    for (uint64_t k = 0; k < n; k++) idx[k] = (u32)(first + k);
    return ps2d_add_fluid_constraint(c, idx.data(), n, density, nullptr);
}

extern "C" int ps2d_create_gas(Ps2dCtx *c, const double *p2, const double *v2, const double *inv_mass, uint64_t n, double density, int open,
                               uint32_t *standard_index) {
    if (!c || !p2 || !inv_mass) { ps_set_error("ps2d_create_gas: null argument"); return PS_ERR_INVALID; }
    if (!(density > 0.)) { ps_set_error("ps2d_create_gas: density must be positive"); return PS_ERR_INVALID; }
    for (uint64_t k = 0; k < n; k++)
        if (inv_mass[k] == 0.0) { ps_set_error("A fluid cannot have a point of infinite mass."); return PS_ERR_INVALID; }  // simulation.cpp:419-422
    std::vector<int32_t> ph(n, PS2D_PHASE_GAS);
    uint64_t first = 0;
    int r = ps2d_add_particles(c, p2, v2, inv_mass, ph.data(), nullptr, nullptr, nullptr, n, &first);
    if (r != PS_OK) return r;
    std::vector<u32> idx(n);
    for (uint64_t k = 0; k < n; k++) idx[k] = (u32)(first + k);
    return ps2d_add_gas_constraint(c, idx.data(), n, density, open, standard_index);
}

extern "C" int ps2d_create_smoke_emitter(Ps2dCtx *c, const double *posn2, double rate, uint32_t standard_index, double timer) {
    if (!c || !posn2) { ps_set_error("ps2d_create_smoke_emitter: null argument"); return PS_ERR_INVALID; }
    if (!(rate > 0.)) { ps_set_error("ps2d_create_smoke_emitter: rate must be positive"); return PS_ERR_INVALID; }
    if (standard_index != UINT32_MAX && (standard_index >= c->standard.size() || c->standard[standard_index].kind != STD_GAS)) {
        ps_set_error("ps2d_create_smoke_emitter: STANDARD constraint %u is not a gas", standard_index); return PS_ERR_INVALID;
    }
    c->emitters.push_back(Emitter{posn2[0], posn2[1], rate, timer, standard_index});
    return PS_OK;
}

extern "C" int ps2d_mouse_pressed(Ps2dCtx *c, double x, double y) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!c->n) return PS_OK;
    CU2(cudaSetDevice(c->device));
    k2d_impulse<<<(c->n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->v, c->p, c->n, x, y);
    CU2(cudaStreamSynchronize(c->stream));
    return PS_OK;
}

extern "C" int ps2d_set_forces(Ps2dCtx *c, const double *f2) {
    if (!c || !f2) { ps_set_error("ps2d_set_forces: null argument"); return PS_ERR_INVALID; }
    CU2(cudaSetDevice(c->device));
    CU2(upload(c, (double *)c->f, f2, 2 * (size_t)c->n));
    return PS_OK;
}
extern "C" int ps2d_body_state(Ps2dCtx *c, uint32_t body, double *center2, double *angle) {
    if (!c || body >= c->nbodies) { ps_set_error("ps2d_body_state: no body %u", body); return PS_ERR_INVALID; }
    CU2(cudaSetDevice(c->device));
    CU2(cudaStreamSynchronize(c->stream));
    if (center2) CU2(cudaMemcpy(center2, c->b_center + body, 16, cudaMemcpyDeviceToHost));
    if (angle) CU2(cudaMemcpy(angle, c->b_angle + body, 8, cudaMemcpyDeviceToHost));
    return PS_OK;
}
extern "C" uint32_t ps2d_num_bodies(Ps2dCtx *c) { return c ? c->nbodies : 0; }

// Ps2dParams.stabilization_iterations of an existing context (the scene builders create theirs with the default, 0)
extern "C" int ps2d_set_stabilization_iterations(Ps2dCtx *c, uint32_t iterations) {
    if (!c || iterations > 64) { ps_set_error("ps2d_set_stabilization_iterations: bad argument"); return PS_ERR_INVALID; }
    c->params.stabilization_iterations = iterations;
    return PS_OK;
}

extern "C" int ps2d_seed_rand(Ps2dCtx *c, uint32_t seed, uint64_t skip) {
    if (!c) return PS_ERR_INVALID;
    c->rng.seed(seed);
    for (uint64_t k = 0; k < skip; k++) c->rng.next();
    c->win_len = 0;  // another stream: the look-ahead window on the device is stale
    return PS_OK;
}
extern "C" uint64_t ps2d_rand_calls(Ps2dCtx *c) { return c ? c->rng.calls : 0; }
// one draw of the context's rand() stream (scene builders jitter particles with it, like the reference's frand())
extern "C" int ps2d_rand(Ps2dCtx *c, int *out) {
    if (!c || !out) { ps_set_error("ps2d_rand: null argument"); return PS_ERR_INVALID; }
    *out = c->rng.next();
    return PS_OK;
}
extern "C" uint64_t ps2d_num_particles(Ps2dCtx *c) { return c ? c->n : 0; }
extern "C" uint32_t ps2d_last_num_boundary_constraints(Ps2dCtx *c) { return c ? c->last_num_boundary : 0; }
extern "C" uint32_t ps2d_last_num_contact_constraints(Ps2dCtx *c) { return c ? c->last_contacts : 0; }
extern "C" uint32_t ps2d_last_num_levels(Ps2dCtx *c) { return c ? c->last_levels : 0; }
extern "C" uint32_t ps2d_launches_per_tick(Ps2dCtx *c) { return c ? c->launches : 0; }

// Level schedule of the static distance constraints: every maximal run of consecutive distance constraints of the
// STANDARD list is sorted (stably) by level and uploaded with its level offsets.
static int rebuild_standard(Ps2dCtx *c) {
    c->runs.clear();
    std::vector<u32> i1, i2, level_off;
    std::vector<double> rest;
    std::vector<u32> last(c->n, 0);
    const size_t m = c->standard.size();
    size_t k = 0;
    while (k < m) {
        if (c->standard[k].kind != STD_DISTANCE) { k++; continue; }
        size_t e = k;
        while (e < m && c->standard[e].kind == STD_DISTANCE) e++;
        std::vector<u32> lv(e - k);
        u32 levels = 0;
        for (size_t q = k; q < e; q++) {
            const StdOp &o = c->standard[q];
            const u32 l = std::max(last[o.i1], last[o.i2]) + 1;
            last[o.i1] = last[o.i2] = l;
            lv[q - k] = l;
            levels = std::max(levels, l);
        }
        for (size_t q = k; q < e; q++) last[c->standard[q].i1] = last[c->standard[q].i2] = 0;
        DistanceRun run{k, e, (u32)i1.size(), (u32)level_off.size(), levels, 0u};
        for (u32 l = 1; l <= levels; l++) {
            level_off.push_back((u32)i1.size());
            for (size_t q = k; q < e; q++)
                if (lv[q - k] == l) { i1.push_back(c->standard[q].i1); i2.push_back(c->standard[q].i2); rest.push_back(c->standard[q].d); }
            run.width = std::max(run.width, (u32)i1.size() - level_off.back());
        }
        level_off.push_back((u32)i1.size());
        c->runs.push_back(run);
        k = e;
    }
    if (i1.size() > c->dc_cap) {
        for (void *q : {(void *)c->dc_i1, (void *)c->dc_i2, (void *)c->dc_rest}) if (q) cudaFree(q);
        c->dc_i1 = c->dc_i2 = nullptr; c->dc_rest = nullptr;
        c->dc_cap = i1.size() * 2;
        if (!dev_alloc(&c->dc_i1, c->dc_cap) || !dev_alloc(&c->dc_i2, c->dc_cap) || !dev_alloc(&c->dc_rest, c->dc_cap)) { ps_set_error("ps2d: distance table allocation failed"); return PS_ERR_CUDA; }
    }
    if (level_off.size() > c->dc_level_cap) {
        if (c->dc_level_off) cudaFree(c->dc_level_off);
        c->dc_level_off = nullptr;
        c->dc_level_cap = level_off.size() * 2;
        if (!dev_alloc(&c->dc_level_off, c->dc_level_cap)) { ps_set_error("ps2d: distance table allocation failed"); return PS_ERR_CUDA; }
    }
    CU2(upload(c, c->dc_i1, i1.data(), i1.size()));
    CU2(upload(c, c->dc_i2, i2.data(), i2.size()));
    CU2(upload(c, c->dc_rest, rest.data(), rest.size()));
    CU2(upload(c, c->dc_level_off, level_off.data(), level_off.size()));
    c->dc_total = (u32)i1.size();
    {   // the list as k2d_tick_fused walks it: fluid / gas constraints one by one, distance constraints run by run
        std::vector<FusedOp> ops;
        size_t run = 0;
        for (size_t q = 0; q < m;) {
            const StdOp &o = c->standard[q];
            FusedOp f{};
            f.kind = (u32)o.kind; f.index = (u32)q;
            if (o.kind == STD_DISTANCE) {
                const DistanceRun &R = c->runs[run++];
                f.level_first = R.dev_level_first; f.levels = R.levels; f.width = R.width;
                q = R.end;
            } else {
                f.open = (u32)o.open; f.p0 = o.p0;
                for (const FluidEmitterRec &fe : c->fluid_emitters) if (fe.standard_index == q) f.keep = 1;
                q++;
            }
            ops.push_back(f);
        }
        if (ops.size() > c->fused_ops_cap) {
            if (c->fused_ops) cudaFree(c->fused_ops);
            c->fused_ops = nullptr;
            c->fused_ops_cap = ops.size() * 2;
            if (!dev_alloc(&c->fused_ops, c->fused_ops_cap)) { ps_set_error("ps2d: op table allocation failed"); return PS_ERR_CUDA; }
        }
        CU2(upload(c, c->fused_ops, ops.data(), ops.size()));
        c->fused_nops = (u32)ops.size();
    }
    c->standard_dirty = false;
    c->standard_version++;
    return PS_OK;
}

// FluidEmitter::tick (cpu/src/fluidemitter.cpp:13-79) after a solver tick: slow fluid particles low in the scene count down
// and freeze into immovable solids that leave the fluid; new fluid particles are emitted for the first 5 seconds.  Host
// logic over the emitter's fluid (a few hundred particles), like the reference's; it reads p, v and the constraint's
// lambdas of the last iteration.  [sic] The reference reads that hash — keyed by particle index — with the position i
// in the member list: `m_fs->lambdas[i]`; a key that is not a member reads 0.  Kept, or the scene would evolve differently.
static int fluid_emitters_tick(Ps2dCtx *c, double dt) {
    if (c->fluid_emitters.empty()) return PS_OK;
    std::vector<double> p, v, lam;
    for (FluidEmitterRec &e : c->fluid_emitters) {
        const u32 n = c->n;
        p.resize(2 * (size_t)n); v.resize(2 * (size_t)n); lam.resize(n);
        CU2(cudaMemcpyAsync(p.data(), c->p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
        CU2(cudaMemcpyAsync(v.data(), c->v, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
        CU2(cudaMemcpyAsync(lam.data(), c->lambda_keep, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
        CU2(cudaStreamSynchronize(c->stream));
        const int id = (int)e.standard_index;
        std::vector<u32> ps;  // the constraint's member list: ascending particle index (append-only, removeAt keeps the order)
        for (u32 k = 0; k < n; k++) if (c->h_group[k] == id) ps.push_back(k);
        for (int i = (int)ps.size() - 1; i >= 0; i--) {
            const u32 k = ps[i];
            const double vx = v[2 * k], vy = v[2 * k + 1];
            if (std::sqrt(vx * vx + vy * vy) < .06 && p[2 * k + 1] <= 5) {
                const double l = ((u32)i < n && c->h_group[i] == id) ? lam[i] : 0.;
                if (l <= 0) {
                    c->h_t[k] -= 1;
                    if (c->h_t[k] <= 0) {
                        c->h_t[k] = 0;
                        const double zero2[2] = {0., 0.}, zero = 0.;
                        const int solid = PS2D_PHASE_SOLID, none = -1;
                        CU2(upload(c, c->imass + k, &zero, 1));
                        CU2(upload(c, c->tmass + k, &zero, 1));
                        CU2(upload(c, c->phase + k, &solid, 1));
                        CU2(upload(c, c->group + k, &none, 1));
                        CU2(upload(c, (double *)(c->ep + k), &p[2 * k], 2));
                        CU2(upload(c, (double *)(c->v + k), zero2, 2));
                        CU2(upload(c, (double *)(c->f + k), zero2, 2));
                        c->h_imass[k] = 0.; c->h_phase[k] = solid; c->h_group[k] = -1;
                        c->any_solid = 1;
                    }
                } else {
                    c->h_t[k] += dt;
                    if (c->h_t[k] > 3) c->h_t[k] = 3;
                }
            }
        }
        e.timer += dt;
        e.total_timer += dt;
        while (e.total_timer < 5 && e.timer >= 1. / e.rate) {
            e.timer -= 1. / e.rate;
            const double pos[2] = {e.x, e.y}, im = 1. / 1.;
            const double vel[2] = {(double)(float)((double)c->rng.next() / (double)2147483647), 1.};  // glm::dvec2(frand(), 1)
            const int32_t ph = PS2D_PHASE_FLUID;
            uint64_t at = 0;
            int r = append_particles(c, pos, vel, &im, &ph, nullptr, nullptr, nullptr, &id, 1, &at);  // a member of the emitter's fluid from the start
            if (r != PS_OK) return r;
        }
    }
    return PS_OK;
}

// createFluidEmitter(posn, particlesPerSec, fs) (simulation.cpp:459-461)
extern "C" int ps2d_create_fluid_emitter(Ps2dCtx *c, const double *posn2, double rate, uint32_t standard_index, double timer, double total_timer) {
    if (!c || !posn2) { ps_set_error("ps2d_create_fluid_emitter: null argument"); return PS_ERR_INVALID; }
    if (!(rate > 0.)) { ps_set_error("ps2d_create_fluid_emitter: rate must be positive"); return PS_ERR_INVALID; }
    if (standard_index >= c->standard.size() || c->standard[standard_index].kind != STD_FLUID) {
        ps_set_error("ps2d_create_fluid_emitter: STANDARD constraint %u is not a fluid", standard_index); return PS_ERR_INVALID;
    }
    CU2(cudaSetDevice(c->device));
    if (!c->lambda_keep) {
        if (!dev_alloc(&c->lambda_keep, c->cap)) { ps_set_error("ps2d_create_fluid_emitter: allocation failed"); return PS_ERR_CUDA; }
        CU2(cudaMemset(c->lambda_keep, 0, c->cap * 8));
    }
    c->fluid_emitters.push_back(FluidEmitterRec{posn2[0], posn2[1], rate, timer, total_timer, standard_index});
    c->standard_dirty = true;  // the op table of the fused tick marks the constraints whose lambda an emitter reads
    return PS_OK;
}
extern "C" int ps2d_set_particle_timers(Ps2dCtx *c, const double *t) {
    if (!c || !t) { ps_set_error("ps2d_set_particle_timers: null argument"); return PS_ERR_INVALID; }
    c->h_t.assign(t, t + c->n);
    return PS_OK;
}
extern "C" int ps2d_get_particle_timers(Ps2dCtx *c, double *t) {
    if (!c || !t) { ps_set_error("ps2d_get_particle_timers: null argument"); return PS_ERR_INVALID; }
    std::copy(c->h_t.begin(), c->h_t.end(), t);
    return PS_OK;
}

// The launch sequence of one tick, from the prediction to the final scalars' copy (what ps2d_tick replays as a graph).
static u32 issue_tick(Ps2dCtx *c, double dt) {
    cudaStream_t s = c->stream;
    const u32 n = c->n, blocks = (n + kBlock - 1) / kBlock;
    const Ps2dParams &P = c->params;
    const double x0 = P.x_bounds[0], x1 = P.x_bounds[1], y0 = P.y_bounds[0], y1 = P.y_bounds[1];
    u32 launches = 0;
    k2d_predict<<<blocks, kBlock, 0, s>>>(c->v, c->ep, c->f, c->tmass, c->p, c->imass, c->phase, n, dt, P.gravity[0], P.gravity[1]);
    k2d_find_contacts<<<(n + kBlock / 32 - 1) / (kBlock / 32), kBlock, 0, s>>>(c->ep, c->imass, c->phase, c->bod, c->static_counts, n, x0, x1, y0, y1, c->nb, c->cnt, c->flags, c->counts,
                                                c->draws, c->scalars + 3, c->any_solid);
    k2d_scan_counts<<<1, 1024, 0, s>>>(c->draws, c->rank, n, c->scalars);
    k2d_contact_levels<<<1, kSerialBlock, 0, s>>>(c->nb, c->cnt, c->flags, n, c->nbq, c->lvl, c->scalars + 1);
    launches += 4;
    Particles2D V{c->ep, c->p, c->tmass, c->sfric, c->kfric, c->phase, c->bod, c->counts, c->sdf_grad, c->sdf_dist, c->b_angle};
    // stabilization passes (the reference's compile-time option USE_STABILIZATION, simulation.cpp:249-271): before the solver iterations
    for (u32 st = 0; st < P.stabilization_iterations; st++) {
        k2d_contact_project<<<1, kSerialBlock, 0, s>>>(V, c->nb, c->cnt, c->flags, c->lvl, c->rank, c->scalars + 1, c->cur, n, c->raw, c->window_base_dev, st, x0, x1, y0, y1, true);
        launches++;
    }
    for (u32 it = 0; it < P.solver_iterations; it++) {
        // CONTACT group (the kernel returns at once when the list is empty)
        k2d_contact_project<<<1, kSerialBlock, 0, s>>>(V, c->nb, c->cnt, c->flags, c->lvl, c->rank, c->scalars + 1, c->cur, n, c->raw, c->window_base_dev,
                                                       P.stabilization_iterations + it, x0, x1, y0, y1, false);
        launches++;
        size_t run = 0;  // STANDARD group, in list order
        for (size_t k = 0; k < c->standard.size();) {
            const StdOp &op = c->standard[k];
            if (op.kind == STD_DISTANCE) {
                const DistanceRun &R = c->runs[run++];
                const u32 count = (u32)(R.end - R.begin);
                const size_t stage = distance_stage_bytes(n, count);
                if (stage <= kDistanceStageMax) {  // the reference's scenes: a few hundred particles
                    const u32 block = std::min<u32>(kSerialBlock, std::max<u32>(32u, (R.width + 31u) & ~31u));
                    if (block == 32)
                        k2d_distance_run_staged<32><<<1, 32, stage, s>>>(c->ep, c->imass, c->counts, c->dc_i1, c->dc_i2, c->dc_rest, c->dc_level_off + R.dev_level_first,
                                                                         R.levels, n, R.dev_first, count);
                    else
                        k2d_distance_run_staged<kSerialBlock><<<1, block, stage, s>>>(c->ep, c->imass, c->counts, c->dc_i1, c->dc_i2, c->dc_rest,
                                                                                      c->dc_level_off + R.dev_level_first, R.levels, n, R.dev_first, count);
                } else {
                    k2d_distance_run<<<1, kSerialBlock, 0, s>>>(c->ep, c->imass, c->counts, c->dc_i1, c->dc_i2, c->dc_rest, c->dc_level_off + R.dev_level_first, R.levels);
                }
                launches++;
                k = R.end;
                continue;
            }
            const FluidConsts K = op.kind == STD_GAS ? FluidConsts{.2, .25, .5, 1, op.open} : FluidConsts{.1, .2, 0., 0, 0};
            const u32 wblocks = (n + kBlock / 32 - 1) / (kBlock / 32);  // one warp per particle
            k2d_fluid_lambda<<<wblocks, kBlock, 0, s>>>(c->ep, c->imass, c->phase, c->group, n, (int)k, op.p0, K, c->lambda, c->nbcount, c->nbr_keep, c->v, c->f);
            if (it + 1 == P.solver_iterations)
                for (const FluidEmitterRec &fe : c->fluid_emitters)
                    if (fe.standard_index == k) CU2(cudaMemcpyAsync(c->lambda_keep, c->lambda, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
            k2d_fluid_delta<<<wblocks, kBlock, 0, s>>>(c->ep, c->imass, c->group, n, (int)k, op.p0, K, c->lambda, c->nbcount, c->nbr_keep, c->counts, c->delta, c->v, c->f);
            k2d_fluid_apply<<<blocks, kBlock, 0, s>>>(c->ep, c->delta, c->group, n, (int)k);
            launches += 3;
            k++;
        }
        if (c->nbodies) {  // SHAPE group
            k2d_shape<<<(c->nbodies + kBlock - 1) / kBlock, kBlock, 0, s>>>(c->ep, c->imass, c->rs, c->b_first, c->b_count, c->b_imass, c->b_stiff, c->b_center,
                                                                             c->b_angle, c->nbodies);
            launches++;
        }
    }
    k2d_finish<<<blocks, kBlock, 0, s>>>(c->p, c->v, c->ep, n, dt);
    launches++;
    cudaMemcpyAsync(c->scalars_host, c->scalars, 16, cudaMemcpyDeviceToHost, s);
    return launches;
}

extern "C" int ps2d_tick(Ps2dCtx *c, double dt) {
    struct Range { Range() { nvtxRangePushA("ps2d_tick"); } ~Range() { nvtxRangePop(); } } nvtx;  // ncu / nsys timelines
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!c->n) return PS_OK;
    CU2(cudaSetDevice(c->device));
    if (c->standard_dirty) {
        int r = rebuild_standard(c);
        if (r != PS_OK) return r;
    }
    cudaStream_t s = c->stream;
    const u32 n = c->n;
    const Ps2dParams &P = c->params;
    // the number of jittered wall constraints decides how many draws of the rand() stream this tick consumes — one per
    // constraint per solver iteration, in list order; at most 2 n per iteration.  The kernels take them from the look-ahead
    // window, which is refilled (from a copy of the generator) when it no longer covers a worst-case tick.
    const u32 passes = P.stabilization_iterations + P.solver_iterations;  // every pass over the wall constraints draws (the stabile copies too)
    const uint64_t worst = (uint64_t)2 * n * passes;
    if (c->any_jitter && (c->rng.calls < c->win_pos || c->rng.calls + worst > c->win_pos + c->win_len)) {
        const size_t want = std::max<size_t>(worst * 16, 1u << 16);
        GlibcRand ahead = c->rng;
        c->h_raw.resize(want);
        for (size_t k = 0; k < want; k++) c->h_raw[k] = ahead.next();
        if (want > c->raw_cap) {
            CU2(cudaStreamSynchronize(s));
            if (c->raw) CU2(cudaFree(c->raw));
            c->raw = nullptr;
            CU2(cudaMalloc(&c->raw, want * sizeof(int)));
            c->raw_cap = want;
        }
        CU2(cudaMemcpyAsync(c->raw, c->h_raw.data(), want * sizeof(int), cudaMemcpyHostToDevice, s));
        CU2(cudaStreamSynchronize(s));  // h_raw is pageable
        c->win_pos = c->rng.calls;
        c->win_len = want;
    }
    // scenes of up to kFusedTickMaxN particles: the whole tick as one kernel (k2d_tick_fused); PS2D_FUSED_MAX_N moves the threshold (0: never)
    static const u32 fused_max_n = [] { const char *e = getenv("PS2D_FUSED_MAX_N"); return e ? std::min<u32>((u32)strtoul(e, nullptr, 10), kFusedTickCapN) : kFusedTickMaxN; }();
    const size_t fused_stage = (size_t)c->dc_total * sizeof(DistSlot) + (size_t)n * (sizeof(double2) + sizeof(double) + sizeof(u32) + 2 * sizeof(int));
    c->last_tick_fused = n <= fused_max_n && c->scalars_out != nullptr && fused_stage <= kFusedStageMax;
    // draws of this tick start at window[window_base]; the kernels of the launch sequence read it from a device word
    c->scalars_host[4] = (u32)(c->rng.calls - c->win_pos);
    if (!c->last_tick_fused) CU2(cudaMemcpyAsync(c->window_base_dev, c->scalars_host + 4, 4, cudaMemcpyHostToDevice, s));  // the fused kernel takes it as an argument
    static const bool no_graph = getenv("PS_NO_GRAPH") != nullptr;
    {   // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: once for each device this process drives
        static bool opted[64] = {};
        if (c->device >= 0 && c->device < 64 && !opted[c->device]) {
            CU2(cudaFuncSetAttribute(k2d_distance_run_staged<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDistanceStageMax));
            CU2(cudaFuncSetAttribute(k2d_distance_run_staged<kSerialBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDistanceStageMax));
            opted[c->device] = true;
        }
    }
    if (c->last_tick_fused) {
        static bool opted_fused[64] = {};
        const size_t stage = fused_stage;
        if (c->device >= 0 && c->device < 64 && !opted_fused[c->device]) {
            CU2(cudaFuncSetAttribute(k2d_tick_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedStageMax));
            opted_fused[c->device] = true;
        }
        FusedTick A{};
        A.v = c->v; A.ep = c->ep; A.f = c->f; A.p = c->p; A.delta = c->delta;
        A.tmass = c->tmass; A.lambda = c->lambda; A.lambda_keep = c->lambda_keep;
        A.imass = c->imass; A.sfric = c->sfric; A.kfric = c->kfric; A.sdf_dist = c->sdf_dist;
        A.sdf_grad = c->sdf_grad; A.rs = c->rs;
        A.phase = c->phase; A.bod = c->bod; A.group = c->group; A.raw = c->raw;
        A.static_counts = c->static_counts;
        A.nb = c->nb; A.cnt = c->cnt; A.flags = c->flags; A.counts = c->counts; A.draws = c->draws; A.rank = c->rank; A.lvl = c->lvl; A.nbcount = c->nbcount; A.nbr_keep = c->nbr_keep;
        A.scalars = c->scalars; A.scalars_out = c->scalars_out;
        A.nbq = c->nbq;
        A.b_first = c->b_first; A.b_count = c->b_count; A.b_imass = c->b_imass; A.b_stiff = c->b_stiff; A.b_center = c->b_center; A.b_angle = c->b_angle;
        A.dc_i1 = c->dc_i1; A.dc_i2 = c->dc_i2; A.dc_level_off = c->dc_level_off; A.dc_rest = c->dc_rest;
        A.ops = c->fused_ops;
        A.n = n; A.nops = c->fused_nops; A.nbodies = c->nbodies; A.ndist = c->dc_total; A.window_base = c->scalars_host[4];
        A.stabilization_iterations = P.stabilization_iterations; A.solver_iterations = P.solver_iterations;
        A.any_solid = c->any_solid;
        A.prof = c->fused_prof;
        c->fused_ticks++;
        A.dt = dt; A.gx = P.gravity[0]; A.gy = P.gravity[1];
        A.x0 = P.x_bounds[0]; A.x1 = P.x_bounds[1]; A.y0 = P.y_bounds[0]; A.y1 = P.y_bounds[1];
        static const u32 cluster_want = [] { const char *e = getenv("PS2D_FUSED_CLUSTER"); return e ? std::max<u32>(1u, std::min<u32>((u32)strtoul(e, nullptr, 10), kFusedClusterMax)) : kFusedCluster; }();
        static u32 cluster_ok[64] = {};  // per device: the largest cluster of this kernel the device schedules (0 = not asked yet)
        u32 &cluster_ctas = cluster_ok[(c->device >= 0 && c->device < 64) ? c->device : 0];
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(kFusedBlock); cfg.dynamicSmemBytes = kFusedStageMax; cfg.stream = s;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (!cluster_ctas) {
            if (cluster_want > 8) CU2(cudaFuncSetAttribute(k2d_tick_fused, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            for (cluster_ctas = cluster_want; cluster_ctas > 1; cluster_ctas /= 2) {  // sizes above 8 are not portable: ask, halve if refused
                int fits = 0;
                cfg.gridDim = dim3(cluster_ctas); attr.val.clusterDim.x = cluster_ctas;
                if (cudaOccupancyMaxActiveClusters(&fits, k2d_tick_fused, &cfg) == cudaSuccess && fits > 0) break;
                cudaGetLastError();
            }
        }
        cfg.gridDim = dim3(cluster_ctas); attr.val.clusterDim.x = cluster_ctas; cfg.dynamicSmemBytes = stage;
        CU2(cudaLaunchKernelEx(&cfg, k2d_tick_fused, A));
        c->launches = 1;
    } else if (no_graph) {
        c->launches = issue_tick(c, dt);
    } else {
        Ps2dCtx::TickKey key;
        memset(&key, 0, sizeof(key));
        key.n = n; key.dt = dt; key.standard_version = c->standard_version; key.params = P;
        key.bodies = c->b_first; key.raw = c->raw; key.lambda_keep = c->lambda_keep;
        key.misc = (uint64_t)c->fluid_emitters.size() | ((uint64_t)c->nbodies << 20) | ((uint64_t)(c->any_solid ? 1 : 0) << 60);
        // a capture costs a few eager ticks: scenes whose emitters add particles every other tick keep changing the key and are
        // issued eagerly; the graph is (re)built once the key has stood still for kStableTicks ticks
        constexpr u32 kStableTicks = 4;
        const bool same = memcmp(&key, &c->tick_key, sizeof(key)) == 0;
        if (!same) {
            if (c->tick_graph) { cudaGraphExecDestroy(c->tick_graph); c->tick_graph = nullptr; }
            memcpy(&c->tick_key, &key, sizeof(key));
            c->tick_key_age = 0;
        }
        c->tick_key_age++;
        if (!c->tick_graph && c->tick_key_age < kStableTicks) {
            c->launches = issue_tick(c, dt);
        } else if (!c->tick_graph) {
            cudaGraph_t graph = nullptr;
            CU2(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            c->launches = issue_tick(c, dt);
            cudaError_t ce = cudaStreamEndCapture(s, &graph);
            if (ce != cudaSuccess) { ps_set_error("ps2d_tick: stream capture failed: %s", cudaGetErrorString(ce)); return PS_ERR_CUDA; }
            ce = cudaGraphInstantiate(&c->tick_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { c->tick_graph = nullptr; ps_set_error("ps2d_tick: graph instantiation failed: %s", cudaGetErrorString(ce)); return PS_ERR_CUDA; }
            CU2(cudaGraphLaunch(c->tick_graph, s));
        } else {
            CU2(cudaGraphLaunch(c->tick_graph, s));
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ps_set_error("ps2d_tick: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
    // the one host round trip of a tick, at its end: what the CONTACT list looked like and how many draws it consumed
    CU2(cudaStreamSynchronize(s));
    const u32 num = c->scalars_host[0];
    c->last_num_boundary = num;
    c->last_levels = c->scalars_host[1];
    c->last_contacts = c->scalars_host[2];
    if (c->scalars_host[3]) {
        ps_set_error("ps2d_tick: a particle has %u contacts, more than PS2D_MAX_CONTACTS = %d", c->scalars_host[3], kMaxC);
        cudaMemsetAsync(c->scalars + 3, 0, 4, s);
        return PS_ERR_CAPACITY;
    }
    for (size_t k = 0, nd = (size_t)num * passes; k < nd; k++) c->rng.next();  // the stream moves on by what was drawn
    // OpenSmokeEmitter::tick (opensmokeemitter.cpp:17-29): particle injection into the gas constraint
    for (Emitter &em : c->emitters) {
        em.timer += dt;
        while (em.timer >= 1. / em.rate) {
            em.timer -= 1. / em.rate;
            if (em.standard_index == UINT32_MAX) continue;
            const double pos[2] = {em.x, em.y}, im = 1. / 1.;
            const int32_t ph = PS2D_PHASE_GAS;
            uint64_t at = 0;
            const int id = (int)em.standard_index;
            int r = append_particles(c, pos, nullptr, &im, &ph, nullptr, nullptr, nullptr, &id, 1, &at);  // a member of the gas from the start
            if (r != PS_OK) return r;
        }
    }
    return fluid_emitters_tick(c, dt);
}

extern "C" int ps2d_download(Ps2dCtx *c, int which, void *host) {
    if (!c) { ps_set_error("ps2d_download: null context"); return PS_ERR_INVALID; }
    if (c->n == 0) return PS_OK;  // nothing to copy (host may be NULL)
    if (!host) { ps_set_error("ps2d_download: null argument"); return PS_ERR_INVALID; }
    CU2(cudaSetDevice(c->device));
    const void *src = nullptr;
    size_t bytes = (size_t)c->n * 16;
    switch (which) {
        case PS2D_ARR_P: src = c->p; break;
        case PS2D_ARR_V: src = c->v; break;
        case PS2D_ARR_EP: src = c->ep; break;
        case PS2D_ARR_F: src = c->f; break;
        case PS2D_ARR_LAMBDA: src = c->lambda; bytes = (size_t)c->n * 8; break;
        case PS2D_ARR_TMASS: src = c->tmass; bytes = (size_t)c->n * 8; break;
        case PS2D_ARR_COUNTS: src = c->counts; bytes = (size_t)c->n * 4; break;
        case PS2D_ARR_IMASS: src = c->imass; bytes = (size_t)c->n * 8; break;
        case PS2D_ARR_SFRICTION: src = c->sfric; bytes = (size_t)c->n * 8; break;
        case PS2D_ARR_KFRICTION: src = c->kfric; bytes = (size_t)c->n * 8; break;
        case PS2D_ARR_SDF_DIST: src = c->sdf_dist; bytes = (size_t)c->n * 8; break;
        case PS2D_ARR_RS: src = c->rs; break;
        case PS2D_ARR_SDF_GRAD: src = c->sdf_grad; break;
        case PS2D_ARR_PHASE: src = c->phase; bytes = (size_t)c->n * 4; break;
        case PS2D_ARR_BOD: src = c->bod; bytes = (size_t)c->n * 4; break;
        case PS2D_ARR_GROUP: src = c->group; bytes = (size_t)c->n * 4; break;
        default: ps_set_error("ps2d_download: unknown array %d", which); return PS_ERR_INVALID;
    }
    CU2(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU2(cudaStreamSynchronize(c->stream));
    return PS_OK;
}

extern "C" int ps2d_kinetic_energy(Ps2dCtx *c, double *out) {
    if (!c || !out) return PS_ERR_INVALID;
    *out = 0.;
    if (c->n == 0) return PS_OK;
    std::vector<double> v((size_t)c->n * 2);
    int r = ps2d_download(c, PS2D_ARR_V, v.data());
    if (r != PS_OK) return r;
    double e = 0;  // same order as Simulation::getKineticEnergy
    for (u32 i = 0; i < c->n; i++)
        if (c->h_imass[i] != 0.) e += .5 * (v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]) / c->h_imass[i];
    *out = e;
    return PS_OK;
}

// ------------------------------------------------------------------ checkpoints ------------------------------------------------------------------
// The restartable state of a Ps2dCtx: parameters, per-particle arrays, rigid bodies, the STANDARD list in
// order, emitters and the rand() stream (its 31 state words).  A run continued from a checkpoint is bit-identical to the
// uninterrupted one (tests/test_gpu_checkpoint.py).  Little-endian, fixed-width fields.
namespace {
const char kMagic2d[8] = {'P', 'S', 'B', '2', 'D', 'C', 'K', '1'};
struct Header2d {
    char magic[8];
    uint32_t version, params_bytes;
    uint64_t n, cap, num_bodies, num_standard, num_emitters, num_fluid_emitters, rand_calls;
};
struct FluidEmitRecord { double x, y, rate, timer, total_timer; uint32_t standard_index, pad; };
struct StdRecord { uint32_t kind, open, i1, i2; double p0, d; };
struct EmitRecord { double x, y, rate, timer; uint32_t standard_index, pad; };
template <class T> bool put2(FILE *f, const T *p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <class T> bool get2(FILE *f, T *p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
template <class T> cudaError_t fetch(std::vector<T> &h, const void *d, size_t count) {
    h.resize(count);
    return count ? cudaMemcpy(h.data(), d, count * sizeof(T), cudaMemcpyDeviceToHost) : cudaSuccess;
}
}  // namespace

extern "C" int ps2d_save(Ps2dCtx *c, const char *path) {
    if (!c || !path) { ps_set_error("ps2d_save: null argument"); return PS_ERR_INVALID; }
    CU2(cudaSetDevice(c->device));
    CU2(cudaStreamSynchronize(c->stream));
    const size_t n = c->n, nb = c->nbodies;
    std::vector<double> p, v, f, rs, sg, im, sf, kf, sd, bim, bst, ban, bce;
    std::vector<int> ph, bod, grp;
    std::vector<u32> bf, bc;
    CU2(fetch(p, c->p, 2 * n)); CU2(fetch(v, c->v, 2 * n)); CU2(fetch(f, c->f, 2 * n)); CU2(fetch(rs, c->rs, 2 * n)); CU2(fetch(sg, c->sdf_grad, 2 * n));
    CU2(fetch(im, c->imass, n)); CU2(fetch(sf, c->sfric, n)); CU2(fetch(kf, c->kfric, n)); CU2(fetch(sd, c->sdf_dist, n));
    CU2(fetch(ph, c->phase, n)); CU2(fetch(bod, c->bod, n)); CU2(fetch(grp, c->group, n));
    CU2(fetch(bf, c->b_first, nb)); CU2(fetch(bc, c->b_count, nb)); CU2(fetch(bim, c->b_imass, nb)); CU2(fetch(bst, c->b_stiff, nb));
    CU2(fetch(ban, c->b_angle, nb)); CU2(fetch(bce, c->b_center, 2 * nb));
    Header2d h{};
    memcpy(h.magic, kMagic2d, 8);
    h.version = 3; h.params_bytes = (uint32_t)sizeof(Ps2dParams);  // 3: Ps2dParams.stabilization_iterations lives in what was padding
    h.n = n; h.cap = c->cap; h.num_bodies = nb; h.num_standard = c->standard.size(); h.num_emitters = c->emitters.size();
    h.num_fluid_emitters = c->fluid_emitters.size(); h.rand_calls = c->rng.calls;
    std::vector<FluidEmitRecord> fem;
    for (const FluidEmitterRec &e : c->fluid_emitters) fem.push_back(FluidEmitRecord{e.x, e.y, e.rate, e.timer, e.total_timer, e.standard_index, 0});
    std::vector<double> lk;
    if (c->lambda_keep) CU2(fetch(lk, c->lambda_keep, n));
    std::vector<StdRecord> st;
    for (const StdOp &o : c->standard) st.push_back(StdRecord{(uint32_t)o.kind, (uint32_t)o.open, o.i1, o.i2, o.p0, o.d});
    std::vector<EmitRecord> em;
    for (const Emitter &e : c->emitters) em.push_back(EmitRecord{e.x, e.y, e.rate, e.timer, e.standard_index, 0});
    FILE *fp = fopen(path, "wb");
    if (!fp) { ps_set_error("ps2d_save: cannot open %s for writing", path); return PS_ERR_INVALID; }
    bool ok = put2(fp, &h, 1) && put2(fp, &c->params, 1) && put2(fp, c->rng.r.data(), 31) && put2(fp, p.data(), p.size()) && put2(fp, v.data(), v.size()) &&
              put2(fp, f.data(), f.size()) && put2(fp, rs.data(), rs.size()) && put2(fp, sg.data(), sg.size()) && put2(fp, im.data(), n) && put2(fp, sf.data(), n) &&
              put2(fp, kf.data(), n) && put2(fp, sd.data(), n) && put2(fp, ph.data(), n) && put2(fp, bod.data(), n) && put2(fp, grp.data(), n) &&
              put2(fp, c->h_static_counts.data(), n) && put2(fp, bf.data(), nb) && put2(fp, bc.data(), nb) && put2(fp, bim.data(), nb) && put2(fp, bst.data(), nb) &&
              put2(fp, ban.data(), nb) && put2(fp, bce.data(), 2 * nb) && put2(fp, st.data(), st.size()) && put2(fp, em.data(), em.size()) &&
              put2(fp, c->h_t.data(), n) && put2(fp, fem.data(), fem.size()) && put2(fp, lk.data(), lk.size());
    ok = (fclose(fp) == 0) && ok;
    if (!ok) { ps_set_error("ps2d_save: short write to %s", path); return PS_ERR_INVALID; }
    return PS_OK;
}

extern "C" int ps2d_load(const char *path, int device, Ps2dCtx **out) {
    if (!path || !out) { ps_set_error("ps2d_load: null argument"); return PS_ERR_INVALID; }
    *out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) { ps_set_error("ps2d_load: cannot open %s", path); return PS_ERR_INVALID; }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{fp};
    Header2d h{};
    if (!get2(fp, &h, 1) || memcmp(h.magic, kMagic2d, 8) != 0) { ps_set_error("ps2d_load: %s is not a 2-D libpsolver checkpoint", path); return PS_ERR_INVALID; }
    if ((h.version != 2 && h.version != 3) || h.params_bytes != sizeof(Ps2dParams)) { ps_set_error("ps2d_load: checkpoint version %u not understood", h.version); return PS_ERR_INVALID; }
    if (h.n > h.cap || h.cap > (1u << 28) || h.num_bodies > h.n || h.num_standard > (1u << 28) || h.num_emitters > 4096 || h.num_fluid_emitters > 4096) { ps_set_error("ps2d_load: implausible sizes in %s", path); return PS_ERR_INVALID; }
    Ps2dParams P;
    uint32_t rng[31];
    const size_t n = h.n, nb = h.num_bodies;
    std::vector<double> p(2 * n), v(2 * n), f(2 * n), rs(2 * n), sg(2 * n), im(n), sf(n), kf(n), sd(n), bim(nb), bst(nb), ban(nb), bce(2 * nb);
    std::vector<int> ph(n), bod(n), grp(n);
    std::vector<u32> sc(n), bf(nb), bc(nb);
    std::vector<StdRecord> st(h.num_standard);
    std::vector<EmitRecord> em(h.num_emitters);
    std::vector<FluidEmitRecord> fem(h.num_fluid_emitters);
    std::vector<double> ht(n), lk(h.num_fluid_emitters ? n : 0);
    bool ok = get2(fp, &P, 1) && (h.version >= 3 || (P.stabilization_iterations = 0, true)) /* version 2: padding */ && get2(fp, rng, 31) && get2(fp, p.data(), p.size()) && get2(fp, v.data(), v.size()) && get2(fp, f.data(), f.size()) &&
              get2(fp, rs.data(), rs.size()) && get2(fp, sg.data(), sg.size()) && get2(fp, im.data(), n) && get2(fp, sf.data(), n) && get2(fp, kf.data(), n) &&
              get2(fp, sd.data(), n) && get2(fp, ph.data(), n) && get2(fp, bod.data(), n) && get2(fp, grp.data(), n) && get2(fp, sc.data(), n) &&
              get2(fp, bf.data(), nb) && get2(fp, bc.data(), nb) && get2(fp, bim.data(), nb) && get2(fp, bst.data(), nb) && get2(fp, ban.data(), nb) &&
              get2(fp, bce.data(), 2 * nb) && get2(fp, st.data(), st.size()) && get2(fp, em.data(), em.size()) && get2(fp, ht.data(), n) &&
              get2(fp, fem.data(), fem.size()) && get2(fp, lk.data(), lk.size());
    if (!ok) { ps_set_error("ps2d_load: truncated file %s", path); return PS_ERR_INVALID; }
    for (size_t b = 0; b < nb; b++) if ((uint64_t)bf[b] + bc[b] > n) { ps_set_error("ps2d_load: body out of range"); return PS_ERR_INVALID; }
    // index fields that address device arrays.  `bod` is the reference's free-form body tag (fluids carry 100 * frand(), ropes -2 / -3,
    // simulation.cpp:411,865,1045); only a SOLID particle's non-negative tag indexes the body table (sdf_data).  Constraint group in
    // [-1, STANDARD records); emitters' STANDARD record (UINT32_MAX = none).
    for (size_t k = 0; k < n; k++)
        if ((ph[k] == PS2D_PHASE_SOLID && bod[k] >= (int)nb) || grp[k] < -1 || (grp[k] >= 0 && (uint64_t)grp[k] >= h.num_standard)) { ps_set_error("ps2d_load: particle %zu refers to a body / constraint group that does not exist", k); return PS_ERR_INVALID; }
    for (const EmitRecord &e : em) if (e.standard_index != 0xffffffffu && e.standard_index >= h.num_standard) { ps_set_error("ps2d_load: emitter refers to a missing STANDARD record"); return PS_ERR_INVALID; }
    for (const FluidEmitRecord &e : fem) if (e.standard_index != 0xffffffffu && e.standard_index >= h.num_standard) { ps_set_error("ps2d_load: fluid emitter refers to a missing STANDARD record"); return PS_ERR_INVALID; }
    for (const StdRecord &o : st) if (o.kind > STD_DISTANCE || (o.kind == STD_DISTANCE && (o.i1 >= n || o.i2 >= n))) { ps_set_error("ps2d_load: bad STANDARD record"); return PS_ERR_INVALID; }
    Ps2dCtx *c = nullptr;
    int r = ps2d_create(device, &P, h.cap, &c);
    if (r != PS_OK) return r;
    auto fail = [&](int code) { ps2d_destroy(c); return code; };
    if (n && (r = ps2d_add_particles(c, p.data(), v.data(), im.data(), ph.data(), bod.data(), sf.data(), kf.data(), n, nullptr)) != PS_OK) return fail(r);
    bool up = upload(c, (double *)c->f, f.data(), 2 * n) == cudaSuccess && upload(c, (double *)c->rs, rs.data(), 2 * n) == cudaSuccess &&
              upload(c, (double *)c->sdf_grad, sg.data(), 2 * n) == cudaSuccess && upload(c, c->sdf_dist, sd.data(), n) == cudaSuccess &&
              upload(c, c->group, grp.data(), n) == cudaSuccess && upload(c, c->static_counts, sc.data(), n) == cudaSuccess;
    c->h_static_counts = sc;
    c->h_group = grp;
    c->h_t = ht;
    for (size_t b = 0; b < nb && up; b++) {
        if (grow_bodies(c) != PS_OK) return fail(PS_ERR_CUDA);
        up = upload(c, c->b_first + b, &bf[b], 1) == cudaSuccess && upload(c, c->b_count + b, &bc[b], 1) == cudaSuccess && upload(c, c->b_imass + b, &bim[b], 1) == cudaSuccess &&
             upload(c, c->b_stiff + b, &bst[b], 1) == cudaSuccess && upload(c, c->b_angle + b, &ban[b], 1) == cudaSuccess && upload(c, (double *)(c->b_center + b), &bce[2 * b], 2) == cudaSuccess;
        c->nbodies++;
    }
    if (!up) { ps_set_error("ps2d_load: upload failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(PS_ERR_CUDA); }
    for (const StdRecord &o : st) {
        StdOp op;
        op.kind = (StdKind)o.kind; op.open = (int)o.open; op.i1 = o.i1; op.i2 = o.i2; op.p0 = o.p0; op.d = o.d;
        c->standard.push_back(op);
    }
    c->standard_dirty = true;
    for (const EmitRecord &e : em) c->emitters.push_back(Emitter{e.x, e.y, e.rate, e.timer, e.standard_index});
    for (const FluidEmitRecord &e : fem) {
        const double posn[2] = {e.x, e.y};
        if ((r = ps2d_create_fluid_emitter(c, posn, e.rate, e.standard_index, e.timer, e.total_timer)) != PS_OK) return fail(r);
    }
    if (!fem.empty() && upload(c, c->lambda_keep, lk.data(), n) != cudaSuccess) { ps_set_error("ps2d_load: upload failed"); return fail(PS_ERR_CUDA); }
    c->rng.r.assign(rng, rng + 31);
    c->rng.calls = h.rand_calls;
    *out = c;
    return PS_OK;
}
