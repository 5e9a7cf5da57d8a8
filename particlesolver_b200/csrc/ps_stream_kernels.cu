// particlesolver_b200/csrc/ps_stream_kernels.cu — the HBM-streaming kernels of the solver step:
//   K1  predict           (reference integrate_functor + copyToXstar, integration.cu:122-135, integration_kernel.cuh:159-184)
//   K8  world bounds      (collide_world_functor, integration_kernel.cuh:57-157)
//   K9  distance          (solveDistanceConstraints, solver.cu:196-231; gather form, no sort/reduce_by_key)
//   K10 point pins        (point_constraint_functor, solver_kernel.cuh:10-25)
//   K11 velocity update   (subtract_functor, integration_kernel.cuh:465-475)
// All are one pass over SoA float4 arrays with 128-bit, L1-bypassing accesses; the bound is HBM bandwidth.
#include "ps_common.cuh"

namespace {
constexpr int kBlock = 256;

// ---- K1: prev <- pos; v' = v + g dt (local); pos += v' dt.  64 B/particle. ----
__global__ void __launch_bounds__(kBlock) k_predict(float4 *__restrict__ pos, const float4 *__restrict__ vel, float4 *__restrict__ prev,
                                                    u32 n, float dt, float3 g) {
    u32 i = blockIdx.x * (kBlock * 2) + threadIdx.x;
    float4 p[2], v[2];
    bool ok[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        u32 j = i + k * kBlock;
        ok[k] = j < n;
        if (ok[k]) { p[k] = ld_stream4(pos + j); v[k] = ld_stream4(vel + j); }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (!ok[k]) continue;
        u32 j = i + k * kBlock;
        st_stream4(prev + j, p[k]);
        float vx = v[k].x + g.x * dt, vy = v[k].y + g.y * dt, vz = v[k].z + g.z * dt;
        p[k].x += vx * dt; p[k].y += vy * dt; p[k].z += vz * dt;
        st_stream4(pos + j, p[k]);
    }
}

// K1 with buoyant gas (PS_FLAG_GAS; not in the reference's GPU solver): GAS particles feel gravity x PS_GAS_ALPHA, like the CPU
// app's prediction (cpu/src/simulation.cpp:144, ALPHA -.2 simulation.h:21).  One more 4-byte read per particle.
__global__ void __launch_bounds__(kBlock) k_predict_gas(float4 *__restrict__ pos, const float4 *__restrict__ vel, float4 *__restrict__ prev,
                                                        const int *__restrict__ phase, u32 n, float dt, float3 g, float alpha) {
    const u32 j = blockIdx.x * kBlock + threadIdx.x;
    if (j >= n) return;
    float4 p = ld_stream4(pos + j);
    const float4 v = ld_stream4(vel + j);
    const float s = phase[j] == 1 ? alpha : 1.f;
    st_stream4(prev + j, p);
    const float vx = v.x + (g.x * s) * dt, vy = v.y + (g.y * s) * dt, vz = v.z + (g.z * s) * dt;
    p.x += vx * dt; p.y += vy * dt; p.z += vz * dt;
    st_stream4(pos + j, p);
}

// ---- K11: V = (Xstar - pos) / -dt on all four components.  48 B/particle. ----
__global__ void __launch_bounds__(kBlock) k_velocity(const float4 *__restrict__ pos, const float4 *__restrict__ prev,
                                                     float4 *__restrict__ vel, u32 n, float neg_dt) {
    u32 i = blockIdx.x * (kBlock * 2) + threadIdx.x;
    float4 p[2], x[2];
    bool ok[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        u32 j = i + k * kBlock;
        ok[k] = j < n;
        if (ok[k]) { p[k] = ld_stream4(pos + j); x[k] = ld_stream4(prev + j); }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (!ok[k]) continue;
        float4 v;
        v.x = __fdividef(x[k].x - p[k].x, neg_dt);
        v.y = __fdividef(x[k].y - p[k].y, neg_dt);
        v.z = __fdividef(x[k].z - p[k].z, neg_dt);
        v.w = __fdividef(x[k].w - p[k].w, neg_dt);
        st_stream4(vel + i + k * kBlock, v);
    }
}

// ---- K8: clamp to the scene box with shared per-iteration jitter, then wall friction.  52 B/particle. ----
__global__ void __launch_bounds__(kBlock) k_collide_world(float4 *__restrict__ pos, const float4 *__restrict__ prev,
                                                          const int *__restrict__ phase, u32 n, const float *__restrict__ rands,
                                                          WorldDesc w) {
    u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    float4 P = ld_stream4(pos + i);
    float ex = P.x, ey = P.y, ez = P.z;
    const float d = w.radius;
    // the six tests do not depend on the phase: particles away from every wall (the vast majority) leave
    // after one 16-byte load, and their stored bits are unchanged exactly as in the reference
    const bool hit_floor = ey < w.min_y + d;
    if (!(hit_floor || ex > w.max_x - d || ex < w.min_x + d || ey > w.max_y - d || ez > w.max_z - d || ez < w.min_z + d)) return;
    const int ph = __ldg(phase + i);
    const float r0 = __ldg(rands + 0), r1 = __ldg(rands + 1), r2 = __ldg(rands + 2), r3 = __ldg(rands + 3), r4 = __ldg(rands + 4),
                r5 = __ldg(rands + 5);
    float nx = 0.f, ny = 0.f, nz = 0.f;
    float eps = d * 0.f;
    if (ph < PH_SOLID) eps = d * 0.01f;
    if (hit_floor) { ey = w.min_y + d + r5 * eps; ny += 1.f; }
    eps = d * 0.01f;
    if (ex > w.max_x - d) { ex = w.max_x - (d + r0 * eps); nx += -1.f; }
    if (ex < w.min_x + d) { ex = w.min_x + (d + r1 * eps); nx += 1.f; }
    if (ey > w.max_y - d) { ey = w.max_y - (d + r2 * eps); ny += -1.f; }
    if (ez > w.max_z - d) { ez = w.max_z - (d + r3 * eps); nz += -1.f; }
    if (ez < w.min_z + d) { ez = w.min_z + (d + r4 * eps); nz += 1.f; }
    float ln = sqrtf(nx * nx + ny * ny + nz * nz);
    if (!(ln < PS_EPS || ph < PH_CLOTH)) {
        float4 X = __ldg(prev + i);
        float dx = ex - X.x, dy = ey - X.y, dz = ez - X.z;
        float dn = dx * nx + dy * ny + dz * nz;
        float tx = dx - dn * nx, ty = dy - dn * ny, tz = dz - dn * nz;
        float lt = sqrtf(tx * tx + ty * ty + tz * tz);
        if (!(lt < PS_EPS)) {
            if (lt < sqrtf(PS_S_FRICTION) * d) {
                ex -= tx; ey -= ty; ez -= tz;
            } else {
                float m = fminf(__fdividef(sqrtf(PS_K_FRICTION) * d, lt), 1.f);
                ex -= tx * m; ey -= ty * m; ez -= tz * m;
            }
        }
    }
    pos[i] = make_float4(ex, ey, ez, P.w);
}

// ---- K10: pos[idx].xyz = point, w kept.  32 B/pin. ----
__global__ void __launch_bounds__(kBlock) k_point(float4 *__restrict__ pos, const u32 *__restrict__ pidx, const float *__restrict__ pxyz,
                                                  u32 np) {
    u32 c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= np) return;
    u32 i = pidx[c];
    float4 p = pos[i];
    p.x = pxyz[3 * c]; p.y = pxyz[3 * c + 1]; p.z = pxyz[3 * c + 2];
    pos[i] = p;
}

// ---- K9 (gather form).  The reference computes one delta per constraint, sort_by_key's the 2M (particle,
// delta) pairs, reduce_by_key's and divides by occurences[] (solver.cu:211-230).  Here a CSR built once at
// constraint-add time lists, for every constrained particle, its constraints in exactly that sorted order
// ("a" roles in constraint order, then "b" roles), so one thread sums them in registers: 44 B/constraint
// endpoint of traffic instead of a 2M-element radix sort of 16-byte values per iteration.
// Two kernels because the update is Jacobi: all deltas come from the positions before any is applied. ----
__global__ void __launch_bounds__(kBlock) k_distance_gather(const float4 *__restrict__ pos, float4 *__restrict__ scratch,
                                                            const u32 *__restrict__ csr_particle, const u32 *__restrict__ csr_off,
                                                            const u32 *__restrict__ csr_other, const float *__restrict__ csr_rest,
                                                            u32 num_constrained) {
    u32 k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= num_constrained) return;
    u32 p = csr_particle[k];
    float4 pp = pos[p];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    u32 b = csr_off[k], e = csr_off[k + 1];
    for (u32 t = b; t < e; t++) {
        u32 o = csr_other[t];
        bool is_b = (o >> 31) != 0;
        float4 pq = pos[o & 0x7fffffffu];
        // relPos = p1 - p2 with p1 the constraint's first endpoint (solver_kernel.cuh:46-52)
        float rx = is_b ? pq.x - pp.x : pp.x - pq.x;
        float ry = is_b ? pq.y - pp.y : pp.y - pq.y;
        float rz = is_b ? pq.z - pp.z : pp.z - pq.z;
        float dist = sqrtf(rx * rx + ry * ry + rz * rz);
        if (dist > 0.0001f) {
            float mag = (csr_rest[t] - dist) * .5f;
            float dx = __fdividef(rx, dist) * mag, dy = __fdividef(ry, dist) * mag, dz = __fdividef(rz, dist) * mag;
            if (is_b) { sx += -dx; sy += -dy; sz += -dz; } else { sx += dx; sy += dy; sz += dz; }
        }
    }
    scratch[k] = make_float4(sx, sy, sz, 0.f);
}
__global__ void __launch_bounds__(kBlock) k_distance_apply(float4 *__restrict__ pos, const float4 *__restrict__ scratch,
                                                           const u32 *__restrict__ csr_particle, const u32 *__restrict__ occ,
                                                           u32 num_constrained, float omega) {
    u32 k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= num_constrained) return;
    u32 p = csr_particle[k];
    float4 s = scratch[k];
    float o = (float)occ[p];
    float4 P = pos[p];
    P.x += omega * __fdividef(s.x, o); P.y += omega * __fdividef(s.y, o); P.z += omega * __fdividef(s.z, o);
    pos[p] = P;
}
}  // namespace

static inline u32 cdiv(u32 a, u32 b) { return (a + b - 1) / b; }

void ps_launch_predict(float4 *pos, const float4 *vel, float4 *prev, u32 n, float dt, float3 g, cudaStream_t s, const int *gas_phase) {
    if (!n) return;
    if (gas_phase) k_predict_gas<<<cdiv(n, kBlock), kBlock, 0, s>>>(pos, vel, prev, gas_phase, n, dt, g, PS_GAS_ALPHA);
    else k_predict<<<cdiv(n, kBlock * 2), kBlock, 0, s>>>(pos, vel, prev, n, dt, g);
}
void ps_launch_velocity(const float4 *pos, const float4 *prev, float4 *vel, u32 n, float dt, cudaStream_t s) {
    if (!n) return;
    k_velocity<<<cdiv(n, kBlock * 2), kBlock, 0, s>>>(pos, prev, vel, n, -dt);
}
void ps_launch_collide_world(float4 *pos, const float4 *prev, const int *phase, u32 n, const float *rands6, WorldDesc w, cudaStream_t s) {
    if (!n) return;
    k_collide_world<<<cdiv(n, kBlock), kBlock, 0, s>>>(pos, prev, phase, n, rands6, w);
}
void ps_launch_point(float4 *pos, const u32 *pidx, const float *pxyz, u32 np, cudaStream_t s) {
    if (!np) return;
    k_point<<<cdiv(np, kBlock), kBlock, 0, s>>>(pos, pidx, pxyz, np);
}
void ps_launch_distance(float4 *pos, float4 *scratch, const u32 *csr_particle, const u32 *csr_off, const u32 *csr_other,
                        const float *csr_rest, const u32 *occ, u32 num_constrained, float omega, cudaStream_t s) {
    if (!num_constrained) return;
    k_distance_gather<<<cdiv(num_constrained, kBlock), kBlock, 0, s>>>(pos, scratch, csr_particle, csr_off, csr_other, csr_rest,
                                                                       num_constrained);
    k_distance_apply<<<cdiv(num_constrained, kBlock), kBlock, 0, s>>>(pos, scratch, csr_particle, occ, num_constrained, omega);
}
