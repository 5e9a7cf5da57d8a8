// particlesolver_b200/csrc/ps_shape_kernels.cu — K12: shape-matching constraint of rigid bodies in 3-D, one warp per body.
//
// NOT in the reference's GPU solver: its rigid_body_functor is an empty stub that is never called
// (gpu/src/cuda/solver_kernel.cuh:289-312) and its CPU solver matches shapes in 2-D by a mass-weighted mean angle
// (cpu/src/solver/particle.cpp:15-57).  This is the 3-D constraint of the paper the reference implements (Macklin et al.
// 2014, section 5.1; Mueller et al. 2005): the goal position of particle i is  c + R r_i  where c is the body's current
// centre of mass, r_i its rest offset and R the rotation of the polar decomposition A = R S of the moment matrix
// A = sum_i m_i (x_i - c) r_i^T.  Parity unpinned (no oracle in the reference); tests/test_gpu_shape.py checks it
// against an SVD-based float64 polar decomposition, and its planar restriction against the reference CPU solver's
// angle estimator on rigid motions.
//
// Per warp: lanes stride over the body's particles; centre of mass and the 9 moment sums are reduced with xor-shuffles, so
// every lane ends up holding c and A; the rotation is then extracted redundantly (uniformly) by all lanes with the
// iteration of Mueller, Bender, Chentanez, Macklin 2016 ("A robust method to extract the rotational part of
// deformations"), warm-started from the body's quaternion of the previous call — robust for flat, degenerate and
// inverted configurations, where Newton / eigen-decomposition based polar decompositions break down.
#include "ps_common.cuh"

namespace {
constexpr int kBlock = 128;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

struct Mat3 { float m[3][3]; };  // m[row][col]

__device__ __forceinline__ Mat3 quat_to_mat(float4 q) {  // q = (x, y, z, w), unit
    Mat3 R;
    const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z, wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    R.m[0][0] = 1.f - 2.f * (yy + zz); R.m[0][1] = 2.f * (xy - wz);       R.m[0][2] = 2.f * (xz + wy);
    R.m[1][0] = 2.f * (xy + wz);       R.m[1][1] = 1.f - 2.f * (xx + zz); R.m[1][2] = 2.f * (yz - wx);
    R.m[2][0] = 2.f * (xz - wy);       R.m[2][1] = 2.f * (yz + wx);       R.m[2][2] = 1.f - 2.f * (xx + yy);
    return R;
}
__device__ __forceinline__ float4 quat_mul(float4 a, float4 b) {  // a * b
    return make_float4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
                       a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}

__global__ void __launch_bounds__(kBlock) k_shape_match(float4 *__restrict__ pos, const u32 *__restrict__ body_off, const u32 *__restrict__ body_idx,
                                                        const float4 *__restrict__ rest, float4 *__restrict__ quat, const float *__restrict__ stiff,
                                                        u32 num_bodies, int max_iters) {
    const u32 b = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= num_bodies) return;
    const u32 b0 = body_off[b], b1 = body_off[b + 1];
    // ---- centre of mass ----
    float mx = 0.f, my = 0.f, mz = 0.f, mt = 0.f;
    for (u32 k = b0 + lane; k < b1; k += 32) {
        const float4 p = pos[body_idx[k]];
        const float m = rest[k].w;
        mx += m * p.x; my += m * p.y; mz += m * p.z; mt += m;
    }
    mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz); mt = warp_sum(mt);
    const float inv_m = 1.f / mt;
    const float cx = mx * inv_m, cy = my * inv_m, cz = mz * inv_m;
    // ---- moment matrix A = sum m (x - c) r^T ----
    float a[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (u32 k = b0 + lane; k < b1; k += 32) {
        const float4 p = pos[body_idx[k]];
        const float4 r = rest[k];
        const float dx = r.w * (p.x - cx), dy = r.w * (p.y - cy), dz = r.w * (p.z - cz);
        a[0][0] += dx * r.x; a[0][1] += dx * r.y; a[0][2] += dx * r.z;
        a[1][0] += dy * r.x; a[1][1] += dy * r.y; a[1][2] += dy * r.z;
        a[2][0] += dz * r.x; a[2][1] += dz * r.y; a[2][2] += dz * r.z;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) a[i][j] = warp_sum(a[i][j]);
    // ---- rotation of the polar decomposition, iterated on the quaternion (uniform across the warp) ----
    float4 q = quat[b];
    Mat3 R = quat_to_mat(q);
    for (int it = 0; it < max_iters; it++) {
        // omega = sum_k R_k x A_k / (|sum_k R_k . A_k| + eps), R_k / A_k the k-th columns
        float ox = 0.f, oy = 0.f, oz = 0.f, den = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float rx = R.m[0][k], ry = R.m[1][k], rz = R.m[2][k], ax = a[0][k], ay = a[1][k], az = a[2][k];
            ox += ry * az - rz * ay; oy += rz * ax - rx * az; oz += rx * ay - ry * ax;
            den += rx * ax + ry * ay + rz * az;
        }
        const float s = 1.f / (fabsf(den) + 1e-9f);
        ox *= s; oy *= s; oz *= s;
        const float w = sqrtf(ox * ox + oy * oy + oz * oz);
        if (w < 1e-7f) break;
        float sn, cs;
        sincosf(0.5f * w, &sn, &cs);
        const float k = sn / w;
        q = quat_mul(make_float4(ox * k, oy * k, oz * k, cs), q);
        const float qn = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        q.x *= qn; q.y *= qn; q.z *= qn; q.w *= qn;
        R = quat_to_mat(q);
    }
    if (lane == 0) quat[b] = q;
    // ---- move every member towards its goal position c + R r ----
    const float st = stiff[b];
    for (u32 k = b0 + lane; k < b1; k += 32) {
        const u32 i = body_idx[k];
        const float4 r = rest[k];
        float4 p = pos[i];
        const float gx = cx + R.m[0][0] * r.x + R.m[0][1] * r.y + R.m[0][2] * r.z;
        const float gy = cy + R.m[1][0] * r.x + R.m[1][1] * r.y + R.m[1][2] * r.z;
        const float gz = cz + R.m[2][0] * r.x + R.m[2][1] * r.y + R.m[2][2] * r.z;
        p.x += st * (gx - p.x); p.y += st * (gy - p.y); p.z += st * (gz - p.z);
        pos[i] = p;
    }
}

// World-frame SDF of every rigid-body member for the contact pass: the stored gradient (rest frame) turned by the body's current
// rotation; depth unchanged.  One thread per member; members without SDF data (w < 0) are left alone.
__global__ void __launch_bounds__(256) k_sdf_world(float4 *__restrict__ sdf_world, const float4 *__restrict__ sdf_rest, const u32 *__restrict__ body_idx,
                                                   const u32 *__restrict__ member_body, const float4 *__restrict__ quat, u32 members) {
    const u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k >= members) return;
    const float4 g = sdf_rest[k];
    if (!(g.w >= 0.f)) return;
    const Mat3 R = quat_to_mat(quat[member_body[k]]);
    sdf_world[body_idx[k]] = make_float4(R.m[0][0] * g.x + R.m[0][1] * g.y + R.m[0][2] * g.z, R.m[1][0] * g.x + R.m[1][1] * g.y + R.m[1][2] * g.z,
                                         R.m[2][0] * g.x + R.m[2][1] * g.y + R.m[2][2] * g.z, g.w);
}
}  // namespace

void ps_launch_sdf_world(float4 *sdf_world, const float4 *sdf_rest, const u32 *body_idx, const u32 *member_body, const float4 *quat, u32 members,
                         cudaStream_t s) {
    if (!members) return;
    k_sdf_world<<<(members + 255) / 256, 256, 0, s>>>(sdf_world, sdf_rest, body_idx, member_body, quat, members);
}

void ps_launch_shape_match(float4 *pos, const u32 *body_off, const u32 *body_idx, const float4 *rest, float4 *quat, const float *stiff, u32 num_bodies,
                           int max_iters, cudaStream_t s) {
    if (!num_bodies) return;
    const u32 warps_per_block = kBlock / 32;
    k_shape_match<<<(num_bodies + warps_per_block - 1) / warps_per_block, kBlock, 0, s>>>(pos, body_off, body_idx, rest, quat, stiff, num_bodies, max_iters);
}
