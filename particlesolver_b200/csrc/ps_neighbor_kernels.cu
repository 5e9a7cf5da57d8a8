// particlesolver_b200/csrc/ps_neighbor_kernels.cu — the neighbour kernels of the solver step:
//   K5  contact + friction  (reference collideD/collideCell,            integration_kernel.cuh:303-462)
//   K6  PBF lambda          (reference findLambdasD/collideCellRadius,  integration_kernel.cuh:480-593)
//   K7  PBF delta-p + s_corr(reference solveFluidsD,                    integration_kernel.cuh:596-642)
//
// What is kept from the reference: one thread per SORTED slot; the neighbour set, its traversal order
// (cells z,y,x outermost-to-innermost, ascending sorted slot inside a cell), the 500-neighbour cap, and the
// arithmetic of every term.  What is new:
//   * no neighbour lists (the reference materialises 500 slots = 2 KB per particle, integration.cu:70 —
//     8 GB at 1M particles); both PBF passes walk the grid and keep everything in registers;
//   * the hash is x-fastest, so the 2r+1 cells of one (dy,dz) row are one CONTIGUOUS range of the sorted
//     arrays: [cell_begin[row+lo], cell_begin[row+hi+1]).  A 9^3 = 729-probe stencil becomes <= 81 range
//     lookups in a dense lower-bound table, and rows/cells that cannot hold a particle within H are pruned
//     (conservatively, see StencilDesc) without changing the neighbour sequence;
//   * candidate positions are read as coalesced float4 through L1 (adjacent threads are adjacent in space,
//     so a warp's 32 loads fall in a handful of 128-byte lines).
// These kernels are bound by FP32 issue and L1 bandwidth, not HBM: ~370 candidate tests and
// ~140 interactions per particle against 36-64 compulsory bytes.
#include "ps_common.cuh"

namespace {
constexpr int kBlock = 128;

// Visit, in the reference's order, every sorted slot j whose cell lies in the stencil rows around gp.
// f(j) is called for each candidate (including j == self; callers test that).
template <class F>
__device__ __forceinline__ void for_each_candidate(const GridDesc &g, const StencilDesc &st, const u32 *__restrict__ cell_begin, int3 gp,
                                                   F &&f) {
    const int rad = st.rad, w = 2 * st.rad + 1;
    for (int dz = -rad; dz <= rad; dz++) {
        const u32 zrow = ((u32)(gp.z + dz) & g.mz) * g.gy;
        for (int dy = -rad; dy <= rad; dy++) {
            const int xr = st.xr[(dz + rad) * w + (dy + rad)];
            if (xr < 0) continue;
            const u32 row = (zrow + ((u32)(gp.y + dy) & g.my)) * g.gx;
            const u32 lo = (u32)(gp.x - xr) & g.mx, hi = (u32)(gp.x + xr) & g.mx;
            if (lo <= hi) {
                u32 b = __ldg(cell_begin + row + lo), e = __ldg(cell_begin + row + hi + 1);
                for (u32 j = b; j < e; j++) f(j);
            } else {  // the row wraps around the power-of-two grid: [lo, gx) then [0, hi]
                u32 b = __ldg(cell_begin + row + lo), e = __ldg(cell_begin + row + g.mx + 1);
                for (u32 j = b; j < e; j++) f(j);
                b = __ldg(cell_begin + row);
                e = __ldg(cell_begin + row + hi + 1);
                for (u32 j = b; j < e; j++) f(j);
            }
        }
    }
}

// ------------------------------------------------------------------ K6: lambda ------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_find_lambdas(float *__restrict__ lambda, u32 *__restrict__ num_neighbors,
                                                         const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                         const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                         const u32 *__restrict__ cell_begin, const float *__restrict__ ros, u32 n,
                                                         u32 n_owned, GridDesc g, StencilDesc st, int zero_nonfluid) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    if (sphase[i] != PH_FLUID) {
        if (zero_nonfluid) lambda[i] = 0.f;
        return;
    }
    const u32 orig = index[i];
    if (orig >= n_owned) return;  // ghost copy of a neighbour slab's particle: its owner computes it
    const float4 pi = spos[i];
    const float inv_w = __fdividef(1.f, sw[i]);
    const float ro0 = ros[orig];
    const float inv_ro0 = __fdividef(1.f, ro0);
    const int3 gp = ps_grid_pos(g, pi.x, pi.y, pi.z);

    u32 nn = 0;
    float ro = 0.f, denom = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    for_each_candidate(g, st, cell_begin, gp, [&](u32 j) {
        const float4 pj = __ldg(spos + j);
        const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
        const float r2 = rx * rx + ry * ry + rz * rz;
        if (r2 < PS_H2 && j != i && nn < PS_MAX_NEIGHBORS) {
            nn++;
            const float inv_r = rsqrtf(r2);
            const float rlen = r2 * inv_r;  // sqrt(r2); r2 == 0 gives NaN here and is handled below
            const float hm2 = PS_H2 - r2;
            ro += (PS_POLY6 * hm2 * hm2 * hm2) * inv_w;
            if (rlen >= 0.0001f) {  // false for NaN as well: coincident particles contribute no gradient
                const float hm = PS_H - rlen;
                const float c = (-PS_SPIKY * hm * hm) * inv_r * inv_ro0;  // spikyGrad = r * c
                const float sx = rx * c, sy = ry * c, sz = rz * c;
                gx -= sx; gy -= sy; gz -= sz;
                denom += sx * sx + sy * sy + sz * sz;
            }
        }
    });
    ro += (PS_POLY6 * PS_H6) * inv_w;
    denom += gx * gx + gy * gy + gz * gz;
    lambda[i] = -__fdividef(ro * inv_ro0 - 1.f, denom + PS_RELAX);
    num_neighbors[i] = nn;
}

// ------------------------------------------------------------------ K7: delta p ------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_solve_fluids(float4 *__restrict__ pos, const float *__restrict__ lambda,
                                                         const float4 *__restrict__ spos, const int *__restrict__ sphase,
                                                         const u32 *__restrict__ index, const u32 *__restrict__ cell_begin,
                                                         const float *__restrict__ ros, u32 n, u32 n_owned, GridDesc g, StencilDesc st,
                                                         float omega) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    if (sphase[i] != PH_FLUID) return;
    const u32 orig = index[i];
    if (orig >= n_owned) return;
    const float4 pi = spos[i];
    const float li = lambda[i];
    const int3 gp = ps_grid_pos(g, pi.x, pi.y, pi.z);
    // s_corr = -K_P * (poly6(r) / poly6(dq*H))^4 ; the POLY6 factors cancel (integration_kernel.cuh:630-634)
    const float term2 = PS_H2 - (PS_DQ_P * PS_DQ_P * PS_H2);
    const float inv_den = __fdividef(1.f, term2 * term2 * term2);

    u32 nn = 0;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    for_each_candidate(g, st, cell_begin, gp, [&](u32 j) {
        const float4 pj = __ldg(spos + j);
        const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
        const float r2 = rx * rx + ry * ry + rz * rz;
        if (r2 < PS_H2 && j != i && nn < PS_MAX_NEIGHBORS) {
            nn++;
            const float lj = __ldg(lambda + j);
            const float inv_r = rsqrtf(r2);
            const float rlen = r2 * inv_r;
            const float hm2 = PS_H2 - r2;
            const float q = (hm2 * hm2 * hm2) * inv_den;
            const float q2 = q * q;
            const float s = li + lj + (-PS_K_P * q2 * q2);
            if (rlen >= 0.0001f) {
                const float hm = PS_H - rlen;
                const float c = s * ((-PS_SPIKY * hm * hm) * inv_r);
                dx += rx * c; dy += ry * c; dz += rz * c;
            } else {  // coincident: the reference nudges along +y, (0,EPS,0,0) * -SPIKY * (H-r)^2 (:625-626)
                const float rl = (r2 > 0.f) ? rlen : 0.f;
                const float hm = PS_H - rl;
                dy += s * (PS_EPS * -PS_SPIKY * hm * hm);
            }
        }
    });
    const float inv_div = __fdividef(omega, ros[orig] + (float)nn);
    float4 P = pos[orig];
    P.x += dx * inv_div; P.y += dy * inv_div; P.z += dz * inv_div;
    pos[orig] = P;
}

// ------------------------------------------------------------------ K5: contacts + friction ------------------------------------------------------------------
// 27 cells = 9 rows of 3.  Two sweeps over the same rows: the first only counts (every per-neighbour term is
// divided by the final neighbour count, and friction is non-linear in it), the second accumulates.
__device__ __forceinline__ float ps_scaled_w(float w, float y) {
    // sW = w != 0 ? 1 / ((1/w) * exp(-y)) : w   (integration_kernel.cuh:403,425)
    return (w != 0.f) ? __fdividef(1.f, __fdividef(1.f, w) * __expf(-y)) : w;
}

__global__ void __launch_bounds__(kBlock) k_collide(float4 *__restrict__ pos, const float4 *__restrict__ prev,
                                                    const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                    const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                    const u32 *__restrict__ cell_begin, u32 *__restrict__ num_neighbors, u32 n, u32 n_owned,
                                                    GridDesc g, StencilDesc st, float radius) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int phase = sphase[i];
    if (phase < PH_CLOTH) return;
    const u32 orig = index[i];
    if (orig >= n_owned) return;
    const float4 pi = spos[i];
    const int3 gp = ps_grid_pos(g, pi.x, pi.y, pi.z);
    const float collide_dist = radius * 2.001f;
    const float collide_dist2 = collide_dist * collide_dist;

    u32 nn = 0;
    for_each_candidate(g, st, cell_begin, gp, [&](u32 j) {
        if (j == i) return;
        const int phase2 = __ldg(sphase + j);
        if (phase > PH_SOLID && phase == phase2) return;
        const float4 pj = __ldg(spos + j);
        const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
        if (rx * rx + ry * ry + rz * rz < collide_dist2 && nn < PS_MAX_NEIGHBORS) nn++;
    });
    num_neighbors[i] = nn;
    float dxs = 0.f, dys = 0.f, dzs = 0.f;
    if (nn) {
        const float w = sw[i];
        const float sW = ps_scaled_w(w, pi.y);
        const float4 pp = __ldg(prev + orig);
        const float fn = (float)nn;
        u32 seen = 0;
        for_each_candidate(g, st, cell_begin, gp, [&](u32 j) {
            if (j == i || seen >= nn) return;
            const int phase2 = __ldg(sphase + j);
            if (phase > PH_SOLID && phase == phase2) return;
            const float4 pj = __ldg(spos + j);
            const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
            const float d2 = rx * rx + ry * ry + rz * rz;
            if (!(d2 < collide_dist2)) return;
            seen++;
            const float w2 = __ldg(sw + j);
            const float dist = sqrtf(d2);
            const float mag = dist - collide_dist;
            float colW = w, colW2 = w2;
            const bool both_solid = phase >= PH_SOLID && phase2 >= PH_SOLID;
            if (both_solid) {
                colW = sW;
                colW2 = ps_scaled_w(w2, pi.y);
            }
            const float wsum = colW + colW2;
            const float sd = __fdividef(__fdividef(mag, wsum), dist);
            const float px = rx * sd, py = ry * sd, pz = rz * sd;  // dp
            const float d1x = __fdividef(-colW * px, fn), d1y = __fdividef(-colW * py, fn), d1z = __fdividef(-colW * pz, fn);
            dxs += d1x; dys += d1y; dzs += d1z;
            if (!both_solid) return;
            const float d2x = __fdividef(colW2 * px, fn), d2y = __fdividef(colW2 * py, fn), d2z = __fdividef(colW2 * pz, fn);
            const float4 pp2 = __ldg(prev + __ldg(index + j));
            const float inv = rsqrtf(d2);
            const float nx = rx * inv, ny = ry * inv, nz = rz * inv;
            // [sic] the reference's second term starts from prevPos of i, not pos2 (integration_kernel.cuh:447)
            const float ex = (pi.x + d1x - pp.x) - (pp.x + d2x - pp2.x);
            const float ey = (pi.y + d1y - pp.y) - (pp.y + d2y - pp2.y);
            const float ez = (pi.z + d1z - pp.z) - (pp.z + d2z - pp2.z);
            const float dn = ex * nx + ey * ny + ez * nz;
            const float tx = ex - dn * nx, ty = ey - dn * ny, tz = ez - dn * nz;
            const float lt = sqrtf(tx * tx + ty * ty + tz * tz);
            if (lt < PS_EPS) return;
            if (lt < PS_S_FRICTION * dist) {
                dxs -= __fdividef(tx * colW, wsum); dys -= __fdividef(ty * colW, wsum); dzs -= __fdividef(tz * colW, wsum);
            } else {
                const float m = fminf(__fdividef(PS_K_FRICTION * dist, lt), 1.f);
                dxs -= tx * m; dys -= ty * m; dzs -= tz * m;
            }
        });
    }
    pos[orig] = make_float4(pi.x + dxs, pi.y + dys, pi.z + dzs, 1.0f);
}
}  // namespace

static inline u32 cdiv(u32 a, u32 b) { return (a + b - 1) / b; }

void ps_launch_collide(float4 *pos, const float4 *prev, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                       const u32 *cell_begin, u32 *num_neighbors, u32 n, u32 n_owned, GridDesc g, float radius, cudaStream_t s) {
    if (!n) return;
    StencilDesc st;  // 3x3x3: every row keeps its full extent (contact radius 2.001r slightly exceeds one cell)
    st.rad = 1;
    for (int k = 0; k < 9; k++) st.xr[k] = 1;
    k_collide<<<cdiv(n, kBlock), kBlock, 0, s>>>(pos, prev, spos, sw, sphase, index, cell_begin, num_neighbors, n, n_owned, g, st, radius);
}

void ps_launch_find_lambdas(float *lambda, u32 *num_neighbors, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                            const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, GridDesc g, const StencilDesc &st,
                            bool zero_nonfluid, cudaStream_t s) {
    if (!n) return;
    k_find_lambdas<<<cdiv(n, kBlock), kBlock, 0, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned, g, st,
                                                      zero_nonfluid ? 1 : 0);
}

void ps_launch_solve_fluids(float4 *pos, const float *lambda, const float4 *spos, const int *sphase, const u32 *index,
                            const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, GridDesc g, const StencilDesc &st, float omega,
                            cudaStream_t s) {
    if (!n) return;
    k_solve_fluids<<<cdiv(n, kBlock), kBlock, 0, s>>>(pos, lambda, spos, sphase, index, cell_begin, ros, n, n_owned, g, st, omega);
}
