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

// ------------------------------------------------------------------ fluid neighbour walk ------------------------------------------------------------------
// Shared by K6 and K7.  One thread per sorted slot, one warp = 32 consecutive slots (= a run of x-adjacent
// particles of one or two grid rows).  Two things keep the warp's issue slots busy:
//   (1) per-particle row pruning: for stencil row (dy,dz) the x-extent that can hold a neighbour follows from the
//       particle's own position inside its cell, ext = sqrt(H^2 - dymin^2 - dzmin^2) — ~200 candidates per
//       particle instead of the ~370 of the full 9^3 stencil.  The pruning is conservative (eps margin), so the
//       accepted neighbour sequence — and with it the 500-cap and the summation order — is exactly the reference's;
//   (2) accept/interact split: the distance test runs over all candidates, accepted neighbours (~40 % of them) are
//       staged in a per-thread shared-memory queue of Q float4 entries (r.x, r.y, r.z, r2 | j) and the expensive
//       interaction body runs over full queues with every lane active, instead of under a divergent branch.
// The row loop is warp-uniform (all lanes walk the same (dz,dy) sequence; the inner loop runs to the warp's longest
// range), which keeps the vote that triggers a queue flush legal.
constexpr int kQ = 8;
constexpr unsigned kFull = 0xffffffffu;

template <bool STORE_J, class Body>
__device__ __forceinline__ u32 walk_fluid_neighbours(const GridDesc &g, const StencilDesc &st, const u32 *__restrict__ cell_begin,
                                                     const float4 *__restrict__ spos, bool act, u32 i, float4 pi, float4 (*q)[kBlock],
                                                     Body &&body) {
    const int tid = threadIdx.x;
    const int rad = st.rad, w = 2 * st.rad + 1;
    const float relx = pi.x - g.ox, rely = pi.y - g.oy, relz = pi.z - g.oz;
    const int3 gp = ps_grid_pos(g, pi.x, pi.y, pi.z);
    // margin: covers the approximate divide of the cell assignment and coordinate rounding (ulp(1000) = 6e-5)
    const float eps = 1e-3f + 2e-6f * fmaxf(fabsf(relx), fmaxf(fabsf(rely), fabsf(relz)));
    const float inv_cx = __fdividef(1.f, g.cx);
    // distance from the particle to the lower / upper face of its own cell (clamped: the divide is approximate)
    const float fy0 = fmaxf(rely - (float)gp.y * g.cy, 0.f), fy1 = fmaxf((float)(gp.y + 1) * g.cy - rely, 0.f);
    const float fz0 = fmaxf(relz - (float)gp.z * g.cz, 0.f), fz1 = fmaxf((float)(gp.z + 1) * g.cz - relz, 0.f);
    u32 nn = 0;
    int cnt = 0;

    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < kQ; k++)
            if (k < cnt) body(q[k][tid]);
        cnt = 0;
    };

    for (int dz = -rad; dz <= rad; dz++) {
        const u32 zrow = ((u32)(gp.z + dz) & g.mz) * g.gy;
        float dzmin = dz == 0 ? 0.f : (dz > 0 ? fz1 + (float)(dz - 1) * g.cz : fz0 + (float)(-dz - 1) * g.cz);
        dzmin = fmaxf(dzmin - eps, 0.f);
        const float remz = PS_H2 - dzmin * dzmin;
        for (int dy = -rad; dy <= rad; dy++) {
            if (st.xr[(dz + rad) * w + (dy + rad)] < 0) continue;  // uniform: row out of reach for every particle
            float dymin = dy == 0 ? 0.f : (dy > 0 ? fy1 + (float)(dy - 1) * g.cy : fy0 + (float)(-dy - 1) * g.cy);
            dymin = fmaxf(dymin - eps, 0.f);
            const float rem = remz - dymin * dymin;
            u32 b0 = 0, len0 = 0, b1 = 0, len1 = 0;
            if (act && rem >= 0.f) {
                const float ext = sqrtf(rem) + eps;
                int lo = (int)floorf((relx - ext) * inv_cx), hi = (int)floorf((relx + ext) * inv_cx);
                lo = max(min(lo, gp.x), gp.x - rad);
                hi = min(max(hi, gp.x), gp.x + rad);
                const u32 row = (zrow + ((u32)(gp.y + dy) & g.my)) * g.gx;
                const u32 lw = (u32)lo & g.mx, hw = (u32)hi & g.mx;
                if (lw <= hw) {
                    b0 = __ldg(cell_begin + row + lw);
                    len0 = __ldg(cell_begin + row + hw + 1) - b0;
                } else {  // the row wraps around the power-of-two grid: [lw, gx) then [0, hw]
                    b0 = __ldg(cell_begin + row + lw);
                    len0 = __ldg(cell_begin + row + g.mx + 1) - b0;
                    b1 = __ldg(cell_begin + row);
                    len1 = __ldg(cell_begin + row + hw + 1) - b1;
                }
            }
#pragma unroll 1
            for (int seg = 0; seg < 2; seg++) {
                const u32 b = seg ? b1 : b0, len = seg ? len1 : len0;
                const u32 maxlen = __reduce_max_sync(kFull, len);
#pragma unroll 1
                for (u32 t = 0; t < maxlen; t++) {
                    if (t < len) {
                        const u32 j = b + t;
                        const float4 pj = __ldg(spos + j);
                        const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                        const float r2 = rx * rx + ry * ry + rz * rz;
                        if (r2 < PS_H2 && j != i && nn < PS_MAX_NEIGHBORS) {
                            q[cnt][tid] = make_float4(rx, ry, rz, STORE_J ? __uint_as_float(j) : r2);
                            cnt++;
                            nn++;
                        }
                    }
                    if (__any_sync(kFull, cnt == kQ)) flush();
                }
            }
        }
    }
    flush();
    return nn;
}

// ------------------------------------------------------------------ K6: lambda ------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_find_lambdas(float *__restrict__ lambda, u32 *__restrict__ num_neighbors,
                                                         const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                         const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                         const u32 *__restrict__ cell_begin, const float *__restrict__ ros, u32 n,
                                                         u32 n_owned, GridDesc g, StencilDesc st, int zero_nonfluid) {
    __shared__ float4 q[kQ][kBlock];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    bool act = i < n;
    u32 orig = 0;
    if (act) {
        if (sphase[i] != PH_FLUID) {
            if (zero_nonfluid) lambda[i] = 0.f;
            act = false;
        } else {
            orig = index[i];
            act = orig < n_owned;  // ghost copy of a neighbour slab's particle: its owner computes it
        }
    }
    if (!__any_sync(kFull, act)) return;
    const float4 pi = act ? spos[i] : make_float4(g.ox, g.oy, g.oz, 0.f);
    const float ro0 = act ? ros[orig] : 1.f;
    const float inv_ro0 = __fdividef(1.f, ro0);
    const float cs = -PS_SPIKY * inv_ro0;

    float ro = 0.f, denom = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    const u32 nn = walk_fluid_neighbours<false>(g, st, cell_begin, spos, act, i, pi, q, [&](const float4 e) {
        const float r2 = e.w;
        const float inv_r = rsqrtf(r2);
        const float rlen = r2 * inv_r;  // sqrt(r2); r2 == 0 gives NaN here and is handled below
        const float hm2 = PS_H2 - r2;
        ro += hm2 * hm2 * hm2;
        if (rlen >= 0.0001f) {  // false for NaN as well: coincident particles contribute no gradient
            const float hm = PS_H - rlen;
            const float c = (cs * hm * hm) * inv_r;  // spikyGrad / rho0 = r * c
            gx += e.x * c; gy += e.y * c; gz += e.z * c;
            denom += (c * c) * r2;
        }
    });
    if (!act) return;
    const float inv_w = __fdividef(1.f, sw[i]);
    ro = (ro + PS_H6) * (PS_POLY6 * inv_w);  // + self term poly6(0) = POLY6 * H^6 (integration_kernel.cuh:589)
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
    __shared__ float4 q[kQ][kBlock];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    bool act = i < n && sphase[i] == PH_FLUID;
    u32 orig = 0;
    if (act) {
        orig = index[i];
        act = orig < n_owned;
    }
    if (!__any_sync(kFull, act)) return;
    const float4 pi = act ? spos[i] : make_float4(g.ox, g.oy, g.oz, 0.f);
    const float li = act ? lambda[i] : 0.f;
    // s_corr = -K_P * (poly6(r) / poly6(dq*H))^4 ; the POLY6 factors cancel (integration_kernel.cuh:630-634)
    const float term2 = PS_H2 - (PS_DQ_P * PS_DQ_P * PS_H2);
    const float inv_den = __fdividef(1.f, term2 * term2 * term2);

    float dx = 0.f, dy = 0.f, dz = 0.f;
    const u32 nn = walk_fluid_neighbours<true>(g, st, cell_begin, spos, act, i, pi, q, [&](const float4 e) {
        const float lj = __ldg(lambda + __float_as_uint(e.w));
        const float r2 = e.x * e.x + e.y * e.y + e.z * e.z;
        const float inv_r = rsqrtf(r2);
        const float rlen = r2 * inv_r;
        const float hm2 = PS_H2 - r2;
        const float qq = (hm2 * hm2 * hm2) * inv_den;
        const float q2 = qq * qq;
        const float s = li + lj + (-PS_K_P * q2 * q2);
        if (rlen >= 0.0001f) {
            const float hm = PS_H - rlen;
            const float c = s * ((-PS_SPIKY * hm * hm) * inv_r);
            dx += e.x * c; dy += e.y * c; dz += e.z * c;
        } else {  // coincident: the reference nudges along +y, (0,EPS,0,0) * -SPIKY * (H-r)^2 (:625-626)
            const float rl = (r2 > 0.f) ? rlen : 0.f;
            const float hm = PS_H - rl;
            dy += s * (PS_EPS * -PS_SPIKY * hm * hm);
        }
    });
    if (!act) return;
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
