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
#include <cuda/std/type_traits>
#include "ps_common.cuh"

namespace {
#ifndef PS_FBLOCK
#define PS_FBLOCK 128
#endif
#ifndef PS_KQ
#define PS_KQ 8
#endif
constexpr int kBlock = PS_FBLOCK;  // tuning knobs of the neighbour kernels (build variants: scripts/bench_variants.sh)

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
// particles of one or two grid rows).  Three things keep the warp's issue slots busy:
//   (1) per-particle row pruning: for stencil row (dy,dz) the x-extent that can hold a neighbour follows from the
//       particle's own position inside its cell, ext = sqrt(H^2 - dymin^2 - dzmin^2) — ~200 candidates per
//       particle instead of the ~370 of the full 9^3 stencil.  The pruning is conservative (eps margin), so the
//       accepted neighbour sequence — and with it the 500-cap and the summation order — is exactly the reference's;
//   (2) per-lane segment lists: the row ranges are short (2.7 candidates on average) and their lengths differ from
//       lane to lane, so walking them row by row in lock-step leaves 45 % of the lanes idle (profiles/r1b).  Instead
//       each lane first writes the non-empty ranges of one z-slab of the stencil (<= 9 rows) to a small list in
//       shared memory, then all lanes walk their own lists in one flat loop whose trip count is the warp's largest
//       candidate total;
//   (3) accept/interact split: the distance test runs over all candidates, accepted neighbours (~65 % of them) are
//       staged in a per-thread shared-memory queue of kQ float4 entries (r.x, r.y, r.z, r2 | j) and the expensive
//       interaction body runs over full queues with (nearly) every lane active, instead of under a divergent branch.
// All loops that contain a vote are warp-uniform.
constexpr int kQ = PS_KQ;
constexpr unsigned kFull = 0xffffffffu;

// Shared memory is carved out of the same 256 KB as L1, and the candidate gather lives on L1 hits: every byte counts.
// Per thread: the queue (kQ entries of 16 B, or of 4 B when only the neighbour's slot is staged, PS_QJ) + the segment
// list of one stencil slab: 2*(2*rad+1) entries of (u32 begin, u16 length).
#ifndef PS_QJ
#define PS_QJ 1  // measured on B200 (1M-particle fluid): 4-byte queue entries + reload beat 16-byte entries by 10 % (occupancy, L1)
#endif
typedef unsigned short u16;
static inline size_t fluid_smem_bytes(int rad) {
    return (size_t)kQ * kBlock * (PS_QJ ? sizeof(u32) : sizeof(float4)) + (size_t)2 * (2 * rad + 1) * kBlock * (sizeof(u32) + sizeof(u16));
}

template <int RAD, bool STORE_J, class Body>
__device__ __forceinline__ u32 walk_fluid_neighbours(const GridDesc &g, const StencilDesc &st, const u32 *__restrict__ cell_begin,
                                                     const float4 *__restrict__ spos, bool act, u32 i, float4 pi, float4 *smem,
                                                     Body &&body) {
    const int tid = threadIdx.x;
    const int rad = RAD ? RAD : st.rad;
    const int nseg = 2 * (2 * rad + 1);  // list capacity per slab: every row may wrap into two ranges
#if PS_QJ
    u32(*q)[kBlock] = reinterpret_cast<u32(*)[kBlock]>(smem);
    u32 *seg_b = reinterpret_cast<u32 *>(smem) + kQ * kBlock + tid;  // entry s of this lane at seg_b[s * kBlock]
#else
    float4(*q)[kBlock] = reinterpret_cast<float4(*)[kBlock]>(smem);
    u32 *seg_b = reinterpret_cast<u32 *>(smem + kQ * kBlock) + tid;  // entry s of this lane at seg_b[s * kBlock]
#endif
    u16 *seg_len = reinterpret_cast<u16 *>(seg_b - tid + nseg * kBlock) + tid;
    const float relx = pi.x - g.ox, rely = pi.y - g.oy, relz = pi.z - g.oz;
    const int3 gp = ps_grid_pos(g, pi.x, pi.y, pi.z);
    // margin: covers the approximate divide of the cell assignment and coordinate rounding (ulp(1000) = 6e-5)
    const float eps = 1e-3f + 2e-6f * fmaxf(fabsf(relx), fmaxf(fabsf(rely), fabsf(relz)));
    const float inv_cx = __fdividef(1.f, g.cx);
    // distance from the particle to the lower / upper face of its own cell (clamped: the divide is approximate)
    const float fy0 = fmaxf(rely - (float)gp.y * g.cy, 0.f), fy1 = fmaxf((float)(gp.y + 1) * g.cy - rely, 0.f);
    const float fz0 = fmaxf(relz - (float)gp.z * g.cz, 0.f), fz1 = fmaxf((float)(gp.z + 1) * g.cz - relz, 0.f);
    auto dmin2 = [&](int d, float f0, float f1, float c) {  // squared distance to the slab of cells at offset d, minus margin
        float m = d == 0 ? 0.f : (d > 0 ? f1 + (float)(d - 1) * c : f0 + (float)(-d - 1) * c);
        m = fmaxf(m - eps, 0.f);
        return m * m;
    };
    float dy2[2 * RAD + 1];
    if (RAD) {
#pragma unroll
        for (int d = 0; d < 2 * RAD + 1; d++) dy2[d] = dmin2(d - RAD, fy0, fy1, g.cy);
    }
    u32 nn = 0;
    int cnt = 0;

    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < kQ; k++)
            if (k < cnt) {
#if PS_QJ
                const u32 j = q[k][tid];
                const float4 pj = __ldg(spos + j);  // an L1 hit: the line was gathered a few instructions ago
                const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                body(make_float4(rx, ry, rz, STORE_J ? __uint_as_float(j) : rx * rx + ry * ry + rz * rz));
#else
                body(q[k][tid]);
#endif
            }
        cnt = 0;
    };

#pragma unroll 1
    for (int dz = -rad; dz <= rad; dz++) {
        // ---- phase 1: this lane's candidate ranges of the slab, compacted into its list ----
        const u32 rowmask = st.rowmask[dz + rad];  // uniform: rows no particle can reach
        const u32 zrow = ((u32)(gp.z + dz) & g.mz) * g.gy;
        const float remz = PS_H2 - dmin2(dz, fz0, fz1, g.cz);
        u32 nlist = 0, total = 0;
        auto push = [&](u32 b, u32 len) {
            if (len) {
                if (len > 0xffffu || nlist >= (u32)nseg) __trap();  // > 65535 particles in one row range: not a particle system
                seg_b[nlist * kBlock] = b;
                seg_len[nlist * kBlock] = (u16)len;
                nlist++;
                total += len;
            }
        };
        auto do_row = [&](int dyi) {
            if (!((rowmask >> dyi) & 1u)) return;
            const int dy = dyi - rad;
            const float rem = remz - (RAD ? dy2[RAD ? dyi : 0] : dmin2(dy, fy0, fy1, g.cy));
            u32 b0 = 0, len0 = 0, row = 0, hw = 0;
            bool wrap = false;
            if (act && rem >= 0.f) {
                const float ext = sqrtf(rem) + eps;
                int lo = (int)floorf((relx - ext) * inv_cx), hi = (int)floorf((relx + ext) * inv_cx);
                lo = max(min(lo, gp.x), gp.x - rad);
                hi = min(max(hi, gp.x), gp.x + rad);
                row = (zrow + ((u32)(gp.y + dy) & g.my)) * g.gx;
                const u32 lw = (u32)lo & g.mx;
                hw = (u32)hi & g.mx;
                wrap = lw > hw;  // the row wraps around the power-of-two grid: [lw, gx) then [0, hw]
                b0 = __ldg(cell_begin + (row + lw));
                len0 = __ldg(cell_begin + (row + (wrap ? g.mx : hw) + 1u)) - b0;
            }
            push(b0, len0);
            if (__any_sync(kFull, wrap)) {
                u32 b1 = 0, len1 = 0;
                if (wrap) {
                    b1 = __ldg(cell_begin + row);
                    len1 = __ldg(cell_begin + (row + hw + 1u)) - b1;
                }
                push(b1, len1);
            }
        };
        if (RAD) {
#pragma unroll
            for (int dyi = 0; dyi < 2 * RAD + 1; dyi++) do_row(dyi);
        } else {
#pragma unroll 1
            for (int dyi = 0; dyi <= 2 * rad; dyi++) do_row(dyi);
        }
        // ---- phase 2: flat walk over the list; the trip count is the warp's largest candidate total ----
        const u32 maxtotal = __reduce_max_sync(kFull, total);
        if (maxtotal == 0) continue;
        const bool capped = __any_sync(kFull, nn + total > PS_MAX_NEIGHBORS);  // the 500-neighbour cap can bite in this slab
        // Software pipeline: the positions of candidates t+1 and t+2 are requested before candidate t is tested, so each
        // lane keeps two gathers in flight (at 8 warps per scheduler one was not enough to cover an L2 hit).
        u32 jb = 0, rem_seg = 0, li = 0;
        auto next_j = [&]() {
            if (rem_seg == 0) {
                jb = seg_b[li * kBlock];
                rem_seg = seg_len[li * kBlock];
                li++;
            }
            rem_seg--;
            return jb++;
        };
        u32 j0 = 0, j1 = 0;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (total > 0) { j0 = next_j(); p0 = __ldg(spos + j0); }
        if (total > 1) { j1 = next_j(); p1 = __ldg(spos + j1); }
        auto walk = [&](auto capped_c) {
            auto test = [&](const u32 j, const float4 pj) {
                const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                const float r2 = rx * rx + ry * ry + rz * rz;
                if (r2 < PS_H2 && j != i && (!decltype(capped_c)::value || nn < PS_MAX_NEIGHBORS)) {
#if PS_QJ
                    q[cnt][tid] = j;
#else
                    q[cnt][tid] = make_float4(rx, ry, rz, STORE_J ? __uint_as_float(j) : r2);
#endif
                    cnt++;
                    nn++;
                }
            };
#pragma unroll 1
            for (u32 t = 0; t < maxtotal; t += 2) {
                if (t < total) {
                    const u32 j = j0;
                    const float4 pj = p0;
                    if (t + 2 < total) { j0 = next_j(); p0 = __ldg(spos + j0); }
                    test(j, pj);
                }
                if (__any_sync(kFull, cnt == kQ)) flush();
                if (t + 1 < total) {
                    const u32 j = j1;
                    const float4 pj = p1;
                    if (t + 3 < total) { j1 = next_j(); p1 = __ldg(spos + j1); }
                    test(j, pj);
                }
                if (__any_sync(kFull, cnt == kQ)) flush();
            }
        };
        if (capped) walk(cuda::std::true_type{});
        else walk(cuda::std::false_type{});
    }
    flush();
    return nn;
}

// ------------------------------------------------------------------ slot mapping ------------------------------------------------------------------
// Which sorted slot a thread works on.  The sorted order is x-fastest, so a CTA of consecutive slots is a 60-unit-long
// pencil of particles whose stencil footprint (~110 KB of positions) thrashes L1 (profiles/r1c: 63 % hit rate, 33 % of
// all stall samples waiting on the candidate gather).  The brick mapping gives each CTA the particles of a compact
// brick of kBrickX x kBrickY x kBrickZ CELLS instead (16 x 2 x 2 world units at the reference's cell size): the brick's
// kBrickY*kBrickZ grid rows are kBrickY*kBrickZ contiguous slices of the sorted arrays, enumerated row by row.  The
// footprint per thread drops ~4x and the working set of one stencil slab fits L1.  Every lane is independent in the
// walk (per-lane segment lists), so any slot -> lane assignment is legal; results do not depend on it.
#ifndef PS_FMAP
#define PS_FMAP 0
#endif
constexpr int kBrickX = 32, kBrickY = 4, kBrickZ = 4, kBrickRows = kBrickY * kBrickZ;

// calls f(slot, have) for every slot of the calling CTA's share, warp-uniformly (have == false pads the last warp)
template <class F>
__device__ __forceinline__ void for_each_slot(const GridDesc &g, const u32 *__restrict__ cell_begin, u32 n, F &&f) {
#if PS_FMAP == 0
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    f(i, i < n);
#else
    __shared__ u32 s_b[kBrickRows], s_pre[kBrickRows + 1];
    const int tid = threadIdx.x;
    const u32 bw = min((u32)kBrickX, g.gx);
    if (tid < 32) {  // one warp: the brick's rows, their slices of the sorted arrays and an exclusive scan of the slice lengths
        u32 b = 0, len = 0;
        if (tid < kBrickRows) {
            const u32 cy = blockIdx.y * kBrickY + (tid % kBrickY), cz = blockIdx.z * kBrickZ + (tid / kBrickY);
            const u32 row = (cz * g.gy + cy) * g.gx + blockIdx.x * bw;
            b = __ldg(cell_begin + row);
            len = __ldg(cell_begin + row + bw) - b;
        }
        u32 incl = len;
#pragma unroll
        for (int o = 1; o < kBrickRows; o <<= 1) {
            const u32 t = __shfl_up_sync(kFull, incl, o);
            if (tid >= o) incl += t;
        }
        if (tid < kBrickRows) {
            s_b[tid] = b;
            s_pre[tid + 1] = incl;
        }
        if (tid == 0) s_pre[0] = 0;
    }
    __syncthreads();
    const u32 total = s_pre[kBrickRows];
    for (u32 base = (u32)(tid & ~31); base < total; base += kBlock) {  // warp-uniform trip count
        const u32 item = base + (tid & 31);
        const bool have = item < total;
        u32 r = 0;
#pragma unroll
        for (int k = 1; k < kBrickRows; k++) r += (have && item >= s_pre[k]) ? 1u : 0u;
        const u32 slot = have ? s_b[r] + (item - s_pre[r]) : 0u;
        f(slot, have);
    }
    (void)n;
#endif
}

// ------------------------------------------------------------------ K6: lambda ------------------------------------------------------------------
template <int RAD>
__global__ void __launch_bounds__(kBlock) k_find_lambdas(float *__restrict__ lambda, u32 *__restrict__ num_neighbors,
                                                         const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                         const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                         const u32 *__restrict__ cell_begin, const float *__restrict__ ros, u32 n,
                                                         u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g, StencilDesc st,
                                                         int zero_nonfluid) {
    extern __shared__ float4 fluid_smem[];
    for_each_slot(g, cell_begin, n, [&](const u32 i, bool act) {
        u32 orig = 0;
        if (act) {
            if (sphase[i] != PH_FLUID) {
                if (zero_nonfluid) lambda[i] = 0.f;
                act = false;
            } else {
                orig = index[i];
            }
        }
        float4 pi = act ? spos[i] : make_float4(g.ox, g.oy, g.oz, 0.f);
        // ghost copies of a neighbour slab's particles: their lambda is needed by the owned particles next to the face
        // (K7 reads lambda_j), and it is exact when the halo holds the ghost's whole neighbourhood, i.e. for ghosts
        // within [ghost_xmin, ghost_xmax]; ghosts further out are never read
        if (act && orig >= n_owned && !(pi.x >= ghost_xmin && pi.x <= ghost_xmax)) act = false;
        if (!__any_sync(kFull, act)) return;
        if (!act) pi = make_float4(g.ox, g.oy, g.oz, 0.f);
        const float ro0 = act ? ros[orig] : 1.f;
        const float inv_ro0 = __fdividef(1.f, ro0);
        const float cs = -PS_SPIKY * inv_ro0;

        float ro = 0.f, denom = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
        const u32 nn = walk_fluid_neighbours<RAD, false>(g, st, cell_begin, spos, act, i, pi, fluid_smem, [&](const float4 e) {
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
    });
}

// ------------------------------------------------------------------ K7: delta p ------------------------------------------------------------------
template <int RAD>
__global__ void __launch_bounds__(kBlock) k_solve_fluids(float4 *__restrict__ pos, const float *__restrict__ lambda,
                                                         const float4 *__restrict__ spos, const int *__restrict__ sphase,
                                                         const u32 *__restrict__ index, const u32 *__restrict__ cell_begin,
                                                         const float *__restrict__ ros, u32 n, u32 n_owned, GridDesc g, StencilDesc st,
                                                         float omega) {
    extern __shared__ float4 fluid_smem[];
    // s_corr = -K_P * (poly6(r) / poly6(dq*H))^4 ; the POLY6 factors cancel (integration_kernel.cuh:630-634)
    const float term2 = PS_H2 - (PS_DQ_P * PS_DQ_P * PS_H2);
    const float inv_den = __fdividef(1.f, term2 * term2 * term2);
    for_each_slot(g, cell_begin, n, [&](const u32 i, bool act) {
        act = act && sphase[i] == PH_FLUID;
        u32 orig = 0;
        if (act) {
            orig = index[i];
            act = orig < n_owned;
        }
        if (!__any_sync(kFull, act)) return;
        const float4 pi = act ? spos[i] : make_float4(g.ox, g.oy, g.oz, 0.f);
        const float li = act ? lambda[i] : 0.f;

        float dx = 0.f, dy = 0.f, dz = 0.f;
        const u32 nn = walk_fluid_neighbours<RAD, true>(g, st, cell_begin, spos, act, i, pi, fluid_smem, [&](const float4 e) {
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
    });
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
static inline dim3 fluid_grid(u32 n, const GridDesc &g) {
#if PS_FMAP == 0
    (void)g;
    return dim3(cdiv(n, kBlock));
#else
    (void)n;  // one CTA per brick of cells; bricks without particles (most of a sparse grid) exit after two loads
    return dim3(cdiv(g.gx, kBrickX), g.gy / kBrickY, g.gz / kBrickZ);
#endif
}

void ps_launch_collide(float4 *pos, const float4 *prev, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                       const u32 *cell_begin, u32 *num_neighbors, u32 n, u32 n_owned, GridDesc g, float radius, cudaStream_t s) {
    if (!n) return;
    StencilDesc st;  // 3x3x3: every row keeps its full extent (contact radius 2.001r slightly exceeds one cell)
    st.rad = 1;
    for (int k = 0; k < 9; k++) st.xr[k] = 1;
    k_collide<<<cdiv(n, kBlock), kBlock, 0, s>>>(pos, prev, spos, sw, sphase, index, cell_begin, num_neighbors, n, n_owned, g, st, radius);
}

void ps_launch_find_lambdas(float *lambda, u32 *num_neighbors, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                            const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g,
                            const StencilDesc &st, bool zero_nonfluid, cudaStream_t s) {
    if (!n) return;
    const size_t sm = fluid_smem_bytes(st.rad);
    static const cudaError_t optin = cudaFuncSetAttribute(k_find_lambdas<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fluid_smem_bytes(PS_MAX_RAD));
    (void)optin;
#ifdef PS_CARVEOUT
    static const cudaError_t carve = cudaFuncSetAttribute(k_find_lambdas<4>, cudaFuncAttributePreferredSharedMemoryCarveout, PS_CARVEOUT);
    (void)carve;
#endif
    if (st.rad == 4)  // the reference's configuration (H = 2, cell = 2r = 0.5): stencil loops fully unrolled
        k_find_lambdas<4><<<fluid_grid(n, g), kBlock, sm, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned, ghost_xmin, ghost_xmax, g, st,
                                                             zero_nonfluid ? 1 : 0);
    else
        k_find_lambdas<0><<<fluid_grid(n, g), kBlock, sm, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned, ghost_xmin, ghost_xmax, g, st,
                                                             zero_nonfluid ? 1 : 0);
}

void ps_launch_solve_fluids(float4 *pos, const float *lambda, const float4 *spos, const int *sphase, const u32 *index,
                            const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, GridDesc g, const StencilDesc &st, float omega,
                            cudaStream_t s) {
    if (!n) return;
    const size_t sm = fluid_smem_bytes(st.rad);
    static const cudaError_t optin = cudaFuncSetAttribute(k_solve_fluids<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fluid_smem_bytes(PS_MAX_RAD));
    (void)optin;
#ifdef PS_CARVEOUT
    static const cudaError_t carve = cudaFuncSetAttribute(k_solve_fluids<4>, cudaFuncAttributePreferredSharedMemoryCarveout, PS_CARVEOUT);
    (void)carve;
#endif
    if (st.rad == 4)
        k_solve_fluids<4><<<fluid_grid(n, g), kBlock, sm, s>>>(pos, lambda, spos, sphase, index, cell_begin, ros, n, n_owned, g, st, omega);
    else
        k_solve_fluids<0><<<fluid_grid(n, g), kBlock, sm, s>>>(pos, lambda, spos, sphase, index, cell_begin, ros, n, n_owned, g, st, omega);
}
