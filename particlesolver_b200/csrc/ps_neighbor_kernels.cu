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
#include <algorithm>
#include <cstdlib>
#include <cuda/std/type_traits>
#include "ps_common.cuh"
#include "ps_fluid_lists.cuh"

namespace {
#ifndef PS_FBLOCK
#define PS_FBLOCK 128
#endif
#ifndef PS_KQ
#define PS_KQ 12  // a multiple of 4 (K7 reads the lists four rows at a time)
#endif
#ifndef PS_FLUSH_PREFETCH
#define PS_FLUSH_PREFETCH 0
#endif
#ifndef PS_VOTE_PAIR
#define PS_VOTE_PAIR 1
#endif

typedef unsigned long long u64;
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
// The neighbour search of the PBF passes.  One thread per sorted slot, one warp = 32 consecutive slots (= a run of
// x-adjacent particles of one or two grid rows).  What keeps the warp's issue slots busy:
//   (1) per-particle row pruning: for stencil row (dy,dz) the x-extent that can hold a neighbour follows from the
//       particle's own position inside its cell, ext = sqrt(H^2 - dymin^2 - dzmin^2) — ~200 candidates per
//       particle instead of the ~370 of the full 9^3 stencil.  The pruning is conservative (eps margin), so the
//       accepted neighbour sequence — and with it the 500-cap and the summation order — is exactly the reference's;
//   (2) per-lane segment lists: the row ranges are short (2.7 candidates on average) and their lengths differ from
//       lane to lane, so walking them row by row in lock-step leaves 45 % of the lanes idle (profiles/r1b).  Instead
//       each lane first writes the non-empty ranges of one z-slab of the stencil (<= 9 rows) to a small list in
//       shared memory, then all lanes walk their own lists in one flat loop whose trip count is the warp's largest
//       candidate total, with a 2-deep software pipeline on the position gather;
//   (3) accept/interact split: the distance test runs over all candidates, accepted neighbours (~65 % of them) are
//       staged in a per-thread shared-memory queue of kQ slots and the expensive interaction body runs over full
//       queues with (nearly) every lane active, instead of under a divergent branch;
// All loops that contain a vote are warp-uniform.  Shared memory is carved out of the same 256 KB as L1, and the
// candidate gather lives on L1 hits, so the per-thread footprint is kept small: kQ 4-byte queue slots (the neighbour's
// sorted slot; its position is re-read, an L1 hit) + 2*(2*rad+1) list entries of (u32 begin, u16 length).
//
// The neighbour lists between K6 and K7 are written by the staged K6 (ps_fluid_staged.cu; format in ps_fluid_lists.cuh).  This walk
// is the general path: every pass when no lists are kept (neighbor_list_rows = 0), and the later passes of the warps without a
// list (kListOverflow: only when neighbor_list_rows is below the 500-neighbour cap).
constexpr int kQ = PS_KQ;
constexpr unsigned kFull = 0xffffffffu;
typedef unsigned short u16;

static inline size_t fluid_smem_bytes(int rad) {
    return (size_t)kQ * kBlock * sizeof(u32) + (size_t)2 * (2 * rad + 1) * kBlock * (sizeof(u32) + sizeof(u16));
}

// body(rx, ry, rz, j) is called for every accepted neighbour, in the reference's traversal order.
template <int RAD, class Body>
__device__ __forceinline__ u32 walk_fluid_neighbours(const GridDesc &g, const StencilDesc &st, const u32 *__restrict__ cell_begin,
                                                     const float4 *__restrict__ spos, bool act, u32 i, float4 pi, u32 *smem, Body &&body) {
    const int tid = threadIdx.x;
    const int rad = RAD ? RAD : st.rad;
    const int nseg = 2 * (2 * rad + 1);  // list capacity per slab: every row may wrap into two ranges
    u32(*q)[kBlock] = reinterpret_cast<u32(*)[kBlock]>(smem);
    u32 *seg_b = smem + kQ * kBlock + tid;  // entry s of this lane at seg_b[s * kBlock]
    u16 *seg_len = reinterpret_cast<u16 *>(smem + kQ * kBlock + nseg * kBlock) + tid;
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
#if PS_FLUSH_PREFETCH
        // positions are re-read (L1 hits: the lines were gathered a few instructions ago) PS_FLUSH_PREFETCH at a time
        // before the interactions that use them, so that each lane has that many loads in flight
#pragma unroll
        for (int k0 = 0; k0 < kQ; k0 += PS_FLUSH_PREFETCH) {
            u32 jj[PS_FLUSH_PREFETCH];
            float4 pp[PS_FLUSH_PREFETCH];
#pragma unroll
            for (int k = 0; k < PS_FLUSH_PREFETCH; k++) {
                jj[k] = (k0 + k < cnt) ? q[k0 + k][tid] : i;
                pp[k] = __ldg(spos + jj[k]);
            }
#pragma unroll
            for (int k = 0; k < PS_FLUSH_PREFETCH; k++)
                if (k0 + k < cnt) body(pi.x - pp[k].x, pi.y - pp[k].y, pi.z - pp[k].z, jj[k]);
        }
#else
#pragma unroll
        for (int k = 0; k < kQ; k++)
            if (k < cnt) {
                const u32 j = q[k][tid];
                const float4 pj = __ldg(spos + j);  // an L1 hit: the line was gathered a few instructions ago
                body(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z, j);
            }
#endif
        nn += cnt;  // accepted neighbours are counted when they leave the queue
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
        const bool capped = __any_sync(kFull, nn + (u32)cnt + total > PS_MAX_NEIGHBORS);  // the 500-neighbour cap can bite in this slab
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
        // A lane that has run out of candidates keeps re-testing its own slot, which is never accepted (j != i): the loop
        // body then needs no per-lane guard, only the refill is predicated.
        u32 j0 = i, j1 = i;
        float4 p0 = pi, p1 = pi;
        if (total > 0) { j0 = next_j(); p0 = __ldg(spos + j0); }
        if (total > 1) { j1 = next_j(); p1 = __ldg(spos + j1); }
        auto walk = [&](auto capped_c) {
            constexpr bool kCapped = decltype(capped_c)::value;
            auto test = [&](const u32 j, const float4 pj) {
                const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                const float r2 = rx * rx + ry * ry + rz * rz;
                if (r2 < PS_H2 && j != i && (!kCapped || nn + cnt < PS_MAX_NEIGHBORS)) q[cnt++][tid] = j;
            };
#pragma unroll 1
            for (u32 t = 0; t < maxtotal; t += 2) {
                {
                    const u32 j = j0;
                    const float4 pj = p0;
                    if (t + 2 < total) { j0 = next_j(); p0 = __ldg(spos + j0); } else { j0 = i; }
                    test(j, pj);
                }
#if !PS_VOTE_PAIR
                if (__any_sync(kFull, cnt == kQ)) flush();
#endif
                {
                    const u32 j = j1;
                    const float4 pj = p1;
                    if (t + 3 < total) { j1 = next_j(); p1 = __ldg(spos + j1); } else { j1 = i; }
                    test(j, pj);
                }
#if PS_VOTE_PAIR
                if (__any_sync(kFull, cnt >= kQ - 1)) flush();  // one vote per two candidates: room for both of the next pair
#else
                if (__any_sync(kFull, cnt == kQ)) flush();
#endif
            }
        };
        if (capped) walk(cuda::std::true_type{});
        else walk(cuda::std::false_type{});
    }
    if (__any_sync(kFull, cnt > 0)) flush();
    return nn;
}

// ------------------------------------------------------------------ K6: lambda, grid-walking form ------------------------------------------------------------------
// The lambda pass when no neighbour lists are kept; with lists the staged K6 (ps_fluid_staged.cu) is the whole pass.
template <int RAD>
// (forcing more resident CTAs through a register cap — 56, 48, 40 registers — makes the walk slower: 0.92 / 0.96 / 1.05 ms against
// 0.78 ms at the compiler's 64; more warps only thrash L1, profiles/r1l)
__global__ void __launch_bounds__(kBlock) k_find_lambdas(float *__restrict__ lambda, u32 *__restrict__ num_neighbors,
                                                         const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                         const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                         const u32 *__restrict__ cell_begin, const float *__restrict__ ros, u32 n,
                                                         u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g, StencilDesc st,
                                                         int zero_nonfluid, u32 *__restrict__ pool, u32 *__restrict__ recs, u32 list_rows) {
    extern __shared__ u32 fluid_smem[];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    const u32 warp = i >> 5;
    const int lane = threadIdx.x & 31;
    // per-warp list status (ps_fluid_lists.cuh); pool == nullptr: no lists are kept
    u32 *rec = pool && (u64)warp * 32 < n ? recs + (size_t)warp * kListRecord : nullptr;
    bool act = i < n;
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
    const bool ghost = act && orig >= n_owned;
    if (ghost && !(pi.x >= ghost_xmin && pi.x <= ghost_xmax)) act = false;
    if (!__any_sync(kFull, act)) {
        if (rec && lane == 0) *rec = 0;
        return;
    }
    if (!act) pi = make_float4(g.ox, g.oy, g.oz, 0.f);
    const float ro0 = act ? ros[orig] : 1.f;
    const float inv_ro0 = __fdividef(1.f, ro0);
    const float cs = -PS_SPIKY * inv_ro0;

    float ro = 0.f, denom = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    // the lane's column of its warp's list region: entry k = the k-th accepted neighbour, appended as the queue is flushed
    u32 *col = pool ? pool + (size_t)warp * list_rows * 32 + lane : nullptr;
    u32 wr = 0;
    bool lost = false;  // an accepted neighbour found no room in the column (only when neighbor_list_rows < 500)
    const u32 nn = walk_fluid_neighbours<RAD>(g, st, cell_begin, spos, act, i, pi, fluid_smem, [&](float rx, float ry, float rz, u32 j) {
        ps_lambda_terms(rx, ry, rz, cs, ro, gx, gy, gz, denom);
        if (col) {
            if (wr < list_rows) __stcs(col + (size_t)wr * 32, j);
            else lost = true;
            wr++;
        }
    });
    if (rec) {
        const bool ovf = __any_sync(kFull, lost);
        if (lane == 0) *rec = ovf ? kListOverflow : 1u;
    }
    if (!act) return;
    const float inv_w = __fdividef(1.f, sw[i]);
    lambda[i] = ps_lambda_from_sums(ro, denom, gx, gy, gz, inv_w, inv_ro0);
    num_neighbors[i] = nn;
}

// ------------------------------------------------------------------ K6: lambda, fused walk (the default with lists) ------------------------------------------------------------------
// The grid walk with the distance test, the interaction and the list append fused into ONE predicated body per candidate — no accept
// queue, no flush, no votes inside the walk.  Measured on the queue form above (ncu source counters, profiles/r2e): 27 SASS instructions
// per candidate for the test + queue, then 48 per queue slot for the flush (guard, slot and position re-read, interaction, list
// store) = ~15.9 k warp-instructions per warp; the fused body is ~38 per candidate = ~10.8 k.  Running the interaction for rejected
// candidates as well (30 % of them, predicated off) is cheaper than separating them.
//   * phase 1 is the walk's: per z-slab every lane writes its non-empty row ranges [begin, end) to a list in shared memory;
//   * phase 2: flat loop, trip count = the warp's largest candidate total of the slab; a candidate's float4 is requested two trips
//     ahead (LDG.128 through L1, predicated on the lane still having candidates) and carries its own sorted slot in .w
//     (ps_launch_reorder(..., slot_in_w)), so the walk tracks nothing but a slot counter and the end of the current range;
//   * an accepted neighbour's slot goes straight to the lane's column of the warp's list region (ps_fluid_lists.cuh).
// Arithmetic: PS_K6_BODY is, association for association, ps_lambda_terms (ps_fluid_lists.cuh) — every K6 variant gives the same bits.
#ifndef PS_K6_LD_HINT
#define PS_K6_LD_HINT ""  // cache hint of the candidate gathers (".L1::evict_last" measured: see DESIGN section 9)
#endif
#ifndef PS_ROW_BATCH
#define PS_ROW_BATCH 9  // stencil rows whose cell-table loads are issued together (fused K6)
#endif
#ifndef PS_FUSED_MINB
#define PS_FUSED_MINB 6  // resident CTAs the register allocation aims at: with the batched row set-up 6 (80 registers) measured 0-1.5 % faster than 7 (72), profiles/r2r
#endif
static inline size_t fused_smem_bytes(int rad) { return (size_t)(2 * (2 * rad + 1) + 1) * kBlock * sizeof(uint2); }

#define PS_K6_FUSED_VISIT(X, Y, Z, W, T, TNEXT, EXTRA_PRED)                                                                      \
    asm volatile("{\n\t"                                                                                                          \
                 ".reg .pred p, q, e, lv;\n\t"                                                                                    \
                 ".reg .f32 rx, ry, rz, r2, ir, rl, h2, hh, hm, a, c, cc;\n\t"                                                     \
                 ".reg .b32 jc;\n\t"                                                                                              \
                 ".reg .b64 wa, la;\n\t"                                                                                          \
                 "sub.ftz.f32 rx, %13, %8;\n\t"                                                                                   \
                 "sub.ftz.f32 ry, %14, %9;\n\t"                                                                                   \
                 "sub.ftz.f32 rz, %15, %10;\n\t"                                                                                  \
                 "mov.b32 jc, %11;\n\t"                                                                                           \
                 "setp.lt.u32 lv, %19, %20;\n\t"                                                                                  \
                 "mul.wide.u32 la, %5, 16;\n\t"                                                                                   \
                 "add.u64 la, la, %24;\n\t"                                                                                       \
                 "@lv ld.global.nc" PS_K6_LD_HINT ".v4.b32 {%8, %9, %10, %11}, [la];\n\t"                                                          \
                 "@lv add.u32 %5, %5, 1;\n\t"                                                                                     \
                 "setp.eq.and.u32 e, %5, %6, lv;\n\t"                                                                             \
                 "@e ld.shared.v2.u32 {%5, %6}, [%7];\n\t"                                                                        \
                 "@e add.u32 %7, %7, %21;\n\t"                                                                                    \
                 "mul.ftz.f32 r2, ry, ry;\n\t"                                                                                    \
                 "fma.rn.ftz.f32 r2, rx, rx, r2;\n\t"                                                                             \
                 "fma.rn.ftz.f32 r2, rz, rz, r2;\n\t"                                                                             \
                 "setp.lt.ftz.f32 p, r2, 0f40800000;\n\t"                                                                         \
                 "setp.lt.and.u32 p, %18, %20, p;\n\t"                                                                            \
                 "setp.ne.and.u32 p, jc, %16, p;\n\t" EXTRA_PRED                                                                  \
                 "rsqrt.approx.ftz.f32 ir, r2;\n\t"                                                                               \
                 "mul.ftz.f32 rl, r2, ir;\n\t"                                                                                    \
                 "sub.ftz.f32 h2, 0f40800000, r2;\n\t"                                                                            \
                 "mul.ftz.f32 hh, h2, h2;\n\t"                                                                                    \
                 "@p fma.rn.ftz.f32 %0, h2, hh, %0;\n\t"                                                                          \
                 "sub.ftz.f32 hm, 0f40000000, rl;\n\t"                                                                            \
                 "setp.ge.and.ftz.f32 q, rl, 0f38D1B717, p;\n\t"                                                                  \
                 "mul.ftz.f32 a, hm, %17;\n\t"                                                                                    \
                 "mul.ftz.f32 a, hm, a;\n\t"                                                                                      \
                 "mul.ftz.f32 c, ir, a;\n\t"                                                                                      \
                 "@q fma.rn.ftz.f32 %1, rx, c, %1;\n\t"                                                                           \
                 "@q fma.rn.ftz.f32 %2, ry, c, %2;\n\t"                                                                           \
                 "@q fma.rn.ftz.f32 %3, rz, c, %3;\n\t"                                                                           \
                 "mul.ftz.f32 cc, c, c;\n\t"                                                                                      \
                 "@q fma.rn.ftz.f32 %4, r2, cc, %4;\n\t"                                                                          \
                 "cvt.u64.u32 wa, %12;\n\t"                                                                                       \
                 "add.u64 wa, wa, %23;\n\t"                                                                                       \
                 "@p st.global.cs.u32 [wa], jc;\n\t"                                                                              \
                 "@p add.u32 %12, %12, 128;\n\t"                                                                                  \
                 "}"                                                                                                              \
                 : "+f"(ro), "+f"(gxs), "+f"(gys), "+f"(gzs), "+f"(denom), "+r"(jn), "+r"(jend), "+r"(sp), "+f"(X), "+f"(Y), "+f"(Z), "+r"(W), "+r"(woff) \
                 : "f"(pi.x), "f"(pi.y), "f"(pi.z), "r"(i), "f"(cs), "r"(T), "r"(TNEXT), "r"(total), "r"((u32)(kBlock * sizeof(uint2))), "r"(room),  \
                   "l"(wfirst_p), "l"(spos)                                                                            \
                 : "memory")

template <int RAD>
__global__ void __launch_bounds__(kBlock, PS_FUSED_MINB) k_find_lambdas_fused(float *__restrict__ lambda, u32 *__restrict__ num_neighbors,
                                                               const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                               const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                               const u32 *__restrict__ cell_begin, const float *__restrict__ ros, u32 n,
                                                               u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g, StencilDesc st,
                                                               int zero_nonfluid, u32 *__restrict__ pool, u32 *__restrict__ recs, u32 list_rows,
                                                               size_t dump_offset, LambdaSinks sinks) {
    extern __shared__ __align__(16) uint2 fused_segs[];
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 i = blockIdx.x * kBlock + tid;
    const u32 warp = i >> 5;
    const bool warp_in = (u64)warp * 32 < n;
    u32 *rec = recs + (size_t)warp * kListRecord;
    bool act = i < n;
    u32 orig = 0;
    if (act) {
        if (sphase[i] != PH_FLUID) {
            if (zero_nonfluid) lambda[i] = 0.f;
            act = false;
        } else {
            orig = index[i];
        }
    }
    const float4 origin = make_float4(g.ox, g.oy, g.oz, 0.f);
    float4 pi = act ? spos[i] : origin;
    // ghost copies of a neighbour slab's particles: lambda only inside [ghost_xmin, ghost_xmax] (see k_find_lambdas)
    if (act && orig >= n_owned && !(pi.x >= ghost_xmin && pi.x <= ghost_xmax)) act = false;
    if (!__any_sync(kFull, act)) {
        if (lane == 0 && warp_in) rec[0] = 0;
        return;
    }
    if (!act) pi = origin;
    const float ro0 = act ? ros[orig] : 1.f;
    const float inv_ro0 = __fdividef(1.f, ro0);
    const float cs = -PS_SPIKY * inv_ro0;

    const int rad = RAD ? RAD : st.rad;
    const int nseg = 2 * (2 * rad + 1);  // list capacity per slab: every row may wrap into two ranges (+ one spare entry the refill may read)
    uint2 *segs = fused_segs + tid;      // entry s of this lane at segs[s * kBlock]
    const u32 seg_addr = (u32)__cvta_generic_to_shared(segs);
    const u32 seg_stride = (u32)(kBlock * sizeof(uint2));
    const float relx = pi.x - g.ox, rely = pi.y - g.oy, relz = pi.z - g.oz;
    const int3 gp = ps_grid_pos(g, pi.x, pi.y, pi.z);
    // margin: covers the approximate divide of the cell assignment and coordinate rounding (ulp(1000) = 6e-5)
    const float eps = 1e-3f + 2e-6f * fmaxf(fabsf(relx), fmaxf(fabsf(rely), fabsf(relz)));
    const float inv_cx = __fdividef(1.f, g.cx);
    const float fy0 = fmaxf(rely - (float)gp.y * g.cy, 0.f), fy1 = fmaxf((float)(gp.y + 1) * g.cy - rely, 0.f);
    const float fz0 = fmaxf(relz - (float)gp.z * g.cz, 0.f), fz1 = fmaxf((float)(gp.z + 1) * g.cz - relz, 0.f);
    auto dmin2 = [&](int d, float f0, float f1, float c) {  // squared distance to the slab of cells at offset d, minus margin
        float m = d == 0 ? 0.f : (d > 0 ? f1 + (float)(d - 1) * c : f0 + (float)(-d - 1) * c);
        m = fmaxf(m - eps, 0.f);
        return m * m;
    };
    float ro = 0.f, denom = 0.f, gxs = 0.f, gys = 0.f, gzs = 0.f;
    // the warp's list region: row k, entry lane = the k-th accepted neighbour of this lane (ps_fluid_lists.cuh)
    u32 *wfirst_p = pool + (size_t)warp * list_rows * 32 + lane;
    u32 woff = 0;  // bytes written to this lane's column: 128 per accepted neighbour
    bool ovf = false;

#pragma unroll 1
    for (int dz = -rad; dz <= rad; dz++) {
        // ---- phase 1: this lane's candidate ranges of the slab ----
        const u32 rowmask = st.rowmask[dz + rad];  // uniform: rows no particle can reach
        const u32 zrow = ((u32)(gp.z + dz) & g.mz) * g.gy;
        const float remz = PS_H2 - dmin2(dz, fz0, fz1, g.cz);
        u32 nlist = 0, total = 0;
        auto push = [&](u32 b, u32 len) {
            if (len) {
                if (nlist >= (u32)nseg) __trap();
                segs[nlist * kBlock] = make_uint2(b, b + len);
                nlist++;
                total += len;
            }
        };
        // Row set-up in two stages per batch of rows: first the address arithmetic and BOTH cell-table loads of every row of the batch
        // (all in flight together), then the list entries.  Done row by row the two dependent loads of a row were 19 % of the
        // kernel's stall samples (profiles/r2i).  A lane's list is its own, so nothing here needs the warp: a row that wraps
        // around the power-of-two grid ([lw, gx) then [0, hw]) gets its second range from the (few) lanes concerned.
        auto row_geometry = [&](int dyi, u32 &row, u32 &lw, u32 &hw) -> bool {
            const int dy = dyi - rad;
            const float rem = remz - dmin2(dy, fy0, fy1, g.cy);
            if (!(act && rem >= 0.f)) return false;
            const float ext = sqrtf(rem) + eps;
            int lo = (int)floorf((relx - ext) * inv_cx), hi = (int)floorf((relx + ext) * inv_cx);
            lo = max(min(lo, gp.x), gp.x - rad);
            hi = min(max(hi, gp.x), gp.x + rad);
            row = (zrow + ((u32)(gp.y + dy) & g.my)) * g.gx;
            lw = (u32)lo & g.mx;
            hw = (u32)hi & g.mx;
            return true;
        };
        auto second_range = [&](int dyi) {  // rare: this lane's window of the row crosses the seam
            u32 row, lw, hw;
            if (row_geometry(dyi, row, lw, hw)) {
                const u32 b1 = __ldg(cell_begin + row);
                push(b1, __ldg(cell_begin + (row + hw + 1u)) - b1);
            }
        };
        if (RAD) {
            constexpr int kRows = 2 * RAD + 1, kBatch = PS_ROW_BATCH;
#pragma unroll
            for (int r0 = 0; r0 < kRows; r0 += kBatch) {
                u32 b0[kBatch], e0[kBatch];
                u32 wraps = 0;
#pragma unroll
                for (int k = 0; k < kBatch; k++) {
                    const int dyi = r0 + k;
                    b0[k] = e0[k] = 0;
                    if (dyi < kRows && ((rowmask >> dyi) & 1u)) {
                        u32 row, lw, hw;
                        if (row_geometry(dyi, row, lw, hw)) {
                            const bool wrap = lw > hw;
                            wraps |= (wrap ? 1u : 0u) << k;
                            b0[k] = __ldg(cell_begin + (row + lw));
                            e0[k] = __ldg(cell_begin + (row + (wrap ? g.mx : hw) + 1u));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < kBatch; k++) {
                    const int dyi = r0 + k;
                    if (dyi < kRows && ((rowmask >> dyi) & 1u)) {
                        push(b0[k], e0[k] - b0[k]);
                        if ((wraps >> k) & 1u) second_range(dyi);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int dyi = 0; dyi <= 2 * rad; dyi++) {
                if (!((rowmask >> dyi) & 1u)) continue;
                u32 row, lw, hw;
                if (!row_geometry(dyi, row, lw, hw)) continue;
                const bool wrap = lw > hw;
                const u32 b0 = __ldg(cell_begin + (row + lw));
                push(b0, __ldg(cell_begin + (row + (wrap ? g.mx : hw) + 1u)) - b0);
                if (wrap) second_range(dyi);
            }
        }
        // ---- phase 2: flat fused walk; the trip count is the warp's largest candidate total ----
        const u32 maxtotal = __reduce_max_sync(kFull, total);
        if (maxtotal == 0) continue;
        // a lane's list would outgrow the warp's region (only when neighbor_list_rows < 500): the warp keeps no list from here on and
        // K7 walks the grid for it; its writes land in the dump region behind the pool
        if (!ovf && __any_sync(kFull, min((woff >> 7) + total, PS_MAX_NEIGHBORS) > list_rows)) {
            ovf = true;
            wfirst_p = pool + dump_offset + lane - (woff >> 2);
        }
        const bool capped = __any_sync(kFull, (woff >> 7) + total > PS_MAX_NEIGHBORS);  // the 500-neighbour cap can bite in this slab
        __syncwarp();  // the lists are written and read by the same lane, but through different address expressions
        // slot of the next candidate to request, end of its range, shared address of the next list entry
        u32 jn = 0, jend = 0, sp = seg_addr;
        if (total) { const uint2 s0 = segs[0]; jn = s0.x; jend = s0.y; sp += seg_stride; }
        float X0 = pi.x, Y0 = pi.y, Z0 = pi.z, X1 = pi.x, Y1 = pi.y, Z1 = pi.z;
        u32 W0 = i, W1 = i;
        auto prime = [&](float &X, float &Y, float &Z, u32 &W, bool live) {
            if (live) {
                const float4 c4 = __ldg(spos + jn);
                X = c4.x; Y = c4.y; Z = c4.z; W = __float_as_uint(c4.w);
                if (++jn == jend) { const uint2 s1 = *reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(fused_segs) + (sp - (u32)__cvta_generic_to_shared(fused_segs))); jn = s1.x; jend = s1.y; sp += seg_stride; }
            }
        };
        prime(X0, Y0, Z0, W0, total > 0);
        prime(X1, Y1, Z1, W1, total > 1);
        if (!capped) {
            const u32 room = 1u;
#pragma unroll 2
            for (u32 t = 0; t < maxtotal; t += 2) {
                const u32 t1 = t + 1, t2 = t + 2, t3 = t + 3;
                PS_K6_FUSED_VISIT(X0, Y0, Z0, W0, t, t2, "");
                PS_K6_FUSED_VISIT(X1, Y1, Z1, W1, t1, t3, "");
            }
        } else {
#pragma unroll 1
            for (u32 t = 0; t < maxtotal; t += 2) {
                const u32 t1 = t + 1, t2 = t + 2, t3 = t + 3;
                u32 room = (woff >> 7) < PS_MAX_NEIGHBORS;
                PS_K6_FUSED_VISIT(X0, Y0, Z0, W0, t, t2, "setp.ne.and.u32 p, %22, 0, p;\n\t");
                room = (woff >> 7) < PS_MAX_NEIGHBORS;
                PS_K6_FUSED_VISIT(X1, Y1, Z1, W1, t1, t3, "setp.ne.and.u32 p, %22, 0, p;\n\t");
            }
        }
    }
    if (lane == 0 && warp_in) rec[0] = ovf ? kListOverflow : 1u;
    if (!act) return;
    const float inv_w = __fdividef(1.f, sw[i]);
    const float lam = ps_lambda_from_sums(ro, denom, gxs, gys, gzs, inv_w, inv_ro0);
    lambda[i] = lam;
    num_neighbors[i] = woff >> 7;
    // slab contexts: a particle of the last halo pack hands its lambda straight to the outgoing messages (what k_slab_pack_lambda would
    // collect in a pass of its own over all sorted slots)
    if (sinks.ranks && orig < n_owned && (pi.x < sinks.left_below || pi.x >= sinks.right_from)) {  // positions are those the pack classified
        const uint2 r = __ldg(sinks.ranks + orig);
        if (r.x < sinks.cap) sinks.left[r.x] = lam;
        if (r.y < sinks.cap) sinks.right[r.y] = lam;
    }
}

// ------------------------------------------------------------------ K7: delta p ------------------------------------------------------------------
// s_corr = -K_P * (poly6(r) / poly6(dq*H))^4 ; the POLY6 factors cancel (integration_kernel.cuh:630-634)
struct DeltaP {
    float li, inv_den, dx, dy, dz;
    const float *__restrict__ lambda;
    // explicit roundings (no compiler-chosen contraction): the list reader and the grid walk instantiate this in different kernels
    // and must produce the same bits (tests/test_gpu_parity.py::test_neighbour_list_paths_agree_bit_for_bit)
    __device__ __forceinline__ void operator()(float rx, float ry, float rz, u32 j) { apply(rx, ry, rz, __ldg(lambda + j)); }
    __device__ __forceinline__ void apply(float rx, float ry, float rz, float lj) {
        const float r2 = __fmaf_rn(rz, rz, __fmaf_rn(rx, rx, __fmul_rn(ry, ry)));
        const float inv_r = rsqrtf(r2);
        const float rlen = __fmul_rn(r2, inv_r);
        const float hm2 = __fsub_rn(PS_H2, r2);
        const float qq = __fmul_rn(__fmul_rn(__fmul_rn(hm2, hm2), hm2), inv_den);
        const float q2 = __fmul_rn(qq, qq);
        const float s = __fmaf_rn(-PS_K_P, __fmul_rn(q2, q2), __fadd_rn(li, lj));
        if (rlen >= 0.0001f) {
            const float hm = __fsub_rn(PS_H, rlen);
            const float c = __fmul_rn(s, __fmul_rn(__fmul_rn(__fmul_rn(-PS_SPIKY, hm), hm), inv_r));
            dx = __fmaf_rn(rx, c, dx); dy = __fmaf_rn(ry, c, dy); dz = __fmaf_rn(rz, c, dz);
        } else {  // coincident: the reference nudges along +y, (0,EPS,0,0) * -SPIKY * (H-r)^2 (:625-626)
            const float rl = (r2 > 0.f) ? rlen : 0.f;
            const float hm = __fsub_rn(PS_H, rl);
            dy = __fmaf_rn(s, __fmul_rn(__fmul_rn(PS_EPS * -PS_SPIKY, hm), hm), dy);
        }
    }
};
__device__ __forceinline__ float delta_p_inv_den() {
    const float term2 = PS_H2 - (PS_DQ_P * PS_DQ_P * PS_H2);
    return __fdividef(1.f, term2 * term2 * term2);
}

// K7 from the neighbour lists K6 left behind: no search, no shared memory (reader: ps_for_each_listed, ps_fluid_lists.cuh).
#ifndef PS_K7_BLOCK
#define PS_K7_BLOCK 256
#endif
constexpr int kListBlock = PS_K7_BLOCK;
__global__ void __launch_bounds__(kListBlock) k_solve_fluids_list(float4 *__restrict__ pos, const float *__restrict__ lambda,
                                                                  const float4 *__restrict__ spos, const int *__restrict__ sphase,
                                                                  const u32 *__restrict__ index, const float *__restrict__ ros,
                                                                  const u32 *__restrict__ nbr_list, const u32 *__restrict__ nbr_rows, u32 list_rows,
                                                                  const u32 *__restrict__ num_neighbors, u32 n, u32 n_owned, float omega) {
    const u32 i = blockIdx.x * kListBlock + threadIdx.x;
    if (i >= n) return;
    const u32 warp = i >> 5;
    const u32 status = nbr_rows[(size_t)warp * kListRecord];
    if (status == kListOverflow || status == 0) return;  // warps without a list are redone by k_solve_fluids
    if (sphase[i] != PH_FLUID) return;
    const u32 orig = index[i];
    if (orig >= n_owned) return;
    const float4 pi = spos[i];
    const u32 nn = num_neighbors[i];
    DeltaP f{lambda[i], delta_p_inv_den(), 0.f, 0.f, 0.f, lambda};
    ps_for_each_listed_lambda(ps_list_column(nbr_list, warp, list_rows, threadIdx.x & 31), nn, i, pi, spos, lambda, f);
    const float inv_div = __fdividef(omega, __fadd_rn(ros[orig], (float)nn));
    float4 P = pos[orig];
    P.x = __fmaf_rn(f.dx, inv_div, P.x); P.y = __fmaf_rn(f.dy, inv_div, P.y); P.z = __fmaf_rn(f.dz, inv_div, P.z);
    pos[orig] = P;
}

// K7 by walking the grid again: the whole pass when no lists are kept, else only the warps whose list overflowed
template <int RAD>
__global__ void __launch_bounds__(kBlock) k_solve_fluids(float4 *__restrict__ pos, const float *__restrict__ lambda,
                                                         const float4 *__restrict__ spos, const int *__restrict__ sphase,
                                                         const u32 *__restrict__ index, const u32 *__restrict__ cell_begin,
                                                         const float *__restrict__ ros, u32 n, u32 n_owned, GridDesc g, StencilDesc st,
                                                         float omega, const u32 *__restrict__ nbr_rows) {
    extern __shared__ u32 fluid_smem[];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (nbr_rows && ((u64)(i >> 5) * 32 >= n || nbr_rows[(size_t)(i >> 5) * kListRecord] != kListOverflow)) return;  // warp-uniform
    bool act = i < n && sphase[i] == PH_FLUID;
    u32 orig = 0;
    if (act) {
        orig = index[i];
        act = orig < n_owned;
    }
    if (!__any_sync(kFull, act)) return;
    const float4 pi = act ? spos[i] : make_float4(g.ox, g.oy, g.oz, 0.f);
    DeltaP f{act ? lambda[i] : 0.f, delta_p_inv_den(), 0.f, 0.f, 0.f, lambda};
    const u32 nn = walk_fluid_neighbours<RAD>(g, st, cell_begin, spos, act, i, pi, fluid_smem, f);
    if (!act) return;
    const float inv_div = __fdividef(omega, __fadd_rn(ros[orig], (float)nn));
    float4 P = pos[orig];
    P.x = __fmaf_rn(f.dx, inv_div, P.x); P.y = __fmaf_rn(f.dy, inv_div, P.y); P.z = __fmaf_rn(f.dz, inv_div, P.z);
    pos[orig] = P;
}

// ------------------------------------------------------------------ K13: XSPH viscosity + vorticity confinement ------------------------------------------------------------------
// Optional, default off, NOT in the reference (neither exists there): the two velocity post-passes of
// Macklin & Mueller 2013 ("Position Based Fluids", eqs. 15-17) on the neighbour structure the PBF passes already built.
//   pass 1  omega_i = sum_j (v_j - v_i) x grad_pj W(p_i - p_j)                       -> omega[i] = (omega, |omega|) by sorted slot
//   pass 2  dv_i = c sum_j (v_j - v_i) W_poly6(p_i - p_j)  +  dt eps (N x omega_i),   N = eta / |eta|, eta = sum_j |omega_j| grad_pi W
//   pass 3  v_i += dv_i   (separate, so that pass 2 reads only old velocities)
// Positions are the sorted copies of the last grid build; W_poly6 / grad W_spiky are the solver's own kernels.  A pass
// walks K6's neighbour lists; warps whose list overflowed (or everything, when no lists are kept) re-walk the grid.
struct OmegaOp {
    const float4 *__restrict__ vel;
    const u32 *__restrict__ index;
    float4 *__restrict__ out;  // omega by sorted slot
    float4 vi;
    float ox, oy, oz;
    __device__ __forceinline__ void begin(u32, u32 orig) { vi = vel[orig]; ox = oy = oz = 0.f; }
    __device__ __forceinline__ void operator()(float rx, float ry, float rz, u32 j) {
        const float r2 = rx * rx + ry * ry + rz * rz;
        const float inv_r = rsqrtf(r2), rlen = r2 * inv_r;
        if (!(rlen >= 0.0001f)) return;
        const float4 vj = __ldg(vel + __ldg(index + j));
        const float hm = PS_H - rlen;
        const float c = (PS_SPIKY * hm * hm) * inv_r;  // grad_pj W(p_i - p_j) = +SPIKY (H - r)^2 r / |r|
        const float gx = rx * c, gy = ry * c, gz = rz * c;
        const float dx = vj.x - vi.x, dy = vj.y - vi.y, dz = vj.z - vi.z;
        ox += dy * gz - dz * gy; oy += dz * gx - dx * gz; oz += dx * gy - dy * gx;
    }
    __device__ __forceinline__ void end(u32 i, u32) { out[i] = make_float4(ox, oy, oz, sqrtf(ox * ox + oy * oy + oz * oz)); }
};
struct ViscosityOp {
    const float4 *__restrict__ vel;
    const u32 *__restrict__ index;
    const float4 *__restrict__ omega;
    float4 *__restrict__ out;  // dv by sorted slot
    float c_xsph, eps_dt;
    float4 vi;
    float sx, sy, sz, ex, ey, ez;
    __device__ __forceinline__ void begin(u32, u32 orig) { vi = vel[orig]; sx = sy = sz = ex = ey = ez = 0.f; }
    __device__ __forceinline__ void operator()(float rx, float ry, float rz, u32 j) {
        const float r2 = rx * rx + ry * ry + rz * rz;
        const float4 vj = __ldg(vel + __ldg(index + j));
        const float hm2 = PS_H2 - r2;
        const float W = PS_POLY6 * (hm2 * hm2 * hm2);
        sx += (vj.x - vi.x) * W; sy += (vj.y - vi.y) * W; sz += (vj.z - vi.z) * W;
        const float inv_r = rsqrtf(r2), rlen = r2 * inv_r;
        if (!(rlen >= 0.0001f)) return;
        const float hm = PS_H - rlen;
        const float c = __ldg(omega + j).w * ((-PS_SPIKY * hm * hm) * inv_r);  // |omega_j| grad_pi W
        ex += rx * c; ey += ry * c; ez += rz * c;
    }
    __device__ __forceinline__ void end(u32 i, u32) {
        float dx = c_xsph * sx, dy = c_xsph * sy, dz = c_xsph * sz;
        const float en = sqrtf(ex * ex + ey * ey + ez * ez);
        if (en > 1e-6f && eps_dt != 0.f) {
            const float s = eps_dt / en;
            const float4 w = omega[i];
            dx += s * (ey * w.z - ez * w.y); dy += s * (ez * w.x - ex * w.z); dz += s * (ex * w.y - ey * w.x);
        }
        out[i] = make_float4(dx, dy, dz, 0.f);
    }
};

// diagnostics: the K6 density estimate of every fluid particle as |rho / rho0 - 1| (the quantity the constraint drives to 0)
struct DensityErrorOp {
    const float *__restrict__ sw;
    const float *__restrict__ ros;
    float4 *__restrict__ out;  // .x by sorted slot
    float ro;
    __device__ __forceinline__ void begin(u32, u32) { ro = 0.f; }
    __device__ __forceinline__ void operator()(float rx, float ry, float rz, u32) {
        const float hm2 = PS_H2 - (rx * rx + ry * ry + rz * rz);
        ro += hm2 * hm2 * hm2;
    }
    __device__ __forceinline__ void end(u32 i, u32 orig) {
        const float rho = (ro + PS_H6) * __fdividef(PS_POLY6, sw[i]);
        out[i] = make_float4(fabsf(__fdividef(rho, ros[orig]) - 1.f), 0.f, 0.f, 0.f);
    }
};

template <class Op>
__global__ void __launch_bounds__(kListBlock) k_fluid_pass_list(Op op, const float4 *__restrict__ spos, const int *__restrict__ sphase,
                                                                const u32 *__restrict__ index, const u32 *__restrict__ nbr_list,
                                                                const u32 *__restrict__ nbr_rows, u32 list_rows,
                                                                const u32 *__restrict__ num_neighbors, u32 n) {
    const u32 i = blockIdx.x * kListBlock + threadIdx.x;
    if (i >= n) return;
    const u32 warp = i >> 5;
    const u32 status = nbr_rows[(size_t)warp * kListRecord];
    if (status == kListOverflow) return;  // redone by k_fluid_pass_walk
    if (sphase[i] != PH_FLUID) return;
    const u32 orig = index[i];
    const float4 pi = spos[i];
    op.begin(i, orig);
    // status 0: K6 found no active fluid particle in the warp (ghosts outside the lambda range): nothing listed
    if (status) ps_for_each_listed(ps_list_column(nbr_list, warp, list_rows, threadIdx.x & 31), num_neighbors[i], i, pi, spos, op);
    op.end(i, orig);
}
template <int RAD, class Op>
__global__ void __launch_bounds__(kBlock) k_fluid_pass_walk(Op op, const float4 *__restrict__ spos, const int *__restrict__ sphase,
                                                            const u32 *__restrict__ index, const u32 *__restrict__ cell_begin, u32 n, GridDesc g,
                                                            StencilDesc st, const u32 *__restrict__ nbr_rows) {
    extern __shared__ u32 fluid_smem[];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (nbr_rows && ((u64)(i >> 5) * 32 >= n || nbr_rows[(size_t)(i >> 5) * kListRecord] != kListOverflow)) return;  // warp-uniform
    const bool act = i < n && sphase[i] == PH_FLUID;
    if (!__any_sync(kFull, act)) return;
    const u32 orig = act ? index[i] : 0u;
    const float4 pi = act ? spos[i] : make_float4(g.ox, g.oy, g.oz, 0.f);
    if (act) op.begin(i, orig);
    walk_fluid_neighbours<RAD>(g, st, cell_begin, spos, act, i, pi, fluid_smem, op);
    if (act) op.end(i, orig);
}
__global__ void __launch_bounds__(256) k_apply_dv(float4 *__restrict__ vel, const float4 *__restrict__ dv, const int *__restrict__ sphase,
                                                  const u32 *__restrict__ index, u32 n) {
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n || sphase[i] != PH_FLUID) return;
    const u32 orig = index[i];
    const float4 d = dv[i];
    float4 v = vel[orig];
    v.x += d.x; v.y += d.y; v.z += d.z;
    vel[orig] = v;
}

// ------------------------------------------------------------------ K5: contacts + friction ------------------------------------------------------------------
// 27 cells = 9 rows of 3.  Two sweeps over the same rows: the first only counts (every per-neighbour term is
// divided by the final neighbour count, and friction is non-linear in it), the second accumulates.
__device__ __forceinline__ float ps_scaled_w(float w, float y) {
    // sW = w != 0 ? 1 / ((1/w) * exp(-y)) : w   (integration_kernel.cuh:403,425)
    return (w != 0.f) ? __fdividef(1.f, __fdividef(1.f, w) * __expf(-y)) : w;
}

// SDF contact of two rigid-body particles: the reference CPU app's RigidContactConstraint (cpu/src/constraint/
// rigidcontactconstraint.cpp:13-66, 2-D) lifted to 3-D.  si / sj = (outward unit gradient in world frame, depth below the body's
// surface) of this particle and its partner, r = x_i - x_j.  The particle that sits shallower in its own body supplies normal and
// depth (ties: the lower particle index, so that both particles of a pair see the same contact).  For a particle of the outermost
// layers (depth < diameter + EPS) the depth is the particles' overlap and the normal is the direction to the partner, mirrored at
// the SDF normal when the partner lies behind the surface (Macklin et al. 2014, eq. 13-14).
// Deviation from the reference's 2-D code, on purpose: it takes x12 = p1 - p2 — pointing from the partner to the particle — so
// its mirror branch is the common case and an oblique contact's sideways component comes out flipped.  Lifted as is, resting
// contacts on a box's edges and corners push sideways and, with the GPU solver's friction constants (0.005 / 0.0002), a tower of
// three boxes walks apart within 300 steps (scripts/diag_rigid_scene.py); with x_ij = (x_j - x_i) / |.| the tower rests.
// Returns false when the pair needs no correction; e = unit normal pointing from i to j.
__device__ __forceinline__ bool sdf_contact(float4 si, float4 sj, bool i_first, float rx, float ry, float rz, float dist, float diam, float &d,
                                            float &ex, float &ey, float &ez) {
    const bool mine = si.w < sj.w || (si.w == sj.w && i_first);
    if (mine) { d = si.w; ex = si.x; ey = si.y; ez = si.z; }
    else { d = sj.w; ex = -sj.x; ey = -sj.y; ez = -sj.z; }
    if (d < diam + PS_EPS) {
        d = diam - dist;
        if (d < PS_EPS) return false;
        float x = 0.f, y = 1.f, z = 0.f;
        if (dist > PS_EPS) { const float inv = __fdividef(-1.f, dist); x = rx * inv; y = ry * inv; z = rz * inv; }
        else if (!i_first) y = -1.f;  // coincident: the lower index is pushed down, the other up
        const float dp = x * ex + y * ey + z * ez;
        if (dp < 0.f) { ex = x - 2.f * dp * ex; ey = y - 2.f * dp * ey; ez = z - 2.f * dp * ez; }
        else { ex = x; ey = y; ez = z; }
    }
    return true;
}

__global__ void __launch_bounds__(kBlock) k_collide(float4 *__restrict__ pos, const float4 *__restrict__ prev,
                                                    const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                    const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                    const u32 *__restrict__ cell_begin, u32 *__restrict__ num_neighbors, u32 n, u32 n_owned,
                                                    GridDesc g, StencilDesc st, float radius, float omega, const u32 *__restrict__ adj_off,
                                                    const u32 *__restrict__ adj, const float4 *__restrict__ sdf_world) {
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

    // Particles of one phase > SOLID never collide with each other in the reference (integration_kernel.cuh:336-337: a cloth
    // does not self-collide).  adj_off != nullptr (PS_FLAG_SELF_COLLISION, not in the reference): such a pair inside
    // the contact distance is skipped only when either particle has no distance constraint (a shape-matched body) or the two
    // are joined by one (CSR adjacency by original index); the lookup runs after the distance test, i.e. for contacts only.
    u32 ab = 0, ae = 0;
    if (adj_off && phase > PH_SOLID) { ab = __ldg(adj_off + orig); ae = __ldg(adj_off + orig + 1); }
    auto same_body = [&](u32 j) {
        if (ab == ae) return true;
        const u32 oj = __ldg(index + j);
        if (__ldg(adj_off + oj) == __ldg(adj_off + oj + 1)) return true;
        for (u32 k = ab; k < ae; k++)
            if (__ldg(adj + k) == oj) return true;
        return false;
    };
    u32 nn = 0;
    for_each_candidate(g, st, cell_begin, gp, [&](u32 j) {
        if (j == i) return;
        const int phase2 = __ldg(sphase + j);
        const bool same = phase > PH_SOLID && phase == phase2;
        if (same && !adj_off) return;
        const float4 pj = __ldg(spos + j);
        const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
        if (!(rx * rx + ry * ry + rz * rz < collide_dist2)) return;
        if (same && same_body(j)) return;
        if (nn < PS_MAX_NEIGHBORS) nn++;
    });
    num_neighbors[i] = nn;
    float dxs = 0.f, dys = 0.f, dzs = 0.f;
    if (nn) {
        const float w = sw[i];
        const float sW = ps_scaled_w(w, pi.y);
        const float4 pp = __ldg(prev + orig);
        // Jacobi averaging over the contacts (integration_kernel.cuh:436-437) with the SOR factor: delta * omega / numNeighbors
        const float fn = (omega == 1.f) ? (float)nn : __fdividef((float)nn, omega);
        // SDF contacts between rigid bodies (ps_set_rigid_body_sdf; not in the reference's GPU solver): w < 0 (or NaN) = no SDF
        const float4 si = sdf_world ? __ldg(sdf_world + orig) : make_float4(0.f, 0.f, 0.f, -1.f);
        u32 seen = 0;
        for_each_candidate(g, st, cell_begin, gp, [&](u32 j) {
            if (j == i || seen >= nn) return;
            const int phase2 = __ldg(sphase + j);
            const bool same = phase > PH_SOLID && phase == phase2;
            if (same && !adj_off) return;
            const float4 pj = __ldg(spos + j);
            const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
            const float d2 = rx * rx + ry * ry + rz * rz;
            if (!(d2 < collide_dist2)) return;
            if (same && same_body(j)) return;
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
            float px = rx * sd, py = ry * sd, pz = rz * sd;  // dp
            float fnx = rx, fny = ry, fnz = rz, fd = dist;    // friction: normal before normalisation, length scale of the cone
            if (both_solid && si.w >= 0.f) {
                const u32 oj = __ldg(index + j);
                const float4 sj = __ldg(sdf_world + oj);
                if (sj.w >= 0.f) {
                    float d, ex, ey, ez;
                    if (!sdf_contact(si, sj, orig < oj, rx, ry, rz, dist, 2.f * radius, d, ex, ey, ez)) return;
                    const float s_ = __fdividef(d, wsum);
                    px = ex * s_; py = ey * s_; pz = ez * s_;
                    fnx = ex; fny = ey; fnz = ez;  // friction about the contact normal; its cone keeps this pass's own scale (dist)
                }
            }
            const float d1x = __fdividef(-colW * px, fn), d1y = __fdividef(-colW * py, fn), d1z = __fdividef(-colW * pz, fn);
            dxs += d1x; dys += d1y; dzs += d1z;
            if (!both_solid) return;
            const float d2x = __fdividef(colW2 * px, fn), d2y = __fdividef(colW2 * py, fn), d2z = __fdividef(colW2 * pz, fn);
            const float4 pp2 = __ldg(prev + __ldg(index + j));
            const float inv = rsqrtf(fnx * fnx + fny * fny + fnz * fnz);  // == rsqrtf(d2) for plain contacts
            const float nx = fnx * inv, ny = fny * inv, nz = fnz * inv;
            // [sic] the reference's second term starts from prevPos of i, not pos2 (integration_kernel.cuh:447)
            const float ex = (pi.x + d1x - pp.x) - (pp.x + d2x - pp2.x);
            const float ey = (pi.y + d1y - pp.y) - (pp.y + d2y - pp2.y);
            const float ez = (pi.z + d1z - pp.z) - (pp.z + d2z - pp2.z);
            const float dn = ex * nx + ey * ny + ez * nz;
            const float tx = ex - dn * nx, ty = ey - dn * ny, tz = ez - dn * nz;
            const float lt = sqrtf(tx * tx + ty * ty + tz * tz);
            if (lt < PS_EPS) return;
            if (lt < PS_S_FRICTION * fd) {
                dxs -= __fdividef(tx * colW, wsum); dys -= __fdividef(ty * colW, wsum); dzs -= __fdividef(tz * colW, wsum);
            } else {
                const float m = fminf(__fdividef(PS_K_FRICTION * fd, lt), 1.f);
                dxs -= tx * m; dys -= ty * m; dzs -= tz * m;
            }
        });
    }
    pos[orig] = make_float4(pi.x + dxs, pi.y + dys, pi.z + dzs, 1.0f);
}
}  // namespace

static inline u32 cdiv(u32 a, u32 b) { return (a + b - 1) / b; }

// The opt-in to > 48 KB of dynamic shared memory (generic stencil radius) is a per-device attribute: once for every device this
// process drives (several contexts on several GPUs may live in one process).
static void ps_optin_smem(int device) {
    static bool opted[64] = {};
    if (device < 0 || device >= 64 || opted[device]) return;
    cudaFuncSetAttribute(k_find_lambdas<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fluid_smem_bytes(PS_MAX_RAD));
    cudaFuncSetAttribute(k_find_lambdas_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes(PS_MAX_RAD));
    cudaFuncSetAttribute(k_solve_fluids<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fluid_smem_bytes(PS_MAX_RAD));
    opted[device] = true;
}

void ps_launch_collide(float4 *pos, const float4 *prev, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                       const u32 *cell_begin, u32 *num_neighbors, u32 n, u32 n_owned, GridDesc g, float radius, float omega, const u32 *adj_off,
                       const u32 *adj, const float4 *sdf_world, cudaStream_t s) {
    if (!n) return;
    StencilDesc st;  // 3x3x3: every row keeps its full extent (contact radius 2.001r slightly exceeds one cell)
    st.rad = 1;
    for (int k = 0; k < 9; k++) st.xr[k] = 1;
    k_collide<<<cdiv(n, kBlock), kBlock, 0, s>>>(pos, prev, spos, sw, sphase, index, cell_begin, num_neighbors, n, n_owned, g, st, radius, omega, adj_off, adj, sdf_world);
}

// list regions for `capacity` particles (rows_per_warp rows of 32 entries per warp) + the staged K6's dump region + the read-ahead
// of the list readers
static size_t list_region_elems(u64 capacity, u32 rows_per_warp) { return (size_t)((capacity + 31) / 32) * rows_per_warp * 32; }
size_t ps_neighbor_list_elems(u64 capacity, u32 rows_per_warp) {
    return list_region_elems(capacity, rows_per_warp) + ((size_t)ps_staged_dump_rows() + PS_LIST_READ_ROWS) * 32;
}
// per-warp status words behind a 4-word header (kept for alignment); the kernels are handed the address of the first word
size_t ps_neighbor_record_elems(u64 capacity) { return 4 + (size_t)((capacity + 31) / 32) * kListRecord; }

// K6.  With lists (nbr_list / nbr_rows / list_rows != 0) the pass also leaves the neighbour lists (ps_fluid_lists.cuh) for K7 and the
// other list readers; else the caller passes nbr_list = nullptr to those as well.  staged: the TMA-staged kernel (ps_fluid_staged.cu;
// needs lists and slot_in_w: spos[j].w = bit pattern of j, ps_launch_reorder(..., slot_in_w = true)) instead of the grid walk.
u32 ps_launch_find_lambdas(float *lambda, u32 *num_neighbors, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                           const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g,
                           const StencilDesc &st, bool zero_nonfluid, u32 *nbr_list, u32 *nbr_rows, u32 list_rows, u64 capacity, bool slot_in_w,
                           bool staged, int device, cudaStream_t s, LambdaSinks sinks, bool *sinks_written) {
    if (sinks_written) *sinks_written = false;
    if (!n) return 0;
    const bool lists = nbr_list && nbr_rows && list_rows && n <= capacity;
    if (lists && staged && slot_in_w) {
        ps_launch_find_lambdas_staged(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned, ghost_xmin, ghost_xmax, g, st,
                                      zero_nonfluid, nbr_list, nbr_rows, list_rows, list_region_elems(capacity, list_rows), device, s);
        return 1;
    }
    ps_optin_smem(device);
    static const bool queue_walk = getenv("PS_K6_QUEUE_WALK") != nullptr;  // tuning aid: the queue form of the walk with lists
    if (lists && slot_in_w && !queue_walk) {  // the default: fused walk
        const size_t fsm = fused_smem_bytes(st.rad);
        const size_t dump = list_region_elems(capacity, list_rows);
        if (st.rad == 4)
            k_find_lambdas_fused<4><<<cdiv(n, kBlock), kBlock, fsm, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned,
                                                                         ghost_xmin, ghost_xmax, g, st, zero_nonfluid ? 1 : 0, nbr_list, nbr_rows, list_rows, dump, sinks);
        else
            k_find_lambdas_fused<0><<<cdiv(n, kBlock), kBlock, fsm, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned,
                                                                         ghost_xmin, ghost_xmax, g, st, zero_nonfluid ? 1 : 0, nbr_list, nbr_rows, list_rows, dump, sinks);
        if (sinks_written) *sinks_written = sinks.ranks != nullptr;
        return 1;
    }
    const size_t sm = fluid_smem_bytes(st.rad);
    u32 *pool = lists ? nbr_list : nullptr;
    if (st.rad == 4)  // the reference's configuration (H = 2, cell = 2r = 0.5): stencil loops fully unrolled
        k_find_lambdas<4><<<cdiv(n, kBlock), kBlock, sm, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned, ghost_xmin,
                                                             ghost_xmax, g, st, zero_nonfluid ? 1 : 0, pool, nbr_rows, list_rows);
    else
        k_find_lambdas<0><<<cdiv(n, kBlock), kBlock, sm, s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned, ghost_xmin,
                                                             ghost_xmax, g, st, zero_nonfluid ? 1 : 0, pool, nbr_rows, list_rows);
    return 1;
}

// nbr_list != nullptr: K7 from the lists K6 wrote, then the grid walk for the warps without a list; else the grid walk for all
u32 ps_launch_solve_fluids(float4 *pos, const float *lambda, const float4 *spos, const int *sphase, const u32 *index,
                           const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, GridDesc g, const StencilDesc &st, float omega,
                           const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows, const u32 *num_neighbors, int device, cudaStream_t s) {
    if (!n) return 0;
    const size_t sm = fluid_smem_bytes(st.rad);
    ps_optin_smem(device);
    u32 launches = 1;
    if (nbr_list && nbr_rows && list_rows) {
        k_solve_fluids_list<<<cdiv(n, kListBlock), kListBlock, 0, s>>>(pos, lambda, spos, sphase, index, ros, nbr_list, nbr_rows, list_rows, num_neighbors, n,
                                                                       n_owned, omega);
        if (list_rows >= PS_MAX_NEIGHBORS) return 1;  // every list fits its region: no warp is left for the grid walk
        launches++;
    } else {
        nbr_rows = nullptr;
    }
    if (st.rad == 4)
        k_solve_fluids<4><<<cdiv(n, kBlock), kBlock, sm, s>>>(pos, lambda, spos, sphase, index, cell_begin, ros, n, n_owned, g, st, omega, nbr_rows);
    else
        k_solve_fluids<0><<<cdiv(n, kBlock), kBlock, sm, s>>>(pos, lambda, spos, sphase, index, cell_begin, ros, n, n_owned, g, st, omega, nbr_rows);
    return launches;
}

template <class Op>
static u32 launch_fluid_pass(Op op, const float4 *spos, const int *sphase, const u32 *index, const u32 *cell_begin, u32 n, GridDesc g,
                             const StencilDesc &st, const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows, const u32 *num_neighbors, int device,
                             cudaStream_t s) {
    const size_t sm = fluid_smem_bytes(st.rad);
    static bool opted[64] = {};  // per device: the opt-in to > 48 KB of dynamic shared memory for the generic-radius walk
    if (device >= 0 && device < 64 && !opted[device]) {
        cudaFuncSetAttribute(k_fluid_pass_walk<0, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fluid_smem_bytes(PS_MAX_RAD));
        opted[device] = true;
    }
    u32 launches = 1;
    if (nbr_list && nbr_rows && list_rows) {
        k_fluid_pass_list<Op><<<cdiv(n, kListBlock), kListBlock, 0, s>>>(op, spos, sphase, index, nbr_list, nbr_rows, list_rows, num_neighbors, n);
        if (list_rows >= PS_MAX_NEIGHBORS) return 1;
        launches++;
    } else {
        nbr_rows = nullptr;
    }
    if (st.rad == 4) k_fluid_pass_walk<4, Op><<<cdiv(n, kBlock), kBlock, sm, s>>>(op, spos, sphase, index, cell_begin, n, g, st, nbr_rows);
    else k_fluid_pass_walk<0, Op><<<cdiv(n, kBlock), kBlock, sm, s>>>(op, spos, sphase, index, cell_begin, n, g, st, nbr_rows);
    return launches;
}

// XSPH viscosity + vorticity confinement (K13): vel += dv, see the kernels.  scratch: float4[2n] (omega | dv by sorted slot).
u32 ps_launch_viscosity(float4 *vel, float4 *scratch, const float4 *spos, const int *sphase, const u32 *index, const u32 *cell_begin, u32 n, GridDesc g,
                        const StencilDesc &st, float c_xsph, float vorticity_eps, float dt, const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows,
                        const u32 *num_neighbors, int device, cudaStream_t s) {
    if (!n) return 0;
    float4 *omega = scratch, *dv = scratch + n;
    u32 launches = 0;
    if (vorticity_eps != 0.f) {
        OmegaOp o1{vel, index, omega};
        launches += launch_fluid_pass(o1, spos, sphase, index, cell_begin, n, g, st, nbr_list, nbr_rows, list_rows, num_neighbors, device, s);
    } else {
        cudaMemsetAsync(omega, 0, (size_t)n * sizeof(float4), s);
    }
    ViscosityOp o2{vel, index, omega, dv, c_xsph, vorticity_eps * dt};
    launches += launch_fluid_pass(o2, spos, sphase, index, cell_begin, n, g, st, nbr_list, nbr_rows, list_rows, num_neighbors, device, s);
    k_apply_dv<<<cdiv(n, 256), 256, 0, s>>>(vel, dv, sphase, index, n);
    return launches + 1;
}

// |rho / rho0 - 1| per fluid particle into scratch[i].x (sorted slot), on the neighbour structure K6 just built
u32 ps_launch_density_error(float4 *scratch, const float4 *spos, const float *sw, const int *sphase, const u32 *index, const float *ros, const u32 *cell_begin,
                            u32 n, GridDesc g, const StencilDesc &st, const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows, const u32 *num_neighbors,
                            int device, cudaStream_t s) {
    if (!n) return 0;
    DensityErrorOp op{sw, ros, scratch};
    return launch_fluid_pass(op, spos, sphase, index, cell_begin, n, g, st, nbr_list, nbr_rows, list_rows, num_neighbors, device, s);
}
