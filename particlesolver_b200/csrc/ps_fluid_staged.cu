// particlesolver_b200/csrc/ps_fluid_staged.cu — K6, PBF lambda (reference findLambdasD / collideCellRadius,
// gpu/src/cuda/integration_kernel.cuh:480-593), with the neighbour rows staged in shared memory by TMA bulk copies.
//
// Kept from the reference: one thread per SORTED slot; the neighbour set (r^2 < H^2, j != i), its traversal order (cells z, y, x
// outermost to innermost, ascending sorted slot inside a cell), the 500-neighbour cap and the arithmetic of every term.
//
// The hash is x-fastest, so the particles of one (y,z) grid row are one contiguous run of the sorted arrays.  A CTA owns kSB
// consecutive sorted slots = a few x-rows of one z-plane (or the tail of one plane and the head of the next): rows ya .. yb.  For
// the stencil slab dz its particles can only meet candidates of rows ya - rad .. yb + rad of plane z + dz, inside the x-window
// [xmin - rad, xmax + rad] of the CTA:
//   * per slab, ONE warp looks the R = yb - ya + 1 + 2 rad row windows up in the dense lower-bound table (a handful of loads per
//     row instead of two per lane per row) and issues one cp.async.bulk (TMA, completion on an mbarrier) per contiguous piece —
//     the window is laid out UNWRAPPED (a window that crosses the power-of-two grid's x-seam is two or three pieces placed one
//     behind the other), so every lane's candidate range of a row is one contiguous run of shared memory;
//   * all threads turn the lower-bound slices of the rows into a table of shared-memory offsets per unwrapped cell, so a lane's
//     row set-up is two 16-bit shared loads (no global loads, no wrap logic);
//   * each lane writes the non-empty ranges of the slab's <= 2 rad + 1 rows to its segment list and walks them in one flat loop
//     whose trip count is the warp's largest candidate total: candidates come out of shared memory (LDS.128, ~30 cycles
//     instead of an L1 / L2 gather), the distance test and the interaction are ONE predicated body (no accept queue, no votes,
//     no flush), and an accepted neighbour's slot goes straight to the lane's column of the slab's list region (ps_fluid_lists.cuh).
// The staged float4 carries its own sorted slot in .w (k_reorder writes it), so the walk never maps shared offsets back to slots.
// Per-particle row pruning (ext = sqrt(H^2 - dymin^2 - dzmin^2), conservative margin) is the grid-walking kernel's: the accepted
// sequence is exactly the reference's.
//
// A CTA the scheme does not cover — its slots touch three z-planes, its rows wrap in y, more than 32 rows or 3072 table entries
// (sparse spray: a few particles per row), a row window that exceeds the stage buffer (a strongly compressed pile) — walks the grid
// itself, lane by lane through L1 (the in-kernel fallback at the end of the kernel: same neighbour sequence, same arithmetic, same
// lists), so K7 never has to search.  Such CTAs are 2-6 % of C3's and sit where neighbourhoods are small; CTAs are issued from both
// ends of the sorted order towards the middle, so the slower ones (a scene's sparse fringe sorts to the ends) never form the tail.
#include <cuda/std/type_traits>
#include "ps_fluid_lists.cuh"

namespace {
#ifndef PS_SBLOCK
#define PS_SBLOCK 256      // threads = sorted slots per CTA
#endif
#ifndef PS_STAGE_CAP
#define PS_STAGE_CAP 1408  // float4 slots of the stage buffer (22 KB; with the tables 55 KB per CTA: four CTAs per SM)
#endif
#ifndef PS_SMINB
#define PS_SMINB 4         // resident CTAs the register allocation aims at
#endif
typedef unsigned long long u64;
typedef unsigned short u16;
constexpr int kSB = PS_SBLOCK, kSW = kSB / 32;
constexpr int kStageCap = PS_STAGE_CAP;
constexpr int kTableCap = 3072;  // u16 entries: rows of a chunk x (window cells + 1)
constexpr int kMaxStageRows = 32;
constexpr unsigned kFull = 0xffffffffu;
static_assert(kStageCap < 0xffff, "stage offsets are 16-bit");

struct RowWin {   // the x-window of one staged row of one slab: up to three contiguous pieces of the sorted arrays
    u32 ga[3];    // first sorted slot of piece A / B / C
    u32 c[3];     // candidates in it
};
enum { kNMin = 0, kNMax, kRMin, kRMax, kFits, kCtlWords = 16 };
// why CTAs left the staged path, counted per CTA since the last reset (diagnostics: ps_debug_staged_stats)
enum { kBailPlanes = 0, kBailRows, kBailTable, kBailStage, kBailNone /* CTAs that stayed */, kBailWords = 8 };
__device__ unsigned int g_staged_stats[kBailWords];

static inline size_t staged_smem_bytes(int rad) {
    return (size_t)kStageCap * sizeof(float4) + kTableCap * sizeof(u16) + (size_t)(2 * rad + 2) * kSB * sizeof(uint2) +
           (size_t)(2 * rad + 1) * kMaxStageRows * sizeof(RowWin) + kCtlWords * sizeof(int) + sizeof(u64);
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 done, polls = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++polls > (1u << 24)) __trap();  // a copy that never lands is an error, not a hang
    } while (!done);
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}

// One candidate of the walk.  Operands: accumulators ro, gxs, gys, gzs, denom (%0-%4), list write offset in bytes (%5), run state cur,
// end, sp (%6-%8), own position (%9-%11), own slot i (%12), cs (%13), trip t and this lane's candidate total (%14, %15), the
// run stride in bytes (%16), room (%17, capped walks only: the lane may still accept), the lane's first list entry (%18).
// Arithmetic, association for association, is what nvcc makes of the grid-walking kernel's C++ (ps_neighbor_kernels.cu:
// k_find_lambdas), so the two paths agree bit for bit (tests/test_gpu_parity.py::test_neighbour_list_paths_agree_bit_for_bit):
//   r2 = fma(rz, rz, fma(rx, rx, ry * ry));  ro += (hm2 * hm2) * hm2;  c = ((hm * cs) * hm) * rsqrt(r2);  g += r * c;  denom += (c * c) * r2
#define PS_K6_VISIT(EXTRA_PRED)                                                                                                  \
    asm volatile("{\n\t"                                                                                                          \
                 ".reg .pred p, q, e;\n\t"                                                                                        \
                 ".reg .f32 x, y, z, w, rx, ry, rz, r2, ir, rl, h2, hh, hm, a, c, cc;\n\t"                                         \
                 ".reg .b32 j;\n\t"                                                                                               \
                 ".reg .b64 wa;\n\t"                                                                                              \
                 "ld.shared.v4.f32 {x, y, z, w}, [%6];\n\t"                                                                       \
                 "add.u32 %6, %6, 16;\n\t"                                                                                        \
                 "setp.eq.u32 e, %6, %7;\n\t"                                                                                     \
                 "@e ld.shared.v2.u32 {%6, %7}, [%8];\n\t"                                                                        \
                 "@e add.u32 %8, %8, %16;\n\t"                                                                                    \
                 "mov.b32 j, w;\n\t"                                                                                              \
                 "sub.ftz.f32 rx, %9, x;\n\t"                                                                                     \
                 "sub.ftz.f32 ry, %10, y;\n\t"                                                                                    \
                 "sub.ftz.f32 rz, %11, z;\n\t"                                                                                    \
                 "mul.ftz.f32 r2, ry, ry;\n\t"                                                                                    \
                 "fma.rn.ftz.f32 r2, rx, rx, r2;\n\t"                                                                             \
                 "fma.rn.ftz.f32 r2, rz, rz, r2;\n\t"                                                                             \
                 "setp.lt.ftz.f32 p, r2, 0f40800000;\n\t"                                                                         \
                 "setp.lt.and.u32 p, %14, %15, p;\n\t"                                                                            \
                 "setp.ne.and.u32 p, j, %12, p;\n\t" EXTRA_PRED                                                                   \
                 "rsqrt.approx.ftz.f32 ir, r2;\n\t"                                                                               \
                 "mul.ftz.f32 rl, r2, ir;\n\t"                                                                                    \
                 "sub.ftz.f32 h2, 0f40800000, r2;\n\t"                                                                            \
                 "mul.ftz.f32 hh, h2, h2;\n\t"                                                                                    \
                 "@p fma.rn.ftz.f32 %0, h2, hh, %0;\n\t"                                                                          \
                 "sub.ftz.f32 hm, 0f40000000, rl;\n\t"                                                                            \
                 "setp.ge.and.ftz.f32 q, rl, 0f38D1B717, p;\n\t"                                                                  \
                 "mul.ftz.f32 a, hm, %13;\n\t"                                                                                    \
                 "mul.ftz.f32 a, hm, a;\n\t"                                                                                      \
                 "mul.ftz.f32 c, ir, a;\n\t"                                                                                      \
                 "@q fma.rn.ftz.f32 %1, rx, c, %1;\n\t"                                                                           \
                 "@q fma.rn.ftz.f32 %2, ry, c, %2;\n\t"                                                                           \
                 "@q fma.rn.ftz.f32 %3, rz, c, %3;\n\t"                                                                           \
                 "mul.ftz.f32 cc, c, c;\n\t"                                                                                      \
                 "@q fma.rn.ftz.f32 %4, r2, cc, %4;\n\t"                                                                          \
                 "mad.wide.u32 wa, %5, 1, %18;\n\t"                                                                               \
                 "@p st.global.cs.u32 [wa], j;\n\t"                                                                               \
                 "@p add.u32 %5, %5, 128;\n\t"                                                                                    \
                 "}"                                                                                                              \
                 : "+f"(ro), "+f"(gxs), "+f"(gys), "+f"(gzs), "+f"(denom), "+r"(woff), "+r"(cur), "+r"(end), "+r"(sp)              \
                 : "f"(pi.x), "f"(pi.y), "f"(pi.z), "r"(i), "f"(cs), "r"(t), "r"(total), "r"((u32)(kSB * sizeof(uint2))), "r"(room),       \
                   "l"(wfirst_p)                                                                                                  \
                 : "memory")

template <int RAD>
__global__ void __launch_bounds__(kSB, PS_SMINB) k_find_lambdas_staged(float *__restrict__ lambda, u32 *__restrict__ num_neighbors,
                                                                      const float4 *__restrict__ spos, const float *__restrict__ sw,
                                                                      const int *__restrict__ sphase, const u32 *__restrict__ index,
                                                                      const u32 *__restrict__ cell_begin, const float *__restrict__ ros, u32 n,
                                                                      u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g, StencilDesc st,
                                                                      int zero_nonfluid, u32 *__restrict__ pool, u32 *__restrict__ recs, u32 list_rows,
                                                                      size_t dump_offset) {
    extern __shared__ __align__(128) unsigned char staged_smem[];
    const int rad = RAD ? RAD : st.rad;
    const int S = 2 * rad + 1;
    float4 *buf = reinterpret_cast<float4 *>(staged_smem);
    u16 *table = reinterpret_cast<u16 *>(buf + kStageCap);
    // this lane's candidate runs of the current chunk as (first, one past last) shared byte addresses: run k at segs[k * kSB + tid],
    // S real runs at most + the closing one
    uint2 *segs = reinterpret_cast<uint2 *>(table + kTableCap);
    RowWin *wins = reinterpret_cast<RowWin *>(segs + (size_t)(S + 1) * kSB);  // [slab][row]
    int *ctl = reinterpret_cast<int *>(wins + (size_t)S * kMaxStageRows);
    u64 *mbar = reinterpret_cast<u64 *>(ctl + kCtlWords);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // CTAs in issue order: first, last, second, last but one, ... (the fringe of the sorted order first, the bulk last)
    const u32 cta = (blockIdx.x & 1u) ? gridDim.x - 1u - (blockIdx.x >> 1) : (blockIdx.x >> 1);
    const u32 i = cta * kSB + tid;
    const u32 warp = i >> 5;
    const bool warp_in = (u64)warp * 32 < n;
    u32 *rec = recs + (size_t)warp * kListRecord;

    bool act = i < n;
    u32 orig = 0;
    const float4 origin = make_float4(g.ox, g.oy, g.oz, 0.f);
    const float4 praw = act ? spos[i] : origin;
    if (act) {
        if (sphase[i] != PH_FLUID) {
            if (zero_nonfluid) lambda[i] = 0.f;
            act = false;
        } else {
            orig = index[i];
        }
    }
    // ghost copies of a neighbour slab's particles: lambda only inside [ghost_xmin, ghost_xmax] (see k_find_lambdas)
    if (act && orig >= n_owned && !(praw.x >= ghost_xmin && praw.x <= ghost_xmax)) act = false;
    const float4 pi = act ? praw : origin;

    if (tid == 0) {
        ctl[kNMin] = 0x7fffffff; ctl[kNMax] = -1; ctl[kRMin] = 0x7fffffff; ctl[kRMax] = -1;
        mbar_init(mbar, 32);
    }
    if (!__syncthreads_or(act)) {  // (also publishes ctl and the barrier)
        if (lane == 0 && warp_in) rec[0] = 0;
        return;
    }
    // ---------------- this lane's particle ----------------
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
    float dy2[2 * RAD + 1];
    if (RAD) {
#pragma unroll
        for (int d = 0; d < 2 * RAD + 1; d++) dy2[d] = dmin2(d - RAD, fy0, fy1, g.cy);
    }
    const float ro0 = act ? ros[orig] : 1.f;
    const float inv_ro0 = __fdividef(1.f, ro0);
    const float cs = -PS_SPIKY * inv_ro0;
    float ro = 0.f, denom = 0.f, gxs = 0.f, gys = 0.f, gzs = 0.f;
    // the warp's list region: row k, entry lane = the k-th accepted neighbour of this lane (ps_fluid_lists.cuh)
    u32 *wfirst_p = pool + (size_t)warp * list_rows * 32 + lane;
    u32 woff = 0;  // bytes written to this lane's column: 128 per accepted neighbour
    bool ovf = false;

    // The staged path; returns why the CTA cannot take it (CTA-uniform), or kBailNone.
    const int why = [&]() -> int {
    // ---------------- CTA geometry ----------------
    // The CTA's slots are rows yaA .. ybA of plane zA and, when they run across a plane boundary, rows yaB .. ybB of the next
    // occupied plane zB (two row groups; three planes or more go to the grid-walking kernel).  x-window [X0, X1] in unwrapped cells.
    const u32 ifirst = cta * kSB, ilast = min(ifirst + (u32)kSB, n) - 1u;
    const float4 pf = __ldg(spos + ifirst), pl = __ldg(spos + ilast);
    const int3 cf = ps_grid_pos(g, pf.x, pf.y, pf.z), cl = ps_grid_pos(g, pl.x, pl.y, pl.z);
    const u32 zA = (u32)cf.z & g.mz, zB = (u32)cl.z & g.mz;
    const int yaA = (int)((u32)cf.y & g.my), ybB = (int)((u32)cl.y & g.my);
    int ybA = ybB, yaB = 0;
    bool ok = zB >= zA;
    if (zB > zA) {
        const u32 sa = __ldg(cell_begin + (zA + 1u) * g.gy * g.gx), sb = __ldg(cell_begin + zB * g.gy * g.gx);  // first slot past plane zA / of plane zB
        ok = sa == sb && sb > ifirst && sb <= ilast;
        if (ok) {
            const float4 pe = __ldg(spos + sb - 1), ps = __ldg(spos + sb);
            ybA = (int)((u32)ps_grid_pos(g, pe.x, pe.y, pe.z).y & g.my);
            yaB = (int)((u32)ps_grid_pos(g, ps.x, ps.y, ps.z).y & g.my);
        }
    }
    if (!ok) return kBailPlanes;
    // staged rows: ya - rad .. yb + rad of each group, wrapped into the grid with '&' like every other cell coordinate
    const int RA = ybA - yaA + 1 + 2 * rad, RB = zB > zA ? ybB - yaB + 1 + 2 * rad : 0, R = RA + RB;
    if (ybA < yaA || (zB > zA && ybB < yaB) || R > kMaxStageRows) return kBailRows;
    int X0, X1;
    const int gxi = (int)g.gx;
    if (zA == zB && yaA == ybA) {  // one row: its slots ascend in x
        X0 = (int)((u32)cf.x & g.mx) - rad;
        X1 = (int)((u32)cl.x & g.mx) + rad;
    } else {  // several rows: extent of the CTA's own cells, on the circle cut at 0 or, failing that, at gx / 2
        const int3 cr = ps_grid_pos(g, praw.x, praw.y, praw.z);
        const u32 hx = i < n ? ((u32)cr.x & g.mx) : ((u32)cf.x & g.mx);
        const u32 hr = (hx + (g.gx >> 1)) & g.mx;
        const u32 a = __reduce_min_sync(kFull, hx), b = __reduce_max_sync(kFull, hx);
        const u32 c = __reduce_min_sync(kFull, hr), d = __reduce_max_sync(kFull, hr);
        if (lane == 0) {
            atomicMin(&ctl[kNMin], (int)a); atomicMax(&ctl[kNMax], (int)b);
            atomicMin(&ctl[kRMin], (int)c); atomicMax(&ctl[kRMax], (int)d);
        }
        __syncthreads();
        const int nmin = ctl[kNMin], nmax = ctl[kNMax], rmin = ctl[kRMin], rmax = ctl[kRMax];
        if (nmax - nmin + 1 + 2 * rad <= gxi) { X0 = nmin - rad; X1 = nmax + rad; }
        else if (rmax - rmin + 1 + 2 * rad <= gxi) { X0 = rmin - (gxi >> 1) - rad; X1 = rmax - (gxi >> 1) + rad; }
        else { X0 = -rad; X1 = gxi - 1 + rad; }
    }
    const int W = X1 - X0 + 1, TS = W + 1;
    if (TS > kTableCap) return kBailTable;
    // the three pieces of the window: unwrapped cells < 0 (A), in [0, gx) (B), >= gx (C); natural first cell and length of each
    const int nA = X0 < 0 ? min(X1, -1) - X0 + 1 : 0, aLo = X0 + gxi;
    const int bLo = max(X0, 0), nB = max(min(X1, gxi - 1) - bLo + 1, 0);
    const int nC = X1 >= gxi ? X1 - gxi + 1 : 0;
    auto row_base = [&](int s, int r) {  // cell index of x = 0 of staged row r in slab s
        const u32 zc = ((r < RA ? zA : zB) + (u32)(s - rad)) & g.mz;
        const u32 yc = (u32)(r < RA ? yaA - rad + r : yaB - rad + (r - RA)) & g.my;
        return (zc * g.gy + yc) * g.gx;
    };

    // ---------------- row windows of every slab, all warps at once: first slot and length of the (up to) three pieces ----------------
    for (int s = wid; s < S; s += kSW)
        if (lane < R) {
            const u32 rb = row_base(s, lane);
            RowWin w;
            w.ga[0] = w.ga[1] = w.ga[2] = 0; w.c[0] = w.c[1] = w.c[2] = 0;
            if (nA) { w.ga[0] = __ldg(cell_begin + rb + aLo); w.c[0] = __ldg(cell_begin + rb + aLo + nA) - w.ga[0]; }
            if (nB) { w.ga[1] = __ldg(cell_begin + rb + bLo); w.c[1] = __ldg(cell_begin + rb + bLo + nB) - w.ga[1]; }
            if (nC) { w.ga[2] = __ldg(cell_begin + rb); w.c[2] = __ldg(cell_begin + rb + nC) - w.ga[2]; }
            wins[s * kMaxStageRows + lane] = w;
        }

    // staged row of (own row - rad); stencil row dyi is staged row own_row + dyi.  Table column of the own cell: dx adds dx.
    const int own_row = ((u32)gp.z & g.mz) == zA ? (int)((u32)gp.y & g.my) - yaA : RA + (int)((u32)gp.y & g.my) - yaB;
    const int own_col = (int)((((u32)gp.x - (u32)(X0 + rad)) & g.mx)) + rad;
    const u32 buf_addr = smem_u32(buf), seg_addr = smem_u32(segs + tid);
    u32 phase = 0;  // mbarrier phase = chunks staged so far
    __syncthreads();  // wins

#pragma unroll 1
    for (int s = 0; s < S; s++) {
        const int dz = s - rad;
        // every warp derives the same chunking of the slab's rows from the same shared data: lane r = staged row r
        const RowWin *swins = wins + s * kMaxStageRows;
        const u32 rcnt = lane < R ? swins[lane].c[0] + swins[lane].c[1] + swins[lane].c[2] : 0u;
        u32 incl = rcnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (__shfl_sync(kFull, incl, 31) == 0) continue;  // nothing in reach in this slab (uniform over the CTA)
        const u32 rowmask = st.rowmask[s];  // uniform: stencil rows no particle can reach
        const float remz = PS_H2 - dmin2(dz, fz0, fz1, g.cz);
#pragma unroll 1
        for (int r0 = 0; r0 < R;) {
            // chunk [r0, r1): the longest run of rows whose candidates fit the stage buffer and whose windows fit the table
            const u32 before = __shfl_sync(kFull, incl - rcnt, r0);
            const bool fit = lane >= r0 && lane < R && incl - before <= (u32)kStageCap && (lane - r0 + 1) * TS <= kTableCap;
            const int r1 = r0 + __popc(__ballot_sync(kFull, fit));
            if (r1 == r0) return kBailStage;  // one row's window alone exceeds the stage buffer (uniform)
            const u32 chunk_total = __shfl_sync(kFull, incl, r1 - 1) - before;
            if (chunk_total == 0) { r0 = r1; continue; }
            __syncthreads();  // every warp is done with the previous chunk's stage buffer, table and runs
            const u32 soff = incl - rcnt - before;  // stage offset of this lane's row
            // ---- warp 0: one TMA bulk copy per non-empty piece of every row of the chunk ----
            if (wid == 0) {
                const bool mine = lane >= r0 && lane < r1 && rcnt;
                if (mine) mbar_arrive_expect_tx(mbar, rcnt * 16u);
                else mbar_arrive(mbar);
                if (mine) {
                    const RowWin w = swins[lane];
                    if (w.c[0]) bulk_g2s(buf + soff, spos + w.ga[0], w.c[0] * 16u, mbar);
                    if (w.c[1]) bulk_g2s(buf + soff + w.c[0], spos + w.ga[1], w.c[1] * 16u, mbar);
                    if (w.c[2]) bulk_g2s(buf + soff + w.c[0] + w.c[1], spos + w.ga[2], w.c[2] * 16u, mbar);
                }
            }
            // ---- all warps: stage offset of the first candidate at or after each unwrapped cell of each row of the chunk ----
            for (int r = r0 + wid; r < r1; r += kSW) {
                const u32 rb = row_base(s, r);
                const u32 so = __shfl_sync(kFull, soff, r);
                const RowWin w = swins[r];
                u16 *trow = table + (r - r0) * TS;
                const u32 *cbA = cell_begin + rb + aLo, *cbB = cell_begin + rb + bLo, *cbC = cell_begin + rb;
                for (int k = lane; k < nA; k += 32) trow[k] = (u16)(so + __ldg(cbA + k) - w.ga[0]);
                for (int k = lane; k < nB; k += 32) trow[nA + k] = (u16)(so + w.c[0] + __ldg(cbB + k) - w.ga[1]);
                for (int k = lane; k < nC; k += 32) trow[nA + nB + k] = (u16)(so + w.c[0] + w.c[1] + __ldg(cbC + k) - w.ga[2]);
                if (lane == 0) trow[W] = (u16)(so + w.c[0] + w.c[1] + w.c[2]);
            }
            __syncthreads();
            // ---- each lane: candidate ranges of its stencil rows inside the chunk -> runs ----
            u32 nlist = 0, total = 0;
            auto do_row = [&](int dyi) {
                if (!((rowmask >> dyi) & 1u)) return;
                const int rr = own_row + dyi - r0;  // row of the chunk
                const float rem = remz - (RAD ? dy2[RAD ? dyi : 0] : dmin2(dyi - rad, fy0, fy1, g.cy));
                if (act && rem >= 0.f && rr >= 0 && rr < r1 - r0) {
                    const float ext = sqrtf(rem) + eps;
                    int lo = (int)floorf((relx - ext) * inv_cx), hi = (int)floorf((relx + ext) * inv_cx);
                    lo = max(min(lo, gp.x), gp.x - rad);
                    hi = min(max(hi, gp.x), gp.x + rad);
                    const u16 *t = table + rr * TS + own_col;
                    const u32 b = t[lo - gp.x], e = t[hi - gp.x + 1];
                    if (e > b) {
                        segs[nlist * kSB + tid] = make_uint2(buf_addr + b * 16u, buf_addr + e * 16u);
                        nlist++;
                        total += e - b;
                    }
                }
            };
            if (RAD) {
#pragma unroll
                for (int dyi = 0; dyi < 2 * RAD + 1; dyi++) do_row(dyi);
            } else {
#pragma unroll 1
                for (int dyi = 0; dyi <= 2 * rad; dyi++) do_row(dyi);
            }
            // closing run: never ends; a lane that has run out of candidates re-reads the head of the stage buffer (always staged
            // data: a lane's total never exceeds the chunk's) and the live test discards what it reads
            segs[nlist * kSB + tid] = make_uint2(buf_addr, 0xffffffffu);
            const u32 maxtotal = __reduce_max_sync(kFull, total);
            // a lane's list would outgrow the warp's region (only when neighbor_list_rows < 500): the warp keeps no list from here on
            // and K7 walks the grid for it; its writes land in the dump region behind the pool
            if (!ovf && __any_sync(kFull, min((woff >> 7) + total, PS_MAX_NEIGHBORS) > list_rows)) {
                ovf = true;
                wfirst_p = pool + dump_offset + lane - (woff >> 2);
            }
            mbar_wait(mbar, phase & 1u);  // the chunk's candidates have landed (every thread observes every phase)
            phase++;
            r0 = r1;
            if (!maxtotal) continue;
            // ---- flat walk: trip count = the warp's largest candidate total ----
            // One candidate = one block of predicated PTX (PS_K6_VISIT): stage-buffer load, run bookkeeping, distance test and
            // interaction with no branch — the compiler's own form of the same C++ is a divergent region (BSSY / BRA / BSYNC) around
            // the interaction plus re-materialised base addresses, 47 SASS instructions per candidate against 37 here.
            const bool capped = __any_sync(kFull, (woff >> 7) + total > PS_MAX_NEIGHBORS);  // the 500-neighbour cap can bite in this chunk
            const uint2 first_run = segs[tid];
            u32 cur = first_run.x, end = first_run.y, sp = seg_addr + (u32)(kSB * sizeof(uint2));
            if (!capped) {
                const u32 room = 1u;
#pragma unroll 4
                for (u32 t = 0; t < maxtotal; t++) PS_K6_VISIT("");
            } else {
#pragma unroll 1
                for (u32 t = 0; t < maxtotal; t++) {
                    const u32 room = (woff >> 7) < PS_MAX_NEIGHBORS;
                    PS_K6_VISIT("setp.ne.and.u32 p, %17, 0, p;\n\t");
                }
            }
        }
    }
    return kBailNone;
    }();
    if (tid == 0) atomicAdd(&g_staged_stats[why], 1u);

    if (why != kBailNone) {
        // ---------------- in-kernel fallback: this lane walks the grid itself ----------------
        // Rows in the reference's order with the same per-particle pruning, candidates through L1, the interaction written with
        // explicit roundings in PS_K6_VISIT's association (the two paths agree bit for bit), accepted neighbours to the lane's column.
        ro = denom = gxs = gys = gzs = 0.f;
        u32 cnt = 0;
        bool lost = false;  // an accepted neighbour that found no room in the column (only when neighbor_list_rows < 500)
        u32 *col = pool + (size_t)warp * list_rows * 32 + lane;
        auto visit = [&](u32 j) {
            const float4 pj = __ldg(spos + j);
            const float rx = __fsub_rn(pi.x, pj.x), ry = __fsub_rn(pi.y, pj.y), rz = __fsub_rn(pi.z, pj.z);
            const float r2 = __fmaf_rn(rz, rz, __fmaf_rn(rx, rx, __fmul_rn(ry, ry)));
            if (r2 < PS_H2 && j != i && cnt < PS_MAX_NEIGHBORS) {
                ps_lambda_terms(rx, ry, rz, cs, ro, gxs, gys, gzs, denom);
                if (cnt < list_rows) __stcs(col + (size_t)cnt * 32, j);
                else lost = true;
                cnt++;
            }
        };
        if (act) {
#pragma unroll 1
            for (int dz = -rad; dz <= rad; dz++) {
                const u32 rowmask = st.rowmask[dz + rad];
                const u32 zrow = ((u32)(gp.z + dz) & g.mz) * g.gy;
                const float remz = PS_H2 - dmin2(dz, fz0, fz1, g.cz);
#pragma unroll 1
                for (int dyi = 0; dyi < S; dyi++) {
                    if (!((rowmask >> dyi) & 1u)) continue;
                    const float rem = remz - dmin2(dyi - rad, fy0, fy1, g.cy);
                    if (!(rem >= 0.f)) continue;
                    const float ext = sqrtf(rem) + eps;
                    int lo = (int)floorf((relx - ext) * inv_cx), hi = (int)floorf((relx + ext) * inv_cx);
                    lo = max(min(lo, gp.x), gp.x - rad);
                    hi = min(max(hi, gp.x), gp.x + rad);
                    const u32 row = (zrow + ((u32)(gp.y + dyi - rad) & g.my)) * g.gx;
                    const u32 lw = (u32)lo & g.mx, hw = (u32)hi & g.mx;
                    const bool wrap = lw > hw;  // the row wraps around the power-of-two grid: [lw, gx) then [0, hw]
                    u32 b = __ldg(cell_begin + row + lw), e = __ldg(cell_begin + row + (wrap ? g.mx : hw) + 1u);
                    for (u32 j = b; j < e; j++) visit(j);
                    if (wrap) {
                        b = __ldg(cell_begin + row);
                        e = __ldg(cell_begin + row + hw + 1u);
                        for (u32 j = b; j < e; j++) visit(j);
                    }
                }
            }
        }
        woff = cnt << 7;
        ovf = __any_sync(kFull, lost);
    }
    if (lane == 0 && warp_in) rec[0] = ovf ? kListOverflow : 1u;
    if (!act) return;
    const float inv_w = __fdividef(1.f, sw[i]);
    lambda[i] = ps_lambda_from_sums(ro, denom, gxs, gys, gzs, inv_w, inv_ro0);
    num_neighbors[i] = woff >> 7;
}
}  // namespace

u32 ps_staged_dump_rows() { return PS_MAX_NEIGHBORS + 8u; }  // an overflowed warp still writes at most 500 rows

// CTAs of the staged K6 since the last reset, by outcome: [0] three z-planes / empty-plane gap, [1] too many rows or rows that wrap
// in y, [2] window table too large, [3] a slab's windows exceeded the stage buffer, [4] stayed on the staged path
extern "C" int ps_debug_staged_stats(unsigned int out[8], int reset) {
    unsigned int h[kBailWords] = {};
    if (cudaMemcpyFromSymbol(h, g_staged_stats, sizeof h) != cudaSuccess) return -1;
    if (out) for (int k = 0; k < kBailWords; k++) out[k] = h[k];
    if (reset) { unsigned int z[kBailWords] = {}; if (cudaMemcpyToSymbol(g_staged_stats, z, sizeof z) != cudaSuccess) return -1; }
    return 0;
}

void ps_launch_find_lambdas_staged(float *lambda, u32 *num_neighbors, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                                   const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g,
                                   const StencilDesc &st, bool zero_nonfluid, u32 *pool, u32 *recs, u32 list_rows, size_t dump_offset, int device,
                                   cudaStream_t s) {
    if (!n) return;
    // the opt-in to > 48 KB of dynamic shared memory is per device: once for each device this process uses
    static bool opted[64] = {};
    if (device >= 0 && device < 64 && !opted[device]) {
        cudaFuncSetAttribute(k_find_lambdas_staged<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem_bytes(4));
        cudaFuncSetAttribute(k_find_lambdas_staged<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem_bytes(PS_MAX_RAD));
        opted[device] = true;
    }
    const u32 blocks = (n + kSB - 1) / kSB;
    if (st.rad == 4)  // the reference's configuration (H = 2, cell = 2r = 0.5): row loop fully unrolled
        k_find_lambdas_staged<4><<<blocks, kSB, staged_smem_bytes(4), s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n, n_owned,
                                                                           ghost_xmin, ghost_xmax, g, st, zero_nonfluid ? 1 : 0, pool, recs, list_rows,
                                                                           dump_offset);
    else
        k_find_lambdas_staged<0><<<blocks, kSB, staged_smem_bytes(st.rad), s>>>(lambda, num_neighbors, spos, sw, sphase, index, cell_begin, ros, n,
                                                                                n_owned, ghost_xmin, ghost_xmax, g, st, zero_nonfluid ? 1 : 0, pool, recs,
                                                                                list_rows, dump_offset);
}
