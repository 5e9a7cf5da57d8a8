// particlesolver_b200/csrc/ps_slab_kernels.cu — spatial slab decomposition: ghost-halo packing, particle migration.
//
// New work (the reference is single-GPU).  A context owns the particles whose x lies in its slab
// [x_lo, x_hi) and keeps ghost copies of its neighbours' particles near the two faces behind them in the same SoA
// arrays (indices [n_owned, n)).  Ghosts take part in the grid build and are read as neighbours; they are never moved.
// Everything here is ORDERED (a stable stream compaction: flags -> exclusive scan -> scatter), so a run is
// reproducible and the in-cell neighbour order (ascending original index) does not depend on scheduling.
//
//   halo record    32 B : pos4 | w | rest density | phase | pad          (what a neighbour needs of a ghost)
//   migrant record 64 B : pos4 | prev4 | vel4 | w | rest density | phase | pad   (the full state of a particle)
#include "ps_common.cuh"

namespace {
constexpr int kBlock = 256;
constexpr int kItems = 8;
constexpr int kTile = kBlock * kItems;  // 2048 particles per CTA

struct HaloRec { float4 pos; float w, ros; int phase; u32 pad; };
struct MigrantRec { float4 pos, prev, vel; float w, ros; int phase; u32 pad; };
static_assert(sizeof(HaloRec) == 32 && sizeof(MigrantRec) == 64, "record sizes are part of the wire format");
constexpr u32 kNoRank = 0xffffffffu;

// class of a particle: 0 = stays / not selected, 1 = left buffer, 2 = right buffer (3 = both, halo of a thin slab)
__device__ __forceinline__ u32 classify(float x, float left_below, float right_from) { return (x < left_below ? 1u : 0u) | (x >= right_from ? 2u : 0u); }

__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *sm, u32 &total) {  // exclusive scan over the CTA's threads
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) {
        const u32 c = sm[w];
        if (w < wid) base += c;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return base + incl - v;
}

// A tile's items are interleaved over the CTA — item (k, t) is particle tile_base + k * kBlock + t, so a warp's loads are contiguous
// (a thread owning 8 CONSECUTIVE particles, the first layout, made every load instruction touch 32 different lines) — and are ranked
// in PARTICLE order: rank(k, t) = number of flagged items (k', t') of the tile with (k', t') < (k, t), i.e. ballots inside a row of
// 32, an exclusive scan over the kItems x kWarps row counts.  sm: 2 * kItems * kWarps + 1 words.
constexpr int kWarps = kBlock / 32;
constexpr int kRankSmem = 2 * kItems * kWarps + 1;
__device__ __forceinline__ void tile_excl_ranks(const bool (&f)[kItems], u32 *sm, u32 (&rank)[kItems], u32 &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    u32 b[kItems];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        b[k] = __ballot_sync(0xffffffffu, f[k]);
        if (lane == 0) sm[k * kWarps + wid] = __popc(b[k]);
    }
    __syncthreads();
    constexpr int kCounts = kItems * kWarps;  // 64: two per lane of warp 0
    static_assert(kCounts == 64, "the scan below takes two counts per lane");
    if (wid == 0) {
        const u32 c0 = sm[2 * lane], c1 = sm[2 * lane + 1];
        u32 incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const u32 excl = incl - (c0 + c1);
        sm[kCounts + 2 * lane] = excl;
        sm[kCounts + 2 * lane + 1] = excl + c0;
        if (lane == 31) sm[2 * kCounts] = incl;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kItems; k++) rank[k] = sm[kCounts + k * kWarps + wid] + __popc(b[k] & lt);
    total = sm[2 * kCounts];
    __syncthreads();  // sm may be reused by the next call
}

// pass 1: per-tile counts of class-1 (left) and class-2 (right) particles -> tile_counts[2*tile + {0,1}]
__global__ void __launch_bounds__(kBlock) k_slab_count(const float4 *__restrict__ pos, u32 n, float left_below, float right_from,
                                                       u32 *__restrict__ tile_counts) {
    __shared__ u32 sm[kBlock / 32];
    const u32 base = blockIdx.x * kTile + threadIdx.x;
    u32 cl = 0, cr = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const u32 i = base + k * kBlock;
        if (i < n) {
            const u32 c = classify(__ldg(&pos[i].x), left_below, right_from);
            cl += c & 1u;
            cr += c >> 1;
        }
    }
    u32 tl, tr;
    block_excl_scan(cl, sm, tl);
    block_excl_scan(cr, sm, tr);
    if (threadIdx.x == 0) { tile_counts[2 * blockIdx.x] = tl; tile_counts[2 * blockIdx.x + 1] = tr; }
}

// pass 2 (one CTA): exclusive scan of the tile counts in place; totals[0..1] = number of left / right records
__global__ void __launch_bounds__(1024) k_slab_scan_tiles(u32 *__restrict__ tile_counts, u32 tiles, u32 *__restrict__ totals) {
    __shared__ u32 sm[2][32];
    __shared__ u32 carry[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (u32 t0 = 0; t0 < tiles; t0 += 1024) {
        const u32 t = t0 + threadIdx.x;
        u32 v[2] = {t < tiles ? tile_counts[2 * t] : 0u, t < tiles ? tile_counts[2 * t + 1] : 0u};
        u32 incl[2] = {v[0], v[1]};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, incl[0], o), b = __shfl_up_sync(0xffffffffu, incl[1], o);
            if (lane >= o) { incl[0] += a; incl[1] += b; }
        }
        if (lane == 31) { sm[0][wid] = incl[0]; sm[1][wid] = incl[1]; }
        __syncthreads();
        u32 base[2] = {carry[0], carry[1]}, tot[2] = {0, 0};
        for (int w = 0; w < 32; w++) {
            if (w < wid) { base[0] += sm[0][w]; base[1] += sm[1][w]; }
            tot[0] += sm[0][w]; tot[1] += sm[1][w];
        }
        if (t < tiles) { tile_counts[2 * t] = base[0] + incl[0] - v[0]; tile_counts[2 * t + 1] = base[1] + incl[1] - v[1]; }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] += tot[0]; carry[1] += tot[1]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

// pass 3a: halo records of the selected particles, in ascending particle index
__global__ void __launch_bounds__(kBlock) k_slab_pack_halo(const float4 *__restrict__ pos, const float *__restrict__ w, const float *__restrict__ ros,
                                                           const int *__restrict__ phase, u32 n, float left_below, float right_from,
                                                           const u32 *__restrict__ tile_offsets, HaloRec *__restrict__ left, HaloRec *__restrict__ right,
                                                           u32 cap, uint2 *__restrict__ ranks) {
    __shared__ u32 sm[kRankSmem];
    const u32 base = blockIdx.x * kTile + threadIdx.x;
    u32 cls[kItems];
    bool fl[kItems], fr[kItems];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const u32 i = base + k * kBlock;
        cls[k] = i < n ? classify(__ldg(&pos[i].x), left_below, right_from) : 0u;
        fl[k] = cls[k] & 1u;
        fr[k] = cls[k] & 2u;
    }
    u32 rl[kItems], rr[kItems], dummy;
    tile_excl_ranks(fl, sm, rl, dummy);
    tile_excl_ranks(fr, sm, rr, dummy);
    const u32 tl = tile_offsets[2 * blockIdx.x], tr = tile_offsets[2 * blockIdx.x + 1];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const u32 i = base + k * kBlock;
        const u32 ol = tl + rl[k], orr = tr + rr[k];
        // where particle i went in the two buffers (kNoRank: not selected), for the lambda exchange that follows K6
        if (ranks && i < n) ranks[i] = make_uint2((cls[k] & 1u) ? ol : kNoRank, (cls[k] & 2u) ? orr : kNoRank);
        if (!cls[k]) continue;
        HaloRec r;
        r.pos = pos[i]; r.w = w[i]; r.ros = ros[i]; r.phase = phase[i]; r.pad = 0;
        if ((cls[k] & 1u) && ol < cap) left[ol] = r;
        if ((cls[k] & 2u) && orr < cap) right[orr] = r;
    }
}

// ---- lambda exchange (the K6 -> K7 dependency across a face) ----
// K7 reads lambda_j of every neighbour j of an owned particle; for a ghost j that value belongs to the rank that owns j.
// After K6 each rank sends the lambdas of the particles it put into the halo buffers — in the order of those records, so
// lambda k of a message belongs to ghost k of the receiver — and the receiver drops them into the ghosts' sorted slots.
// lambda lives by SORTED slot (index[slot] = particle), the halo ranks by particle.
__global__ void __launch_bounds__(kBlock) k_slab_pack_lambda(const float *__restrict__ lambda, const u32 *__restrict__ index, const uint2 *__restrict__ ranks,
                                                             u32 n, u32 n_owned, float *__restrict__ left, float *__restrict__ right, u32 cap) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const u32 orig = __ldg(index + i);
    if (orig >= n_owned) return;
    const uint2 r = __ldg(ranks + orig);
    if (r.x == kNoRank && r.y == kNoRank) return;
    const float l = lambda[i];
    if (r.x < cap) left[r.x] = l;
    if (r.y < cap) right[r.y] = l;
}
__global__ void __launch_bounds__(kBlock) k_slab_unpack_lambda(float *__restrict__ lambda, const u32 *__restrict__ index, u32 n, u32 n_owned,
                                                               const float *__restrict__ from_left, u32 n_left, const float *__restrict__ from_right) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const u32 orig = __ldg(index + i);
    if (orig < n_owned) return;
    const u32 k = orig - n_owned;  // ghosts sit behind the owned particles, the left neighbour's first (k_slab_unpack_halo)
    lambda[i] = k < n_left ? __ldg(from_left + k) : __ldg(from_right + (k - n_left));
}

// ghosts received from the two neighbours -> the tails of the SoA arrays, left neighbour's first
__global__ void __launch_bounds__(kBlock) k_slab_unpack_halo(float4 *__restrict__ pos, float *__restrict__ w, float *__restrict__ ros, int *__restrict__ phase,
                                                             u32 first, const HaloRec *__restrict__ from_left, u32 n_left,
                                                             const HaloRec *__restrict__ from_right, u32 n_right) {
    const u32 k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= n_left + n_right) return;
    const HaloRec r = k < n_left ? from_left[k] : from_right[k - n_left];
    const u32 i = first + k;
    pos[i] = r.pos; w[i] = r.w; ros[i] = r.ros; phase[i] = r.phase;
}

// pass 3b: migrant records (class 1 -> left, class 2 -> right; a particle cannot be both: x_lo < x_hi)
__global__ void __launch_bounds__(kBlock) k_slab_pack_migrants(const float4 *__restrict__ pos, const float4 *__restrict__ prev, const float4 *__restrict__ vel,
                                                               const float *__restrict__ w, const float *__restrict__ ros, const int *__restrict__ phase,
                                                               u32 n, float left_below, float right_from, const u32 *__restrict__ tile_offsets,
                                                               MigrantRec *__restrict__ left, MigrantRec *__restrict__ right, u32 cap) {
    __shared__ u32 sm[kRankSmem];
    const u32 base = blockIdx.x * kTile + threadIdx.x;
    u32 cls[kItems];
    bool fl[kItems], fr[kItems];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const u32 i = base + k * kBlock;
        cls[k] = i < n ? classify(__ldg(&pos[i].x), left_below, right_from) : 0u;
        fl[k] = cls[k] & 1u;
        fr[k] = (cls[k] & 1u) == 0 && (cls[k] & 2u);
    }
    u32 rl[kItems], rr[kItems], dummy;
    tile_excl_ranks(fl, sm, rl, dummy);
    tile_excl_ranks(fr, sm, rr, dummy);
    const u32 tl = tile_offsets[2 * blockIdx.x], tr = tile_offsets[2 * blockIdx.x + 1];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        if (!cls[k]) continue;
        const u32 i = base + k * kBlock;
        MigrantRec r;
        r.pos = pos[i]; r.prev = prev[i]; r.vel = vel[i]; r.w = w[i]; r.ros = ros[i]; r.phase = phase[i]; r.pad = 0;
        if (cls[k] & 1u) { if (tl + rl[k] < cap) left[tl + rl[k]] = r; }
        else { if (tr + rr[k] < cap) right[tr + rr[k]] = r; }
    }
}

// stable compaction of the stayers: element i moves to i - (#migrants before i).  tile_offsets holds the exclusive
// left/right migrant counts per tile, so the destination is known without another scan.
template <class T>
__global__ void __launch_bounds__(kBlock) k_slab_compact(const float4 *__restrict__ pos, const T *__restrict__ src, T *__restrict__ dst, u32 n,
                                                         float left_below, float right_from, const u32 *__restrict__ tile_offsets) {
    __shared__ u32 sm[kRankSmem];
    const u32 base = blockIdx.x * kTile + threadIdx.x;
    bool gone[kItems];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const u32 i = base + k * kBlock;
        gone[k] = i < n && classify(__ldg(&pos[i].x), left_below, right_from) != 0u;
    }
    u32 rg[kItems], dummy;
    tile_excl_ranks(gone, sm, rg, dummy);
    const u32 before_tile = tile_offsets[2 * blockIdx.x] + tile_offsets[2 * blockIdx.x + 1];
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const u32 i = base + k * kBlock;
        if (i < n && !gone[k]) dst[i - (before_tile + rg[k])] = src[i];
    }
}

__global__ void __launch_bounds__(kBlock) k_slab_append_migrants(float4 *__restrict__ pos, float4 *__restrict__ prev, float4 *__restrict__ vel, float *__restrict__ w,
                                                                 float *__restrict__ ros, int *__restrict__ phase, u32 first,
                                                                 const MigrantRec *__restrict__ from_left, u32 n_left,
                                                                 const MigrantRec *__restrict__ from_right, u32 n_right) {
    const u32 k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= n_left + n_right) return;
    const MigrantRec r = k < n_left ? from_left[k] : from_right[k - n_left];
    const u32 i = first + k;
    pos[i] = r.pos; prev[i] = r.prev; vel[i] = r.vel; w[i] = r.w; ros[i] = r.ros; phase[i] = r.phase;
}
}  // namespace

static inline u32 cdiv(u32 a, u32 b) { return (a + b - 1) / b; }

size_t ps_slab_scratch_elems(u32 n) { return 2 * (size_t)cdiv(n ? n : 1, kTile) + 2; }

// counts + exclusive tile offsets for the split (x < left_below | x >= right_from); totals land in scratch[2*tiles .. +1]
void ps_launch_slab_select(const float4 *pos, u32 n, float left_below, float right_from, u32 *scratch, cudaStream_t s) {
    const u32 tiles = cdiv(n ? n : 1, kTile);
    k_slab_count<<<tiles, kBlock, 0, s>>>(pos, n, left_below, right_from, scratch);
    k_slab_scan_tiles<<<1, 1024, 0, s>>>(scratch, tiles, scratch + 2 * tiles);
}
void ps_launch_slab_pack_halo(const float4 *pos, const float *w, const float *ros, const int *phase, u32 n, float left_below, float right_from,
                              const u32 *scratch, void *left, void *right, u32 cap, cudaStream_t s, u32 *ranks) {
    if (!n) return;
    k_slab_pack_halo<<<cdiv(n, kTile), kBlock, 0, s>>>(pos, w, ros, phase, n, left_below, right_from, scratch, (HaloRec *)left, (HaloRec *)right, cap,
                                                      (uint2 *)ranks);
}
void ps_launch_slab_pack_lambda(const float *lambda, const u32 *index, const u32 *ranks, u32 n, u32 n_owned, float *left, float *right, u32 cap,
                                cudaStream_t s) {
    if (!n) return;
    k_slab_pack_lambda<<<cdiv(n, kBlock), kBlock, 0, s>>>(lambda, index, (const uint2 *)ranks, n, n_owned, left, right, cap);
}
void ps_launch_slab_unpack_lambda(float *lambda, const u32 *index, u32 n, u32 n_owned, const float *from_left, u32 n_left, const float *from_right,
                                  cudaStream_t s) {
    if (!n || n == n_owned) return;
    k_slab_unpack_lambda<<<cdiv(n, kBlock), kBlock, 0, s>>>(lambda, index, n, n_owned, from_left, n_left, from_right);
}
void ps_launch_slab_unpack_halo(float4 *pos, float *w, float *ros, int *phase, u32 first, const void *from_left, u32 n_left, const void *from_right,
                                u32 n_right, cudaStream_t s) {
    if (!(n_left + n_right)) return;
    k_slab_unpack_halo<<<cdiv(n_left + n_right, kBlock), kBlock, 0, s>>>(pos, w, ros, phase, first, (const HaloRec *)from_left, n_left,
                                                                        (const HaloRec *)from_right, n_right);
}
void ps_launch_slab_pack_migrants(const float4 *pos, const float4 *prev, const float4 *vel, const float *w, const float *ros, const int *phase, u32 n,
                                  float left_below, float right_from, const u32 *scratch, void *left, void *right, u32 cap, cudaStream_t s) {
    if (!n) return;
    k_slab_pack_migrants<<<cdiv(n, kTile), kBlock, 0, s>>>(pos, prev, vel, w, ros, phase, n, left_below, right_from, scratch, (MigrantRec *)left,
                                                          (MigrantRec *)right, cap);
}
void ps_launch_slab_compact4(const float4 *pos, const float4 *src, float4 *dst, u32 n, float left_below, float right_from, const u32 *scratch,
                             cudaStream_t s) {
    if (!n) return;
    k_slab_compact<float4><<<cdiv(n, kTile), kBlock, 0, s>>>(pos, src, dst, n, left_below, right_from, scratch);
}
void ps_launch_slab_compact1(const float4 *pos, const u32 *src, u32 *dst, u32 n, float left_below, float right_from, const u32 *scratch,
                             cudaStream_t s) {
    if (!n) return;
    k_slab_compact<u32><<<cdiv(n, kTile), kBlock, 0, s>>>(pos, src, dst, n, left_below, right_from, scratch);
}
void ps_launch_slab_append_migrants(float4 *pos, float4 *prev, float4 *vel, float *w, float *ros, int *phase, u32 first, const void *from_left,
                                    u32 n_left, const void *from_right, u32 n_right, cudaStream_t s) {
    if (!(n_left + n_right)) return;
    k_slab_append_migrants<<<cdiv(n_left + n_right, kBlock), kBlock, 0, s>>>(pos, prev, vel, w, ros, phase, first, (const MigrantRec *)from_left,
                                                                            n_left, (const MigrantRec *)from_right, n_right);
}

// ---- load balancing: histogram of the owned particles' x over [x_min, x_max) (bins <= 65536), for re-cutting the slabs ----
namespace {
__global__ void __launch_bounds__(256) k_slab_x_histogram(const float4 *__restrict__ pos, u32 n, float x_min, float inv_width, u32 bins, u32 *__restrict__ hist) {
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float t = (__ldg(&pos[i].x) - x_min) * inv_width;
    int b = (int)floorf(t);
    b = b < 0 ? 0 : (b >= (int)bins ? (int)bins - 1 : b);
    atomicAdd(hist + b, 1u);  // x-sorted-ish input: neighbouring threads hit neighbouring bins, few conflicts
}
}  // namespace
void ps_launch_slab_x_histogram(const float4 *pos, u32 n, float x_min, float x_max, u32 bins, u32 *hist, cudaStream_t s) {
    cudaMemsetAsync(hist, 0, (size_t)bins * sizeof(u32), s);
    if (!n) return;
    k_slab_x_histogram<<<(n + 255) / 256, 256, 0, s>>>(pos, n, x_min, (float)bins / (x_max - x_min), bins, hist);
}

// A few 32-bit words between device memory and PINNED host memory (either direction) by a kernel instead of a copy-engine
// transfer: the counts a slab step waits for must not queue behind a bulk host transfer of a streamed step's results on the same
// engine (ps_step_streamed / ps_io_end: 2 x 128 MB at 8M particles held every count fetch for 2.7 ms, profiles/r2s).
namespace {
__global__ void k_copy_words(u32 *__restrict__ dst, const u32 *__restrict__ src, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
    __threadfence_system();
}
}  // namespace
void ps_launch_copy_words(u32 *dst, const u32 *src, u32 n, cudaStream_t s) {
    if (!n) return;
    k_copy_words<<<(n + 63) / 64, 64, 0, s>>>(dst, src, n);
}
