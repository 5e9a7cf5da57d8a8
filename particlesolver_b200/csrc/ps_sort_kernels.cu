// particlesolver_b200/csrc/ps_sort_kernels.cu — K3: stable LSD radix sort of (cell key, particle index) pairs.
//
// Replaces the reference's thrust::sort_by_key (integration.cu:270-275).  Contract:
// ascending by key, ties in ascending original index — i.e. a STABLE sort, which is what thrust's radix
// dispatch delivers and what the reference's cellStart/cellEnd + neighbour traversal order depend on.
//
// Design (single-pass-per-digit "onesweep" with decoupled look-back, hand-written for sm_100a):
//   * one upfront kernel builds the global 256-bin histogram of EVERY digit place in one read of the keys;
//   * per 8-bit digit ONE kernel: each CTA takes a tile of 4096 pairs through a dynamic ticket, ranks them
//     stably in-warp with match.any, publishes its per-digit counts in one 32-bit status word per digit
//     (2 flag bits | 30 count bits — flag and value travel together, so no fence protocol is needed),
//     looks back over predecessor tiles to turn them into exclusive prefixes, stages the tile digit-sorted
//     in shared memory and streams it out in coalesced per-digit runs.
//   * only ceil(log2(cells)/8) passes run (the reference sorts all 32 bits although 18 are live).
// HBM traffic: 4 B (histogram) + 16 B per pass per pair — the model in SURVEY §8(d).
#include "ps_common.cuh"

#ifndef PS_SORT_LOOK
#define PS_SORT_LOOK 8
#endif
namespace {
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;                       // keys per thread (8, i.e. 2048-pair tiles, was measured slower even at 1M pairs: 94 vs 86 us per sort)
constexpr int kTile = kThreads * kItems;         // 4096 pairs per CTA
constexpr int kWarpTile = 32 * kItems;           // 512 contiguous pairs per warp (keeps the rank order stable)
constexpr u32 kFlagAgg = 1u << 30, kFlagPrefix = 2u << 30, kValMask = (1u << 30) - 1;

// global histograms of all digit places, one pass over the keys
__global__ void __launch_bounds__(kThreads) k_radix_hist(const u32 *__restrict__ keys, u32 n, int passes, u32 *__restrict__ hist) {
    __shared__ u32 sh[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += kThreads) sh[i] = 0;
    __syncthreads();
    u32 stride = gridDim.x * kThreads * 4;
    for (u32 i = (blockIdx.x * kThreads + threadIdx.x) * 4; i < n; i += stride) {
        u32 k[4];
        if (i + 4 <= n) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(keys + i));
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
        } else {
            for (int j = 0; j < 4; j++) k[j] = (i + j < n) ? keys[i + j] : 0xffffffffu;
        }
        // Cell keys of consecutive particles share their high digits: when a digit is uniform across the warp one lane
        // adds 32, otherwise plain shared atomics (few conflicts: the digit varies).  No match.any, which is slow.
        const u32 active = __activemask();
        const u32 lane = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool ok = i + j < n;
            for (int p = 0; p < passes; p++) {
                const u32 d = (k[j] >> (8 * p)) & 255u;
                bool uniform = false;
                if (active == 0xffffffffu) {  // whole warp in the loop (always, except in the last grid-stride round)
                    const u32 d0 = __shfl_sync(0xffffffffu, d, 0);
                    uniform = __all_sync(0xffffffffu, ok && d == d0);
                }
                if (uniform) {
                    if (lane == 0) atomicAdd(&sh[p * 256 + d], 32u);
                } else if (ok) {
                    atomicAdd(&sh[p * 256 + d], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += kThreads) {
        u32 c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template <bool IDENTITY_VALS>
__global__ void __launch_bounds__(kThreads) k_radix_pass(const u32 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
                                                         u32 *__restrict__ keys_out, u32 *__restrict__ vals_out, u32 n, int shift,
                                                         const u32 *__restrict__ hist, u32 *__restrict__ status, u32 *__restrict__ ticket) {
    __shared__ u32 s_keys[kTile];
    __shared__ u32 s_vals[kTile];
    __shared__ u32 s_warp_cnt[kWarps][256];  // per-warp digit counts, then exclusive over warps
    __shared__ u32 s_tile_off[256];          // exclusive scan of tile digit counts (position in s_keys)
    __shared__ int s_gofs[256];              // global position of a digit's run minus its tile offset
    __shared__ u32 s_tile;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < kWarps * 256; i += kThreads) (&s_warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 tile_base = tile * kTile;
    const u32 nt = min((u32)kTile, n - tile_base);  // valid pairs in this tile

    // ---- load: warp w owns the contiguous run [w*512, w*512+512) of the tile, item k at k*32+lane ----
    u32 key[kItems], val[kItems];
    const u32 wbase = wid * kWarpTile;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        u32 loc = wbase + k * 32 + lane;
        bool ok = loc < nt;
        key[k] = ok ? __ldg(keys_in + tile_base + loc) : 0xffffffffu;
        if (IDENTITY_VALS) val[k] = tile_base + loc;
        else val[k] = ok ? __ldg(vals_in + tile_base + loc) : 0u;
    }

    // ---- stable in-warp ranking: rank = (#earlier items of this warp with the same digit) ----
    u32 rank[kItems];
    const u32 lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        u32 loc = wbase + k * 32 + lane;
        bool ok = loc < nt;
        const u32 d = (key[k] >> shift) & 255u;
        // uniform digit across the warp (the rule for the high digits of nearly sorted cell keys): rank = lane, no match.any
        const u32 d0 = __shfl_sync(0xffffffffu, d, 0);
        if (__all_sync(0xffffffffu, ok && d == d0)) {
            const u32 prev = s_warp_cnt[wid][d0];
            __syncwarp();
            if (lane == 0) s_warp_cnt[wid][d0] = prev + 32u;
            __syncwarp();
            rank[k] = prev + lane;
            continue;
        }
        u32 grp = __match_any_sync(0xffffffffu, ok ? d : 256u);
        u32 prev = 0;
        if (ok) prev = s_warp_cnt[wid][d];
        __syncwarp();
        if (ok && (grp & lt_mask) == 0) s_warp_cnt[wid][d] = prev + __popc(grp);  // lowest lane of the group
        __syncwarp();
        rank[k] = prev + __popc(grp & lt_mask);
    }
    __syncthreads();

    // ---- thread d owns digit d: exclusive scan over warps, tile count, look-back ----
    u32 cnt = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        u32 c = s_warp_cnt[w][tid];
        s_warp_cnt[w][tid] = cnt;
        cnt += c;
    }
    u32 *my_status = status + (size_t)tile * 256 + tid;
    u32 excl = 0;
    if (tile == 0) {
        atomicExch(my_status, kFlagPrefix | cnt);
    } else {
        atomicExch(my_status, kFlagAgg | cnt);
        // decoupled look-back: walk predecessors until one has published an inclusive prefix.  The tiles in flight publish their
        // aggregates at about the same time, so the front of inclusive prefixes advances by at most kLook tiles per L2 round trip:
        // a pass over 245 tiles (1M pairs) is that chain, not bandwidth.  kLook status words are requested per round trip
        // (8; 16 and 32 measured slower — 66 / 71 / 76 us per 1M-pair sort, 0.35 / 0.39 / 0.44 ms at 8M: the registers they cost take a resident CTA away, profiles/r2k).
        constexpr int kLook = PS_SORT_LOOK;
        int p = (int)tile - 1;
        bool done = false;
        while (!done) {
            u32 sv[kLook];
#pragma unroll
            for (int q = 0; q < kLook; q++)
                sv[q] = (p - q >= 0) ? *((volatile u32 *)(status + (size_t)(p - q) * 256 + tid)) : (2u << 30);  // before tile 0: an inclusive prefix of 0 (kFlagPrefix)
            int used = 0;
#pragma unroll
            for (int q = 0; q < kLook; q++) {
                if (done || used != q) continue;       // stopped at an unpublished word: re-read from there
                if ((sv[q] >> 30) == 0) continue;      // not published yet (tiles are ticketed in order, so it is running)
                excl += sv[q] & kValMask;
                used = q + 1;
                if (sv[q] & kFlagPrefix) done = true;
            }
            p -= used;
        }
        atomicExch(my_status, kFlagPrefix | (excl + cnt));
    }
    // digit base = exclusive scan of the global histogram over digits; tile offset = exclusive scan of cnt
    u32 hcount = hist[tid];
    // block-wide exclusive scans of (hcount, cnt) over the 256 digits
    u32 h_incl = hcount, c_incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 th = __shfl_up_sync(0xffffffffu, h_incl, o), tc = __shfl_up_sync(0xffffffffu, c_incl, o);
        if (lane >= o) { h_incl += th; c_incl += tc; }
    }
    __shared__ u32 s_hsum[kWarps], s_csum[kWarps];
    if (lane == 31) { s_hsum[wid] = h_incl; s_csum[wid] = c_incl; }
    __syncthreads();
    u32 hbase = 0, cbase = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        if (w < wid) { hbase += s_hsum[w]; cbase += s_csum[w]; }
    }
    const u32 digit_base = hbase + h_incl - hcount;  // #keys with a smaller digit (whole array)
    const u32 tile_off = cbase + c_incl - cnt;       // #keys with a smaller digit (this tile)
    s_tile_off[tid] = tile_off;
    s_gofs[tid] = (int)(digit_base + excl) - (int)tile_off;
    __syncthreads();

    // ---- stage digit-sorted in shared memory ----
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        u32 loc = wbase + k * 32 + lane;
        if (loc < nt) {
            u32 d = (key[k] >> shift) & 255u;
            u32 pos = s_tile_off[d] + s_warp_cnt[wid][d] + rank[k];
            s_keys[pos] = key[k];
            s_vals[pos] = val[k];
        }
    }
    __syncthreads();

    // ---- stream out: consecutive threads write consecutive addresses inside each digit run ----
#pragma unroll 4
    for (u32 i = tid; i < nt; i += kThreads) {
        u32 kk = s_keys[i];
        u32 d = (kk >> shift) & 255u;
        u32 dst = (u32)((int)i + s_gofs[d]);
        keys_out[dst] = kk;
        vals_out[dst] = s_vals[i];
    }
}
}  // namespace

int ps_sort_passes(u32 num_cells) {
    int bits = 0;
    while (bits < 32 && ((u32)1 << bits) < num_cells) bits++;
    if (num_cells > (1u << 31)) bits = 32;
    int p = (bits + 7) / 8;
    return p < 1 ? 1 : p;
}

constexpr size_t kSortHeaderElems = 4 * 256 + 8;  // [4][256] histograms, 4 tickets (+4 pad)
size_t ps_sort_status_elems(u32 n, int passes) {
    size_t tiles = ((size_t)n + kTile - 1) / kTile;
    return kSortHeaderElems + tiles * 256 * (size_t)passes;
}
SortScratch ps_sort_scratch_layout(u32 *base) {
    SortScratch sc;
    sc.hist = base;
    sc.ticket = base + 4 * 256;
    sc.status = base + kSortHeaderElems;
    return sc;
}

// zeroes the histograms, the tickets and the look-back status words of one sort (one allocation, ps_sort_scratch_layout: one memset node)
void ps_launch_sort_prepare(u32 n, int passes, SortScratch sc, cudaStream_t s) {
    if (!n) return;
    const u32 tiles = (n + kTile - 1) / kTile;
    cudaMemsetAsync(sc.hist, 0, (kSortHeaderElems + (size_t)tiles * 256 * passes) * sizeof(u32), s);
}

// hist_ready: ps_launch_sort_prepare has run and the digit histograms of kA are already in sc.hist (ps_launch_calc_hash_hist)
void ps_launch_sort(u32 *kA, u32 *vA, u32 *kB, u32 *vB, u32 n, int passes, bool identity_vals, SortScratch sc, cudaStream_t s, bool hist_ready) {
    if (!n) return;
    const u32 tiles = (n + kTile - 1) / kTile;
    if (!hist_ready) {
        ps_launch_sort_prepare(n, passes, sc, s);
        u32 hist_blocks = (n + kThreads * 4 * 4 - 1) / (kThreads * 4 * 4);
        if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
        k_radix_hist<<<hist_blocks, kThreads, 0, s>>>(kA, n, passes, sc.hist);
    }
    u32 *kin = kA, *vin = vA, *kout = kB, *vout = vB;
    for (int p = 0; p < passes; p++) {
        u32 *st = sc.status + (size_t)p * tiles * 256;
        if (p == 0 && identity_vals)
            k_radix_pass<true><<<tiles, kThreads, 0, s>>>(kin, vin, kout, vout, n, 8 * p, sc.hist + 256 * p, st, sc.ticket + p);
        else
            k_radix_pass<false><<<tiles, kThreads, 0, s>>>(kin, vin, kout, vout, n, 8 * p, sc.hist + 256 * p, st, sc.ticket + p);
        u32 *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
}
