// particlesolver_b200/csrc/ps_grid_kernels.cu — uniform-grid build around the radix sort:
//   K2  cell hash                         (reference calcHashD, integration_kernel.cuh:206-225)
//   K4  reorder + cell ranges             (reference reorderDataAndFindCellStartD, integration_kernel.cuh:229-299)
//       The cell ranges are kept as ONE dense lower-bound table cell_begin[c] = #particles with key < c (it lets
//       the neighbour kernels turn the reference's (2r+1)^3 cellStart/cellEnd probes into (2r+1)^2 contiguous row
//       ranges); the reference's cellStart/cellEnd pair is derived from it on demand, bit for bit.
// Integer outputs are the bit-exact contract with the reference.
#include "ps_common.cuh"

namespace {
constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_calc_hash(u32 *__restrict__ hash, u32 *__restrict__ index, const float4 *__restrict__ pos, u32 n,
                                                      GridDesc g) {
    u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    float4 p = ld_stream4(pos + i);
    hash[i] = ps_grid_hash(g, ps_grid_pos(g, p.x, p.y, p.z));
    if (index) index[i] = i;  // the fused grid build sorts with implicit identity values and skips this write
}

// K2 fused with the radix sort's histogram kernel: the keys are in registers when they are written, so the 256-bin histograms of
// all digit places (what k_radix_hist builds from a second read of the keys, ps_sort_kernels.cu) are accumulated here — one
// launch and one 4 B/particle read less per grid build.  Shared-memory bins per CTA, a warp whose 32 keys share a digit adds 32
// with one atomic (cell keys of consecutive particles share their high digits), non-empty bins flushed once per CTA.
constexpr int kHashItems = 4;
__global__ void __launch_bounds__(kBlock) k_calc_hash_hist(u32 *__restrict__ hash, const float4 *__restrict__ pos, u32 n, GridDesc g, int passes,
                                                           u32 *__restrict__ hist) {
    __shared__ u32 sh[4 * 256];
    for (int i = threadIdx.x; i < passes * 256; i += kBlock) sh[i] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    for (u32 base = blockIdx.x * (kBlock * kHashItems); base < n; base += gridDim.x * (kBlock * kHashItems)) {
        u32 key[kHashItems];
        bool ok[kHashItems];
#pragma unroll
        for (int k = 0; k < kHashItems; k++) {
            const u32 i = base + k * kBlock + threadIdx.x;
            ok[k] = i < n;
            if (ok[k]) {
                const float4 p = ld_stream4(pos + i);
                key[k] = ps_grid_hash(g, ps_grid_pos(g, p.x, p.y, p.z));
                hash[i] = key[k];
            } else {
                key[k] = 0xffffffffu;
            }
        }
#pragma unroll
        for (int k = 0; k < kHashItems; k++) {
            const bool full = __all_sync(0xffffffffu, ok[k]);  // (the loop bounds are CTA-uniform: every warp arrives here whole)
            for (int p = 0; p < passes; p++) {
                const u32 d = (key[k] >> (8 * p)) & 255u;
                const u32 d0 = __shfl_sync(0xffffffffu, d, 0);
                if (full && __all_sync(0xffffffffu, d == d0)) {
                    if (lane == 0) atomicAdd(&sh[p * 256 + d0], 32u);
                } else if (ok[k]) {
                    atomicAdd(&sh[p * 256 + d], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += kBlock) {
        const u32 c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---------------- K4: gather into sorted order + chunk table ----------------
// One thread per sorted slot: sortedPos/W/Phase[i] = pos/W/phase[index[i]] (exact copies, the contract).
// The reference scatters cellStart/cellEnd from this kernel after a memset of the whole table
// (integration.cu:199, integration_kernel.cuh:262-285).  Here the kernel only records, for every CHUNK of
// kCellsPerBlock cells, the lower bound of the chunk's first cell in the sorted key array (chunk_lb, a table
// 2048x smaller than the cell table); k_cell_begin then builds each chunk of the dense table from its own
// slice of the sorted keys with no memset, no global scatter and no cross-chunk carry.  Gaps between
// consecutive keys can span thousands of chunks (empty space), so a gap is filled by the whole warp.
constexpr int kCellsPerThread = 8;
constexpr int kCellsPerBlock = kBlock * kCellsPerThread;  // 2048
constexpr int kChunkShift = 11;
static_assert((1 << kChunkShift) == kCellsPerBlock, "chunk size");

__global__ void __launch_bounds__(kBlock) k_reorder(float4 *__restrict__ spos, float *__restrict__ sw, int *__restrict__ sphase,
                                                    u32 *__restrict__ chunk_lb, const u32 *__restrict__ hash,
                                                    const u32 *__restrict__ index, const float4 *__restrict__ pos,
                                                    const float *__restrict__ w, const int *__restrict__ phase, u32 n, u32 num_chunks,
                                                    int gas_as_fluid, int slot_in_w, float4 *__restrict__ exact_copy) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    const bool ok = i < n;
    const int lane = threadIdx.x & 31;
    u32 k0 = 0, k1 = 0;  // this slot is the lower bound of chunks [k0, k1)
    if (ok) {
        const u32 h = hash[i];
        const u32 src = index[i];
        // issue the three gathers first so their latency overlaps the chunk bookkeeping
        const float4 p = __ldg(pos + src);
        const float wi = __ldg(w + src);
        const int ph = __ldg(phase + src);
        k1 = (h >> kChunkShift) + 1;
        k0 = i > 0 ? (hash[i - 1] >> kChunkShift) + 1 : 0;
        if (exact_copy) st_stream4(exact_copy + i, p);
        // the staged K6 stages candidates by TMA and takes a candidate's sorted slot from .w (nothing else reads spos.w)
        st_stream4_wbits(spos + i, p.x, p.y, p.z, slot_in_w ? i : __float_as_uint(p.w));
        sw[i] = wi;
        // PS_FLAG_GAS: GAS particles take part in the density constraint — the neighbour kernels see them as fluid (the
        // reference's GPU kernels ignore phase 1 altogether; its CPU app solves gas as a PBF fluid, gasconstraint.cpp)
        sphase[i] = (gas_as_fluid && ph == 1) ? PH_FLUID : ph;
    }
    unsigned todo = __ballot_sync(0xffffffffu, k1 > k0);
    while (todo) {
        const int src_lane = __ffs(todo) - 1;
        todo &= todo - 1;
        const u32 a = __shfl_sync(0xffffffffu, k0, src_lane), b = __shfl_sync(0xffffffffu, k1, src_lane);
        const u32 v = __shfl_sync(0xffffffffu, i, src_lane);
        for (u32 k = a + lane; k < b; k += 32) chunk_lb[k] = v;
    }
    // chunks after the last key (and the end sentinel) take n
    const unsigned last = __ballot_sync(0xffffffffu, ok && i == n - 1);
    if (last) {
        const u32 a = __shfl_sync(0xffffffffu, k1, __ffs(last) - 1);
        for (u32 k = a + lane; k <= num_chunks; k += 32) chunk_lb[k] = n;
    }
}

// ---------------- dense table cell_begin[c] = #keys < c, one chunk of 2048 cells per CTA ----------------
// suffix-min across the warp: lane l gets min over lanes >= l
__device__ __forceinline__ u32 warp_suffix_min_incl(u32 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v = min(v, t);
    }
    return v;
}

__global__ void __launch_bounds__(kBlock) k_cell_begin(u32 *__restrict__ cell_begin, const u32 *__restrict__ hash,
                                                       const u32 *__restrict__ chunk_lb, u32 num_cells, u32 n) {
    __shared__ __align__(16) u32 cs[kCellsPerBlock];  // first sorted slot of each cell of the chunk, 0xffffffff = empty
    __shared__ u32 sm[kBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 c0 = blockIdx.x * kCellsPerBlock;
    const u32 base = c0 + threadIdx.x * kCellsPerThread;
    const u32 lb = chunk_lb[blockIdx.x], ub = chunk_lb[blockIdx.x + 1];
    uint4 *out = reinterpret_cast<uint4 *>(cell_begin + base);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cell_begin[num_cells] = n;
    if (lb == ub) {  // no key in this chunk (most chunks of a sparse grid): every cell's lower bound is ub
        out[0] = make_uint4(ub, ub, ub, ub);
        out[1] = make_uint4(ub, ub, ub, ub);
        return;
    }
    uint4 *cs4 = reinterpret_cast<uint4 *>(cs);
    cs4[threadIdx.x * 2] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    cs4[threadIdx.x * 2 + 1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    __syncthreads();
    for (u32 s = lb + threadIdx.x; s < ub; s += kBlock) {
        const u32 h = hash[s];
        if (s == lb || hash[s - 1] != h) cs[h - c0] = s;
    }
    __syncthreads();
    u32 c[kCellsPerThread];
    {
        const uint4 a = cs4[threadIdx.x * 2], b = cs4[threadIdx.x * 2 + 1];
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    }
    // thread-local inclusive suffix-min, then across the warp, then across the CTA; the carry from later chunks is ub
#pragma unroll
    for (int k = kCellsPerThread - 2; k >= 0; k--) c[k] = min(c[k], c[k + 1]);
    const u32 incl = warp_suffix_min_incl(c[0], lane);
    if (lane == 0) sm[wid] = incl;
    __syncthreads();
    u32 later = ub;
    for (int w2 = wid + 1; w2 < kBlock / 32; w2++) later = min(later, sm[w2]);
    u32 excl = __shfl_down_sync(0xffffffffu, incl, 1);
    if (lane == 31) excl = 0xffffffffu;
    excl = min(excl, later);  // min over every cell after this thread's 8
#pragma unroll
    for (int k = 0; k < kCellsPerThread; k++) c[k] = min(c[k], excl);
    out[0] = make_uint4(c[0], c[1], c[2], c[3]);
    out[1] = make_uint4(c[4], c[5], c[6], c[7]);
}

// ---------------- the reference's own table format, derived on demand ----------------
// cellStart[c] = first sorted slot of cell c or 0xffffffff; cellEnd[c] = one past its last slot.  The reference
// never clears cellEnd (integration.cu:199), so its entries for empty cells are unspecified; here they read 0.
__global__ void __launch_bounds__(kBlock) k_emit_reference_tables(u32 *__restrict__ cell_start, u32 *__restrict__ cell_end,
                                                                  const u32 *__restrict__ cell_begin, u32 num_cells) {
    const u32 c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= num_cells) return;
    const u32 b = cell_begin[c], e = cell_begin[c + 1];
    cell_start[c] = e > b ? b : 0xffffffffu;
    cell_end[c] = e > b ? e : 0u;
}
__global__ void __launch_bounds__(kBlock) k_export_sorted_pos(float4 *__restrict__ out, const float4 *__restrict__ spos,
                                                              const float4 *__restrict__ pos, const u32 *__restrict__ index, u32 n) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float4 p = spos[i];
    out[i] = make_float4(p.x, p.y, p.z, pos[index[i]].w);
}
}  // namespace

static inline u32 cdiv(u32 a, u32 b) { return (a + b - 1) / b; }

void ps_launch_calc_hash(u32 *hash, u32 *index, const float4 *pos, u32 n, GridDesc g, cudaStream_t s) {
    if (!n) return;
    k_calc_hash<<<cdiv(n, kBlock), kBlock, 0, s>>>(hash, index, pos, n, g);
}

// K2 + the sort's histograms in one launch; hist = SortScratch::hist, zeroed by ps_launch_sort_prepare on the same stream before
void ps_launch_calc_hash_hist(u32 *hash, const float4 *pos, u32 n, GridDesc g, int passes, u32 *hist, cudaStream_t s) {
    if (!n) return;
    u32 blocks = cdiv(n, kBlock * kHashItems);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_calc_hash_hist<<<blocks, kBlock, 0, s>>>(hash, pos, n, g, passes, hist);
}

size_t ps_chunk_table_elems(u32 num_cells) { return (size_t)cdiv(num_cells, kCellsPerBlock) + 1; }

void ps_launch_reorder(float4 *spos, float *sw, int *sphase, u32 *chunk_lb, const u32 *hash, const u32 *index, const float4 *pos,
                       const float *w, const int *phase, u32 n, u32 num_cells, cudaStream_t s, bool gas_as_fluid, bool slot_in_w, float4 *exact_copy) {
    if (!n) return;
    k_reorder<<<cdiv(n, kBlock), kBlock, 0, s>>>(spos, sw, sphase, chunk_lb, hash, index, pos, w, phase, n, cdiv(num_cells, kCellsPerBlock),
                                                 gas_as_fluid ? 1 : 0, slot_in_w ? 1 : 0, exact_copy);
}

void ps_launch_export_sorted_pos(float4 *out, const float4 *spos, const float4 *pos, const u32 *index, u32 n, cudaStream_t s) {
    if (!n) return;
    k_export_sorted_pos<<<cdiv(n, kBlock), kBlock, 0, s>>>(out, spos, pos, index, n);
}

void ps_launch_cell_begin(u32 *cell_begin, const u32 *hash, const u32 *chunk_lb, u32 n, u32 num_cells, cudaStream_t s) {
    if (!n) return;
    k_cell_begin<<<cdiv(num_cells, kCellsPerBlock), kBlock, 0, s>>>(cell_begin, hash, chunk_lb, num_cells, n);
}

void ps_launch_emit_reference_tables(u32 *cell_start, u32 *cell_end, const u32 *cell_begin, u32 num_cells, cudaStream_t s) {
    k_emit_reference_tables<<<cdiv(num_cells, kBlock), kBlock, 0, s>>>(cell_start, cell_end, cell_begin, num_cells);
}
