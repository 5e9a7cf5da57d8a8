// particlesolver_b200/csrc/ps_grid_kernels.cu — uniform-grid build around the radix sort:
//   K2  cell hash                         (reference calcHashD, integration_kernel.cuh:206-225)
//   K4  reorder + cellStart/cellEnd       (reference reorderDataAndFindCellStartD, integration_kernel.cuh:229-299)
//   +   dense lower-bound table cell_begin[c] = #particles with key < c  (new; lets the neighbour kernels turn the
//       reference's (2r+1)^3 cellStart/cellEnd probes into (2r+1)^2 contiguous row ranges)
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

// One thread per sorted slot.  cellStart has been memset to 0xffffffff; cellEnd is deliberately NOT cleared
// (the reference does not either, integration.cu:199) — it is only meaningful where cellStart != 0xffffffff.
__global__ void __launch_bounds__(kBlock) k_reorder(u32 *__restrict__ cell_start, u32 *__restrict__ cell_end, float4 *__restrict__ spos,
                                                    float *__restrict__ sw, int *__restrict__ sphase, const u32 *__restrict__ hash,
                                                    const u32 *__restrict__ index, const float4 *__restrict__ pos,
                                                    const float *__restrict__ w, const int *__restrict__ phase, u32 n) {
    u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    u32 h = hash[i];
    u32 src = index[i];
    // issue the three gathers first so their latency overlaps the cell-range bookkeeping
    float4 p = __ldg(pos + src);
    float wi = __ldg(w + src);
    int ph = __ldg(phase + src);
    u32 hp = i > 0 ? hash[i - 1] : 0xffffffffu;
    if (i == 0 || h != hp) {
        cell_start[h] = i;
        if (i > 0) cell_end[hp] = i;
    }
    if (i == n - 1) cell_end[h] = n;
    st_stream4(spos + i, p);
    sw[i] = wi;
    sphase[i] = ph;
}

// ---------------- dense table: suffix-min scan of cellStart over the cells ----------------
constexpr int kCellsPerThread = 8;
constexpr int kCellsPerBlock = kBlock * kCellsPerThread;  // 2048

__device__ __forceinline__ u32 warp_min(u32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// suffix-min across the warp: lane l gets min over lanes >= l
__device__ __forceinline__ u32 warp_suffix_min_incl(u32 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v = min(v, t);
    }
    return v;
}

__device__ __forceinline__ void load_cells(const u32 *__restrict__ cell_start, u32 base, u32 num_cells, u32 (&c)[kCellsPerThread]) {
    if (base + kCellsPerThread <= num_cells) {
        uint4 a = __ldg(reinterpret_cast<const uint4 *>(cell_start + base));
        uint4 b = __ldg(reinterpret_cast<const uint4 *>(cell_start + base + 4));
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kCellsPerThread; k++) c[k] = (base + k < num_cells) ? __ldg(cell_start + base + k) : 0xffffffffu;
    }
}

__global__ void __launch_bounds__(kBlock) k_cell_block_min(const u32 *__restrict__ cell_start, u32 *__restrict__ block_min, u32 num_cells) {
    __shared__ u32 sm[kBlock / 32];
    u32 base = blockIdx.x * kCellsPerBlock + threadIdx.x * kCellsPerThread;
    u32 c[kCellsPerThread];
    load_cells(cell_start, base, num_cells, c);
    u32 m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < kCellsPerThread; k++) m = min(m, c[k]);
    m = warp_min(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        u32 v = threadIdx.x < kBlock / 32 ? sm[threadIdx.x] : 0xffffffffu;
        v = warp_min(v);
        if (threadIdx.x == 0) block_min[blockIdx.x] = v;
    }
}

// single CTA: block_min[b] <- min(n, min over b' > b of block_min[b'])   (exclusive suffix-min, in place)
__global__ void __launch_bounds__(1024) k_cell_carry(u32 *__restrict__ block_min, u32 num_blocks, u32 n) {
    __shared__ u32 sm[32];
    __shared__ u32 s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = n;
    __syncthreads();
    u32 chunks = (num_blocks + 1023) / 1024;
    for (u32 ch = chunks; ch-- > 0;) {
        u32 b = ch * 1024 + threadIdx.x;
        u32 v = b < num_blocks ? block_min[b] : 0xffffffffu;
        u32 incl = warp_suffix_min_incl(v, lane);  // min over lanes >= lane
        if (lane == 0) sm[wid] = incl;             // == min of the whole warp
        __syncthreads();
        const u32 carry = s_carry;  // min over every later chunk (n if none)
        u32 later = carry, chunk_min = 0xffffffffu;
        for (int w2 = 0; w2 < 32; w2++) {
            u32 m = sm[w2];
            chunk_min = min(chunk_min, m);
            if (w2 > wid) later = min(later, m);
        }
        u32 excl = __shfl_down_sync(0xffffffffu, incl, 1);
        if (lane == 31) excl = 0xffffffffu;
        excl = min(excl, later);
        __syncthreads();  // every read of s_carry / sm[] is done
        if (b < num_blocks) block_min[b] = excl;
        if (threadIdx.x == 0) s_carry = min(carry, chunk_min);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock) k_cell_begin(u32 *__restrict__ cell_begin, const u32 *__restrict__ cell_start,
                                                       const u32 *__restrict__ block_carry, u32 num_cells, u32 n) {
    __shared__ u32 sm[kBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 base = blockIdx.x * kCellsPerBlock + threadIdx.x * kCellsPerThread;
    u32 c[kCellsPerThread];
    load_cells(cell_start, base, num_cells, c);
    // thread-local inclusive suffix-min
#pragma unroll
    for (int k = kCellsPerThread - 2; k >= 0; k--) c[k] = min(c[k], c[k + 1]);
    u32 incl = warp_suffix_min_incl(c[0], lane);
    if (lane == 0) sm[wid] = incl;
    __syncthreads();
    u32 later = block_carry[blockIdx.x];
    for (int w2 = wid + 1; w2 < kBlock / 32; w2++) later = min(later, sm[w2]);
    u32 excl = __shfl_down_sync(0xffffffffu, incl, 1);
    if (lane == 31) excl = 0xffffffffu;
    excl = min(excl, later);  // min over every cell after this thread's 8
#pragma unroll
    for (int k = 0; k < kCellsPerThread; k++) c[k] = min(c[k], excl);
    if (base + kCellsPerThread <= num_cells) {
        reinterpret_cast<uint4 *>(cell_begin + base)[0] = make_uint4(c[0], c[1], c[2], c[3]);
        reinterpret_cast<uint4 *>(cell_begin + base)[1] = make_uint4(c[4], c[5], c[6], c[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kCellsPerThread; k++)
            if (base + k < num_cells) cell_begin[base + k] = c[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cell_begin[num_cells] = n;
}
}  // namespace

static inline u32 cdiv(u32 a, u32 b) { return (a + b - 1) / b; }

void ps_launch_calc_hash(u32 *hash, u32 *index, const float4 *pos, u32 n, GridDesc g, cudaStream_t s) {
    if (!n) return;
    k_calc_hash<<<cdiv(n, kBlock), kBlock, 0, s>>>(hash, index, pos, n, g);
}

void ps_launch_reorder(u32 *cell_start, u32 *cell_end, float4 *spos, float *sw, int *sphase, const u32 *hash, const u32 *index,
                       const float4 *pos, const float *w, const int *phase, u32 n, u32 num_cells, cudaStream_t s) {
    cudaMemsetAsync(cell_start, 0xff, (size_t)num_cells * sizeof(u32), s);
    if (!n) return;
    k_reorder<<<cdiv(n, kBlock), kBlock, 0, s>>>(cell_start, cell_end, spos, sw, sphase, hash, index, pos, w, phase, n);
}

size_t ps_cell_begin_scratch_elems(u32 num_cells) { return cdiv(num_cells, kCellsPerBlock); }

void ps_launch_cell_begin(u32 *cell_begin, const u32 *cell_start, u32 *block_min, u32 n, u32 num_cells, cudaStream_t s) {
    u32 nb = cdiv(num_cells, kCellsPerBlock);
    k_cell_block_min<<<nb, kBlock, 0, s>>>(cell_start, block_min, num_cells);
    k_cell_carry<<<1, 1024, 0, s>>>(block_min, nb, n);
    k_cell_begin<<<nb, kBlock, 0, s>>>(cell_begin, cell_start, block_min, num_cells, n);
}
