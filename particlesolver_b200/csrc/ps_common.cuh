// particlesolver_b200/csrc/ps_common.cuh — shared device/host declarations of libpsolver (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpsolver is written for sm_100a (B200) only"
#endif

typedef uint32_t u32;

// phase codes (reference gpu/src/cuda/shared_variables.cuh:4-9)
#define PH_FLUID 0
#define PH_CLOTH 2
#define PH_SOLID 3

// solver constants (reference gpu/src/cuda/integration_kernel.cuh:20-40)
#define PS_EPS 0.001f
#define PS_MAX_NEIGHBORS 500u
#define PS_H 2.f
#define PS_H2 4.f
#define PS_H6 64.f
#define PS_POLY6 0.00305992474f
#define PS_SPIKY 0.22381163872f
#define PS_RELAX .01f
#define PS_K_P .1f
#define PS_DQ_P .2f
#define PS_S_FRICTION .005f
#define PS_K_FRICTION .0002f
#define PS_GAS_ALPHA -.2f  // buoyancy of GAS particles (reference CPU app: ALPHA, cpu/src/simulation.h:21); only with PS_FLAG_GAS

#define PS_MAX_RAD 8
#define PS_LIST_RECORD_WORDS 1  // per-warp neighbour-list status word (ps_fluid_lists.cuh)

// uniform grid descriptor, passed by value (kernel parameter space == constant bank, one per launch, so
// several contexts can coexist; the reference uses a single __constant__ SimParams, integration_kernel.cuh:55)
struct GridDesc {
    float ox, oy, oz;  // worldOrigin
    float cx, cy, cz;  // cellSize
    u32 gx, gy, gz;    // gridSize (powers of two)
    u32 mx, my, mz;    // gridSize-1
    u32 num_cells;
};

// Row stencil: for each (dz,dy) the largest |dx| whose cell can still hold a particle within the support
// radius, or -1 if the whole row is out of reach.  rad = ceil(H/cell) as in integration_kernel.cuh:546.
struct StencilDesc {
    int rad;
    signed char xr[(2 * PS_MAX_RAD + 1) * (2 * PS_MAX_RAD + 1)];
    u32 rowmask[2 * PS_MAX_RAD + 1];  // per dz: bit (dy+rad) set when xr >= 0
};

struct WorldDesc {
    float radius;
    int min_x, min_y, min_z;
    int max_x, max_y, max_z;
};

// ---- calcGridPos / calcGridHash semantics (reference integration_kernel.cuh:187-203) ----
// floor((p-origin)/cell) with the approximate divide the reference's -use_fast_math build emits
// (exact for power-of-two cell sizes), then '&' wrap and x-fastest linearisation.
__device__ __forceinline__ int3 ps_grid_pos(const GridDesc &g, float x, float y, float z) {
    int3 gp;
    gp.x = (int)floorf(__fdividef(x - g.ox, g.cx));
    gp.y = (int)floorf(__fdividef(y - g.oy, g.cy));
    gp.z = (int)floorf(__fdividef(z - g.oz, g.cz));
    return gp;
}
__device__ __forceinline__ u32 ps_grid_hash(const GridDesc &g, int3 gp) {
    u32 x = (u32)gp.x & g.mx, y = (u32)gp.y & g.my, z = (u32)gp.z & g.mz;
    return (z * g.gy + y) * g.gx + x;
}

// streaming (read-once / write-once) 128-bit accesses that do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// the same with the fourth component given as raw bits (a sorted slot travelling in .w must not pass through float arithmetic)
__device__ __forceinline__ void st_stream4_wbits(float4 *p, float x, float y, float z, u32 wbits) {
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)), "r"(__float_as_uint(z)),
                 "r"(wbits) : "memory");
}

// ---------------- launchers (host side, all asynchronous on `s`) ----------------
// ps_stream_kernels.cu
// gas_phase != nullptr (PS_FLAG_GAS): phase array; GAS particles are predicted with gravity x PS_GAS_ALPHA
void ps_launch_predict(float4 *pos, const float4 *vel, float4 *prev, u32 n, float dt, float3 g, cudaStream_t s, const int *gas_phase = nullptr);
void ps_launch_velocity(const float4 *pos, const float4 *prev, float4 *vel, u32 n, float dt, cudaStream_t s);
void ps_launch_collide_world(float4 *pos, const float4 *prev, const int *phase, u32 n, const float *rands6, WorldDesc w, cudaStream_t s);
void ps_launch_point(float4 *pos, const u32 *pidx, const float *pxyz, u32 np, cudaStream_t s);
void ps_launch_distance(float4 *pos, float4 *scratch, const u32 *csr_particle, const u32 *csr_off, const u32 *csr_other,
                        const float *csr_rest, const u32 *occ, u32 num_constrained, float omega, cudaStream_t s);
// ps_grid_kernels.cu
void ps_launch_calc_hash(u32 *hash, u32 *index, const float4 *pos, u32 n, GridDesc g, cudaStream_t s);
// K4: gather into sorted order + per-chunk lower bounds (chunk_lb[ps_chunk_table_elems(num_cells)])
// slot_in_w: spos[i].w = bit pattern of i instead of pos.w (the staged K6 reads a candidate's slot from there; nothing else reads
// spos.w); exact_copy != nullptr: also the reference's sortedPos, exact float4 copies (the reference ABI hands that array out)
void ps_launch_reorder(float4 *spos, float *sw, int *sphase, u32 *chunk_lb, const u32 *hash, const u32 *index, const float4 *pos,
                       const float *w, const int *phase, u32 n, u32 num_cells, cudaStream_t s, bool gas_as_fluid = false, bool slot_in_w = false,
                       float4 *exact_copy = nullptr);
// out[i] = (spos[i].xyz, pos[index[i]].w): the reference's sortedPos from a slot_in_w array (downloads / diagnostics)
void ps_launch_export_sorted_pos(float4 *out, const float4 *spos, const float4 *pos, const u32 *index, u32 n, cudaStream_t s);
size_t ps_chunk_table_elems(u32 num_cells);
// dense lower-bound table cell_begin[c] = #particles with key < c, c in [0,num_cells]; from the sorted keys + chunk_lb
void ps_launch_cell_begin(u32 *cell_begin, const u32 *hash, const u32 *chunk_lb, u32 n, u32 num_cells, cudaStream_t s);
// the reference's cellStart/cellEnd pair, derived from cell_begin (cellEnd of empty cells reads 0)
void ps_launch_emit_reference_tables(u32 *cell_start, u32 *cell_end, const u32 *cell_begin, u32 num_cells, cudaStream_t s);
// ps_sort_kernels.cu
struct SortScratch {
    u32 *hist;    // [4][256] global digit histograms
    u32 *status;  // [passes][tiles][256] decoupled look-back words
    u32 *ticket;  // [4] dynamic tile ids
};
size_t ps_sort_status_elems(u32 n, int passes);  // whole scratch: histograms + tickets + status words
SortScratch ps_sort_scratch_layout(u32 *base);
int ps_sort_passes(u32 num_cells);
// Stable LSD radix sort of (key,val) pairs on the low 8*passes key bits, ping-ponging between (kA,vA) and
// (kB,vB): input in A, result in A when `passes` is even and in B when it is odd (the caller picks where the
// unsorted keys are written so that the result lands where it wants it).  1 + passes launches + 1 memset.
// identity_vals: vals[i]==i on entry is assumed and vA is never read (saves one 4 B/particle read).
void ps_launch_sort(u32 *kA, u32 *vA, u32 *kB, u32 *vB, u32 n, int passes, bool identity_vals, SortScratch sc, cudaStream_t s, bool hist_ready = false);
// the fused form: ps_launch_sort_prepare (zeroes the sort's scratch), ps_launch_calc_hash_hist (K2 + the digit histograms of the keys it
// writes), then ps_launch_sort(..., hist_ready = true)
void ps_launch_sort_prepare(u32 n, int passes, SortScratch sc, cudaStream_t s);
void ps_launch_calc_hash_hist(u32 *hash, const float4 *pos, u32 n, GridDesc g, int passes, u32 *hist, cudaStream_t s);
// ps_neighbor_kernels.cu
void ps_launch_collide(float4 *pos, const float4 *prev, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                       const u32 *cell_begin, u32 *num_neighbors, u32 n, u32 n_owned, GridDesc g, float radius, float omega, const u32 *adj_off,
                       const u32 *adj, const float4 *sdf_world, cudaStream_t s);
// nbr_list / nbr_rows: interleaved per-warp neighbour lists written by K6 and consumed by K7 and the other passes (nullptr: they
// re-walk the grid); format in ps_fluid_lists.cuh.  `list_rows` = rows of a warp's region (PsParams.neighbor_list_rows), `capacity`
// = the particle capacity the buffers were sized for, `nbr_rows` = the first per-warp status word of the buffer sized by
// ps_neighbor_record_elems; `device` = the context's device (per-device shared-memory opt-in).
// Where K6 drops the lambdas of the particles of the last halo pack (slab contexts, ps_slab_set_lambda_sinks): ranks[orig] = the record's
// position in the left / right halo buffer (0xffffffff: none), left / right = the outgoing lambda messages.  ranks == nullptr: off.
struct LambdaSinks {
    const uint2 *ranks = nullptr;
    float *left = nullptr, *right = nullptr;
    u32 cap = 0;
    float left_below = 0.f, right_from = 0.f;  // the pack's membership test (x < left_below | x >= right_from): only members look their rank up
};
size_t ps_neighbor_list_elems(unsigned long long capacity, u32 rows_per_warp);
size_t ps_neighbor_record_elems(unsigned long long capacity);
u32 ps_launch_find_lambdas(float *lambda, u32 *num_neighbors, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                           const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g,
                           const StencilDesc &st, bool zero_nonfluid, u32 *nbr_list, u32 *nbr_rows, u32 list_rows, unsigned long long capacity,
                           bool slot_in_w, bool staged, int device, cudaStream_t s, LambdaSinks sinks = LambdaSinks(),
                           bool *sinks_written = nullptr);  // returns the number of launches; *sinks_written: the pass filled the sinks itself
u32 ps_launch_solve_fluids(float4 *pos, const float *lambda, const float4 *spos, const int *sphase, const u32 *index,
                           const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, GridDesc g, const StencilDesc &st, float omega,
                           const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows, const u32 *num_neighbors, int device,
                           cudaStream_t s);  // returns the number of launches
// K13 (optional, not in the reference): XSPH viscosity + vorticity confinement on the PBF neighbour structure; returns #launches
u32 ps_launch_viscosity(float4 *vel, float4 *scratch, const float4 *spos, const int *sphase, const u32 *index, const u32 *cell_begin, u32 n, GridDesc g,
                        const StencilDesc &st, float c_xsph, float vorticity_eps, float dt, const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows,
                        const u32 *num_neighbors, int device, cudaStream_t s);
u32 ps_launch_density_error(float4 *scratch, const float4 *spos, const float *sw, const int *sphase, const u32 *index, const float *ros, const u32 *cell_begin,
                            u32 n, GridDesc g, const StencilDesc &st, const u32 *nbr_list, const u32 *nbr_rows, u32 list_rows,
                            const u32 *num_neighbors, int device, cudaStream_t s);
// ps_shape_kernels.cu — K12 (not in the reference's GPU solver): shape matching, one warp per rigid body
void ps_launch_sdf_world(float4 *sdf_world, const float4 *sdf_rest, const u32 *body_idx, const u32 *member_body, const float4 *quat, u32 members,
                         cudaStream_t s);
void ps_launch_shape_match(float4 *pos, const u32 *body_off, const u32 *body_idx, const float4 *rest, float4 *quat, const float *stiff, u32 num_bodies,
                           int max_iters, cudaStream_t s);
// ps_slab_kernels.cu — slab decomposition: ordered selection / packing / compaction
size_t ps_slab_scratch_elems(u32 n);
void ps_launch_slab_select(const float4 *pos, u32 n, float left_below, float right_from, u32 *scratch, cudaStream_t s);
void ps_launch_slab_pack_halo(const float4 *pos, const float *w, const float *ros, const int *phase, u32 n, float left_below, float right_from,
                              const u32 *scratch, void *left, void *right, u32 cap, cudaStream_t s, u32 *ranks = nullptr);
// lambda exchange: ranks = uint2[n_owned] written by the halo pack (position of each owned particle in the two halo buffers)
void ps_launch_slab_pack_lambda(const float *lambda, const u32 *index, const u32 *ranks, u32 n, u32 n_owned, float *left, float *right, u32 cap,
                                cudaStream_t s);
void ps_launch_slab_unpack_lambda(float *lambda, const u32 *index, u32 n, u32 n_owned, const float *from_left, u32 n_left, const float *from_right,
                                  cudaStream_t s);
void ps_launch_slab_unpack_halo(float4 *pos, float *w, float *ros, int *phase, u32 first, const void *from_left, u32 n_left, const void *from_right,
                                u32 n_right, cudaStream_t s);
void ps_launch_slab_pack_migrants(const float4 *pos, const float4 *prev, const float4 *vel, const float *w, const float *ros, const int *phase, u32 n,
                                  float left_below, float right_from, const u32 *scratch, void *left, void *right, u32 cap, cudaStream_t s);
void ps_launch_slab_compact4(const float4 *pos, const float4 *src, float4 *dst, u32 n, float left_below, float right_from, const u32 *scratch,
                             cudaStream_t s);
void ps_launch_slab_compact1(const float4 *pos, const u32 *src, u32 *dst, u32 n, float left_below, float right_from, const u32 *scratch,
                             cudaStream_t s);
void ps_launch_slab_append_migrants(float4 *pos, float4 *prev, float4 *vel, float *w, float *ros, int *phase, u32 first, const void *from_left,
                                    u32 n_left, const void *from_right, u32 n_right, cudaStream_t s);
// words between device memory and pinned (device-visible) host memory by a kernel, not by a copy engine
void ps_launch_copy_words(u32 *dst, const u32 *src, u32 n, cudaStream_t s);
void ps_launch_slab_x_histogram(const float4 *pos, u32 n, float x_min, float x_max, u32 bins, u32 *hist, cudaStream_t s);
