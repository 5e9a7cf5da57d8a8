// particlesolver_b200/csrc/ps_context.h — internal definition of the opaque PsCtx (not installed).
#pragma once
#include <cuda_runtime.h>
#include <curand.h>
#include <string>
#include <vector>
#include "../../include/psolver.h"
#include "ps_common.cuh"
#include <nvtx3/nvToolsExt.h>

// NVTX range for the scope (shows up in ncu / nsys timelines; a no-op without a tool attached)
struct PsNvtxRange {
    explicit PsNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~PsNvtxRange() { nvtxRangePop(); }
};

// ps_stream_io.cu: staging frames, copy streams and events of ps_step_streamed
struct PsStreamIo {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    float4 *in[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}, *out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [frame][pos | vel]
    cudaEvent_t in_ready[2] = {nullptr, nullptr}, in_free[2] = {nullptr, nullptr}, out_ready[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    uint64_t cap = 0, calls = 0;
    bool begun = false;  // between ps_io_begin and ps_io_end
    bool prefetched[2] = {false, false}, frames_used[2] = {false, false};
    uint64_t prefetch_n[2] = {0, 0};
    const float *prefetch_src[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
};

struct PsComm;  // ps_comm.cu: NCCL communicator, slab geometry and record buffers of a multi-GPU context

struct PsCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    PsParams params{};
    GridDesc grid{};
    StencilDesc stencil{};
    WorldDesc world{};
    u32 num_cells = 0;
    int sort_passes = 0;

    uint64_t limit = 0;     // hard particle limit (ps_create's max_particles)
    uint64_t capacity = 0;  // allocated slots
    u32 n = 0;              // owned + ghosts
    u32 n_ghost = 0;
    // phase census of the appended particles, so that a step skips the contact pass of an all-fluid scene and the
    // fluid passes of a scene without fluid; unknown after a raw phase upload (then nothing is skipped)
    bool census_known = true;
    u32 n_fluid = 0, n_contact = 0, n_gas = 0;  // phase == FLUID, phase >= CLOTH, phase == GAS
    // contact-phase (>= CLOTH) particles ever handed to this context by its host, +1 per raw phase upload (unknown contents): unlike the
    // census above it survives the slab operations, so a decomposed run can agree once, globally, that no rank holds a contact particle
    uint64_t contact_sources = 0;
    uint64_t nonfluid_sources = 0;   // the same for phase != FLUID

    // per-particle state (SoA).  pos may be caller-owned in the reference-ABI shim, hence the indirection.
    float4 *pos = nullptr, *vel = nullptr, *prev = nullptr, *spos = nullptr;
    float *w = nullptr, *ros = nullptr, *sw = nullptr, *lambda = nullptr;
    int *phase = nullptr, *sphase = nullptr;
    u32 *hash = nullptr, *index = nullptr, *hash_tmp = nullptr, *index_tmp = nullptr, *num_neighbors = nullptr, *occ = nullptr;
    // neighbour lists between K6 and K7 (ps_neighbor_kernels.cu): row pool, per-warp records (rows used or overflow mark, chunk
    // ids; the pool's bump allocator sits 4 words before the first record), number of chunks in the pool
    u32 *nbr_list = nullptr, *nbr_rows = nullptr;
    u32 nbr_max_rows = 0;
    // per-cell state
    u32 *cell_begin = nullptr, *chunk_lb = nullptr;
    u32 *cell_start = nullptr, *cell_end = nullptr;  // the reference's table format: allocated and filled on demand
    bool ref_tables_valid = false;
    uint64_t cell_capacity = 0;
    // sort scratch
    u32 *sort_status = nullptr;  // histograms | tickets | look-back status words (ps_sort_scratch_layout)
    size_t sort_status_elems = 0;
    // wall jitter uniforms, [iterations][6]
    float *rands = nullptr;
    u32 rands_iters = 0;
    curandGenerator_t gen = nullptr;
    uint64_t rand_calls = 0;  // curandGenerateUniform(gen, ., 6) calls so far (a checkpoint replays them to reposition the stream)

    // constraints: host mirror + device arrays
    std::vector<u32> h_dist_idx;    // 2 per constraint
    std::vector<float> h_dist_rest;
    std::vector<u32> h_point_idx;
    std::vector<float> h_point_xyz;  // 3 per pin
    std::vector<u32> h_occ;
    bool constraints_dirty = false;
    u32 num_constrained = 0, num_points = 0;
    u32 *csr_particle = nullptr, *csr_off = nullptr, *csr_other = nullptr, *d_point_idx = nullptr;
    float *csr_rest = nullptr, *d_point_xyz = nullptr;
    u32 *adj_off = nullptr, *adj = nullptr;  // distance-constraint adjacency by particle index (n + 1 offsets, 2m partners); null when m == 0
    float4 *dist_scratch = nullptr;
    bool dist_prefix_ok = true;  // constrained particles are the index prefix [0,K) (see K9 note)

    // rigid bodies (shape matching, K12): CSR over particle indices, rest offsets (xyz, mass), rotation quaternion per body
    std::vector<u32> h_body_off{0u}, h_body_idx;
    std::vector<float> h_body_rest, h_body_stiff;  // 4 per member; 1 per body
    u32 num_bodies = 0, bodies_uploaded = 0;
    u32 *body_off = nullptr, *body_idx = nullptr;
    float4 *body_rest = nullptr, *body_quat = nullptr;
    float *body_stiff = nullptr;
    // SDF contacts between rigid bodies (ps_set_rigid_body_sdf): per member (unit outward gradient in the rest frame, depth), w < 0 = none;
    // sdf_world = the same by particle index in the world frame, refreshed before every contact pass
    std::vector<float> h_body_sdf;  // 4 per member
    bool has_sdf = false, sdf_dirty = false;
    float4 *body_sdf = nullptr, *sdf_world = nullptr;
    u32 *member_body = nullptr;
    uint64_t sdf_capacity = 0;
    // XSPH viscosity / vorticity confinement (K13): coefficients (0 = off) and scratch (omega | dv, float4[2 * capacity])
    float xsph_c = 0.f, vorticity_eps = 0.f;
    float4 *visc_scratch = nullptr;
    uint64_t visc_scratch_cap = 0;

    // CUDA graph of one whole step
    cudaGraphExec_t graph_exec = nullptr;
    // everything issue_step bakes into the captured graph: sizes, parameters, which passes run (the phase census decides whether the
    // contact pass and the fluid passes are issued at all) and the ghost-lambda range of a slab context
    struct GraphKey {
        u32 n, n_ghost, m, p, iters, flags; float dt, omega; u32 bodies; float xsph, vort; u32 passes; float lam_lo, lam_hi;
        bool operator==(const GraphKey &o) const {
            return n == o.n && n_ghost == o.n_ghost && m == o.m && p == o.p && iters == o.iters && flags == o.flags && dt == o.dt && omega == o.omega &&
                   bodies == o.bodies && xsph == o.xsph && vort == o.vort && passes == o.passes && lam_lo == o.lam_lo && lam_hi == o.lam_hi;
        }
    } graph_key{};
    u32 launches_per_step = 0;
    u32 launch_counter = 0;  // counts launches while a step is being issued
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
    bool grid_valid = false;
    // slab decomposition
    u32 *slab_scratch = nullptr;
    size_t slab_scratch_elems = 0;
    u32 *slab_counts_host = nullptr;  // pinned, 2 words
    u32 *slab_ranks = nullptr;        // uint2 per owned particle: its record's position in the last halo pack's two buffers (lambda exchange)
    uint64_t slab_ranks_cap = 0;
    u32 slab_halo_counts[2] = {0, 0}; // records of the last halo pack
    bool slab_ranks_valid = false;
    // ps_slab_set_lambda_sinks: the outgoing lambda messages K6 fills itself; lam_sinks_written: the last lambda pass did
    float *lam_sink[2] = {nullptr, nullptr};
    uint64_t lam_sink_cap = 0;
    bool lam_sinks_written = false;
    float slab_left_below = 0.f, slab_right_from = 0.f;  // membership thresholds of the last halo pack
    bool slab_used = false;           // a slab call has compacted / appended particles: index-based constraints and bodies are refused from then on
    float lambda_xmin = -3.0e38f, lambda_xmax = 3.0e38f;
    PsStreamIo io;
    PsComm *comm = nullptr;
};

// internal helpers shared with the reference-ABI shim
int ps_create_internal(int device, const PsParams *params, uint64_t max_particles, bool legacy_default_stream, PsCtx **out);
int ps_ctx_ensure_capacity(PsCtx *c, uint64_t want);
int ps_ctx_alloc_lists(PsCtx *c, uint64_t cap);
const int *ps_ctx_gas_phase(PsCtx *c);
int ps_ctx_ensure_cells(PsCtx *c);
int ps_ctx_emit_reference_tables(PsCtx *c);
int ps_ctx_sync_constraints(PsCtx *c);
void ps_ctx_refresh_descs(PsCtx *c);
void ps_set_error(const char *fmt, ...);
SortScratch ps_ctx_sort_scratch(PsCtx *c, u32 n);

// ps_extensions.cu: rigid bodies and viscosity (issued from the step when present / enabled); each returns #launches
int ps_ext_sync_bodies(PsCtx *c);
int ps_ext_prepare_step(PsCtx *c);
u32 ps_ext_issue_shapes(PsCtx *c);
u32 ps_ext_issue_sdf(PsCtx *c);  // world-frame SDF for the next contact pass (0 launches when no body carries one)
u32 ps_ext_issue_viscosity(PsCtx *c, float dt);
void ps_ext_free(PsCtx *c);
void ps_io_free(PsCtx *c);  // ps_stream_io.cu
void ps_comm_free(PsCtx *c);  // ps_comm.cu

// stage issue functions on explicit arrays (used by both ABIs); each returns the number of launches issued
u32 ps_issue_build_grid(PsCtx *c, const float4 *pos);
