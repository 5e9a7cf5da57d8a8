// particlesolver_b200/csrc/ps_context.cu — context, memory, step orchestration and the ps_* C ABI
// (include/psolver.h).  Host-side equivalent of the reference's wrapper layer
// (gpu/src/cuda/{integration,solver,shared_variables,util}.cu) + ParticleSystem::update
// (gpu/src/particlesystem.cpp:144-246), without process-global state: everything lives in a PsCtx.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include "ps_context.h"

// ------------------------------------------------------------------ errors ------------------------------------------------------------------
static thread_local char g_err[512] = "";
void ps_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
#define CU(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            ps_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return PS_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)
#define CR(x)                                                                                         \
    do {                                                                                              \
        curandStatus_t e_ = (x);                                                                      \
        if (e_ != CURAND_STATUS_SUCCESS) {                                                            \
            ps_set_error("%s failed: curand status %d (%s:%d)", #x, (int)e_, __FILE__, __LINE__);     \
            return PS_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)
#define NEED(c)                                  \
    do {                                         \
        if (!(c)) {                              \
            ps_set_error("null context");        \
            return PS_ERR_INVALID;               \
        }                                        \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" const char *ps_last_error(void) { return g_err; }
extern "C" const char *ps_version(void) { return "particlesolver_b200 0.1 (sm_100a)"; }

extern "C" void ps_default_params(PsParams *p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->gravity[0] = 0.f; p->gravity[1] = -9.8f; p->gravity[2] = 0.f;  // particlesystem.cpp:65
    p->global_damping = 1.0f;
    p->particle_radius = 0.25f;                                        // particleapp.cpp:24
    p->grid_size[0] = p->grid_size[1] = p->grid_size[2] = 64;          // particleapp.cpp:25
    p->cell_size[0] = p->cell_size[1] = p->cell_size[2] = 0.5f;        // 2 * radius, particlesystem.cpp:62
    p->min_bounds[0] = -50; p->min_bounds[1] = 0; p->min_bounds[2] = -50;  // particleapp.cpp:38
    p->max_bounds[0] = 50; p->max_bounds[1] = 200; p->max_bounds[2] = 50;
    p->solver_iterations = 5;
    p->omega = 1.0f;
    p->flags = PS_FLAG_NONE;
    // tuning aid (A/B runs of bench.py and the CLI without a rebuild): PS_EXTRA_FLAGS=<bits> is ORed into the default flags
    if (const char *e = getenv("PS_EXTRA_FLAGS")) p->flags |= (uint32_t)strtoul(e, nullptr, 0);
    p->neighbor_list_rows = 512;  // rows of a warp's list region: >= the 500-neighbour cap, so a list always fits (2 KB of address space per particle)
}

static bool is_pow2(u32 v) { return v && !(v & (v - 1)); }

static int validate_params(const PsParams *p) {
    for (int c = 0; c < 3; c++) {
        if (!is_pow2(p->grid_size[c])) { ps_set_error("grid_size[%d]=%u is not a power of two (the hash wraps with '&')", c, p->grid_size[c]); return PS_ERR_INVALID; }
        if (!(p->cell_size[c] > 0.f)) { ps_set_error("cell_size[%d] must be positive", c); return PS_ERR_INVALID; }
        if (p->grid_size[c] >= (1u << 24)) { ps_set_error("grid_size[%d] must stay below 2^24 (reference hash uses __umul24)", c); return PS_ERR_INVALID; }
    }
    uint64_t cells = (uint64_t)p->grid_size[0] * p->grid_size[1] * p->grid_size[2];
    if (cells > (1ull << 31)) { ps_set_error("grid has %llu cells; the 32-bit cell key allows 2^31", (unsigned long long)cells); return PS_ERR_INVALID; }
    if ((uint64_t)p->grid_size[2] * p->grid_size[1] >= (1ull << 24) && p->grid_size[0] > 1) {
        // (z&mask)*gy must fit the 24-bit multiplier of the reference's __umul24 to stay bit-identical
        ps_set_error("grid_size y*z must stay below 2^24 for a hash identical to the reference's __umul24 form");
        return PS_ERR_INVALID;
    }
    int rad = (int)ceilf(PS_H / p->cell_size[0]);
    if (rad > PS_MAX_RAD) { ps_set_error("cell_size %.4g gives a fluid stencil radius %d > %d", p->cell_size[0], rad, PS_MAX_RAD); return PS_ERR_INVALID; }
    for (int c = 0; c < 3; c++)
        if (p->grid_size[c] < (u32)(2 * rad + 2)) { ps_set_error("grid_size[%d]=%u is smaller than the fluid stencil (%d cells)", c, p->grid_size[c], 2 * rad + 2); return PS_ERR_INVALID; }
    if (p->solver_iterations > 64) { ps_set_error("solver_iterations > 64"); return PS_ERR_INVALID; }
    if (p->neighbor_list_rows % 8 || p->neighbor_list_rows > 4096) { ps_set_error("neighbor_list_rows must be a multiple of 8, at most 4096"); return PS_ERR_INVALID; }
    return PS_OK;
}

void ps_ctx_refresh_descs(PsCtx *c) {
    const PsParams &p = c->params;
    GridDesc &g = c->grid;
    g.ox = p.world_origin[0]; g.oy = p.world_origin[1]; g.oz = p.world_origin[2];
    g.cx = p.cell_size[0]; g.cy = p.cell_size[1]; g.cz = p.cell_size[2];
    g.gx = p.grid_size[0]; g.gy = p.grid_size[1]; g.gz = p.grid_size[2];
    g.mx = g.gx - 1; g.my = g.gy - 1; g.mz = g.gz - 1;
    g.num_cells = g.gx * g.gy * g.gz;
    c->num_cells = g.num_cells;
    c->sort_passes = ps_sort_passes(g.num_cells);
    c->world.radius = p.particle_radius;
    c->world.min_x = p.min_bounds[0]; c->world.min_y = p.min_bounds[1]; c->world.min_z = p.min_bounds[2];
    c->world.max_x = p.max_bounds[0]; c->world.max_y = p.max_bounds[1]; c->world.max_z = p.max_bounds[2];
    // Row stencil of the fluid kernels.  rad follows the reference (integration_kernel.cuh:546).  A cell at offset
    // d along an axis is at least (|d|-1)*cell away, so a (dy,dz) row with ay^2+az^2 > H^2 holds no neighbour
    // and within a kept row only |dx| <= 1 + sqrt(H^2-ay^2-az^2)/cell can.  Margins keep borderline cells.
    StencilDesc &st = c->stencil;
    memset(&st, 0, sizeof st);
    st.rad = (int)ceilf(PS_H / g.cx);
    const int w = 2 * st.rad + 1;
    for (int dz = -st.rad; dz <= st.rad; dz++)
        for (int dy = -st.rad; dy <= st.rad; dy++) {
            double ay = std::max(abs(dy) - 1, 0) * (double)g.cy, az = std::max(abs(dz) - 1, 0) * (double)g.cz;
            double rem = (double)PS_H2 * (1.0 + 1e-4) - ay * ay - az * az;
            int xr = -1;
            if (rem >= 0.0) xr = std::min(st.rad, 1 + (int)floor(sqrt(rem) / g.cx + 1e-4));
            st.xr[(dz + st.rad) * w + (dy + st.rad)] = (signed char)xr;
            if (xr >= 0) st.rowmask[dz + st.rad] |= 1u << (dy + st.rad);
        }
}

// ------------------------------------------------------------------ memory ------------------------------------------------------------------
template <class T>
static int grow(T **p, uint64_t old_n, uint64_t new_n, cudaStream_t s, bool zero_new) {
    T *np = nullptr;
    CU(cudaMalloc((void **)&np, std::max<uint64_t>(new_n, 1) * sizeof(T)));
    if (*p && old_n) CU(cudaMemcpyAsync(np, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (zero_new && new_n > old_n) CU(cudaMemsetAsync(np + old_n, 0, (new_n - old_n) * sizeof(T), s));
    if (*p) { CU(cudaStreamSynchronize(s)); CU(cudaFree(*p)); }
    *p = np;
    return PS_OK;
}

// neighbour-list row pool + per-warp records for `cap` particles (ps_neighbor_kernels.cu); params.neighbor_list_rows rows per
// warp on average, 0 = no lists
int ps_ctx_alloc_lists(PsCtx *c, uint64_t cap) {
    if (c->nbr_list) { CU(cudaFree(c->nbr_list)); c->nbr_list = nullptr; }
    if (c->nbr_rows) { CU(cudaFree(c->nbr_rows - 4)); c->nbr_rows = nullptr; }
    c->nbr_max_rows = 0;
    if (!c->params.neighbor_list_rows || !cap) return PS_OK;
    c->nbr_max_rows = c->params.neighbor_list_rows;  // rows of a warp's list region
    if (!c->nbr_max_rows) return PS_OK;
    u32 *rec = nullptr;
    CU(cudaMalloc((void **)&c->nbr_list, ps_neighbor_list_elems(cap, c->params.neighbor_list_rows) * sizeof(u32)));
    CU(cudaMalloc((void **)&rec, ps_neighbor_record_elems(cap) * sizeof(u32)));
    CU(cudaMemsetAsync(rec, 0, ps_neighbor_record_elems(cap) * sizeof(u32), c->stream));
    c->nbr_rows = rec + 4;
    return PS_OK;
}

int ps_ctx_ensure_capacity(PsCtx *c, uint64_t want) {
    if (want <= c->capacity) return PS_OK;
    uint64_t cap = std::max<uint64_t>(want, c->capacity ? c->capacity * 2 : 1024);
    if (c->limit && cap > c->limit) cap = std::max<uint64_t>(want, c->limit);
    cap = (cap + 3) & ~3ull;
    const uint64_t o = c->capacity;
    cudaStream_t s = c->stream;
    int r;
#define G(arr, zero) if ((r = grow(&c->arr, o, cap, s, zero)) != PS_OK) return r;
    G(pos, true) G(vel, true) G(prev, true) G(spos, true)
    G(w, true) G(ros, true) G(sw, true) G(lambda, true)   // lambda zero-initialised like the reference's resize (integration.cu:68)
    G(phase, true) G(sphase, true)
    G(hash, true) G(index, true) G(hash_tmp, true) G(index_tmp, true) G(num_neighbors, true) G(occ, true)
#undef G
    c->capacity = cap;
    if ((r = ps_ctx_alloc_lists(c, cap)) != PS_OK) return r;
    // sort look-back status words for the largest n this capacity allows
    size_t need = ps_sort_status_elems((u32)cap, 4);
    if (need > c->sort_status_elems) {
        if (c->sort_status) CU(cudaFree(c->sort_status));
        CU(cudaMalloc((void **)&c->sort_status, need * sizeof(u32)));
        c->sort_status_elems = need;
    }
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    return PS_OK;
}

int ps_ctx_ensure_cells(PsCtx *c) {
    uint64_t cells = c->num_cells;
    if (cells <= c->cell_capacity && c->cell_begin) return PS_OK;
    if (c->cell_begin) { CU(cudaFree(c->cell_begin)); CU(cudaFree(c->chunk_lb)); c->cell_begin = c->chunk_lb = nullptr; }
    if (c->cell_start) { CU(cudaFree(c->cell_start)); CU(cudaFree(c->cell_end)); c->cell_start = c->cell_end = nullptr; }
    CU(cudaMalloc((void **)&c->cell_begin, (cells + 8) * sizeof(u32)));
    CU(cudaMalloc((void **)&c->chunk_lb, ps_chunk_table_elems((u32)cells) * sizeof(u32)));
    c->cell_capacity = cells;
    c->ref_tables_valid = false;
    return PS_OK;
}

// cellStart / cellEnd in the reference's format (m_dCellStart / m_dCellEnd), derived from the dense table of the
// last grid build.  Not on the step's path: only parity checks and foreign consumers ask for them.
int ps_ctx_emit_reference_tables(PsCtx *c) {
    if (c->ref_tables_valid) return PS_OK;
    if (!c->grid_valid) { ps_set_error("cellStart/cellEnd requested before a grid build"); return PS_ERR_STATE; }
    const uint64_t cells = c->num_cells;
    if (!c->cell_start) {
        CU(cudaMalloc((void **)&c->cell_start, cells * sizeof(u32)));
        CU(cudaMalloc((void **)&c->cell_end, cells * sizeof(u32)));
    }
    ps_launch_emit_reference_tables(c->cell_start, c->cell_end, c->cell_begin, (u32)cells, c->stream);
    CU(cudaGetLastError());
    c->ref_tables_valid = true;
    return PS_OK;
}

SortScratch ps_ctx_sort_scratch(PsCtx *c, u32) {
    SortScratch sc;
    sc = ps_sort_scratch_layout(c->sort_status);
    return sc;
}

// ------------------------------------------------------------------ lifetime ------------------------------------------------------------------
extern "C" int ps_create(int device, const PsParams *params, uint64_t max_particles, PsCtx **out) {
    return ps_create_internal(device, params, max_particles, false, out);
}

// legacy_default_stream: the reference-ABI shim issues everything on stream 0 so that the caller's own blocking
// cudaMemcpy calls keep the ordering they have with the reference's default-stream wrappers.
int ps_create_internal(int device, const PsParams *params, uint64_t max_particles, bool legacy_default_stream, PsCtx **out) {
    if (!params || !out) { ps_set_error("ps_create: null argument"); return PS_ERR_INVALID; }
    *out = nullptr;
    int r = validate_params(params);
    if (r != PS_OK) return r;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { ps_set_error("ps_create: device %d of %d", device, ndev); return PS_ERR_INVALID; }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { ps_set_error("ps_create: device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor); return PS_ERR_CUDA; }
    if (max_particles >= (1ull << 31)) { ps_set_error("ps_create: max_particles must be < 2^31 per context"); return PS_ERR_INVALID; }
    const uint64_t initial = max_particles ? max_particles : 1024;  // 0 = unlimited, grows on demand (reference-ABI shim)
    DeviceGuard dg(device);
    PsCtx *c = new PsCtx();
    c->device = device;
    c->params = *params;
    c->limit = max_particles;
    ps_ctx_refresh_descs(c);
    auto fail = [&](int code) { ps_destroy(c); return code; };
    c->own_stream = !legacy_default_stream;
    if (c->own_stream && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { ps_set_error("cudaStreamCreate failed"); return fail(PS_ERR_CUDA); }
    if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) { ps_set_error("cudaEventCreate failed"); return fail(PS_ERR_CUDA); }
    c->rands_iters = 64;
    if (cudaMalloc((void **)&c->rands, c->rands_iters * 6 * sizeof(float)) != cudaSuccess) { ps_set_error("cudaMalloc failed"); return fail(PS_ERR_CUDA); }
    cudaMemsetAsync(c->rands, 0, c->rands_iters * 6 * sizeof(float), c->stream);
    // same generator the reference creates in initIntegration (integration.cu:46-47): XORWOW, seed 1234
    if (curandCreateGenerator(&c->gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS ||
        curandSetPseudoRandomGeneratorSeed(c->gen, 1234ULL) != CURAND_STATUS_SUCCESS || curandSetStream(c->gen, c->stream) != CURAND_STATUS_SUCCESS) {
        ps_set_error("curand generator setup failed"); return fail(PS_ERR_CUDA);
    }
    if ((r = ps_ctx_ensure_cells(c)) != PS_OK) return fail(r);
    if ((r = ps_ctx_ensure_capacity(c, initial)) != PS_OK) return fail(r);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { ps_set_error("stream sync failed"); return fail(PS_ERR_CUDA); }
    *out = c;
    return PS_OK;
}

extern "C" int ps_destroy(PsCtx *c) {
    if (!c) return PS_OK;
    DeviceGuard dg(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->gen) curandDestroyGenerator(c->gen);
    void *ptrs[] = {c->pos, c->vel, c->prev, c->spos, c->w, c->ros, c->sw, c->lambda, c->phase, c->sphase, c->hash, c->index, c->hash_tmp,
                    c->index_tmp, c->num_neighbors, c->occ, c->cell_start, c->cell_end, c->cell_begin, c->chunk_lb,
                    c->sort_status, c->rands, c->slab_scratch, c->slab_ranks, c->nbr_list, c->nbr_rows ? c->nbr_rows - 4 : nullptr, c->csr_particle, c->csr_off, c->csr_other, c->d_point_idx, c->csr_rest,
                    c->d_point_xyz, c->dist_scratch, c->adj_off, c->adj};
    for (void *p : ptrs) if (p) cudaFree(p);
    ps_comm_free(c);
    ps_ext_free(c);
    ps_io_free(c);
    if (c->slab_counts_host) cudaFreeHost(c->slab_counts_host);
    if (c->tm0) { cudaEventDestroy(c->tm0); cudaEventDestroy(c->tm1); }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return PS_OK;
}

extern "C" int ps_set_params(PsCtx *c, const PsParams *p) {
    NEED(c);
    if (!p) { ps_set_error("ps_set_params: null params"); return PS_ERR_INVALID; }
    int r = validate_params(p);
    if (r != PS_OK) return r;
    DeviceGuard dg(c->device);
    const bool rows_changed = p->neighbor_list_rows != c->params.neighbor_list_rows;
    c->params = *p;
    ps_ctx_refresh_descs(c);
    if ((r = ps_ctx_ensure_cells(c)) != PS_OK) return r;
    if (rows_changed) {
        CU(cudaStreamSynchronize(c->stream));
        c->params.neighbor_list_rows = p->neighbor_list_rows;
        int lr = ps_ctx_alloc_lists(c, c->capacity);
        if (lr != PS_OK) return lr;
    }
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    c->grid_valid = false;
    return PS_OK;
}
extern "C" int ps_get_params(PsCtx *c, PsParams *out) { NEED(c); if (!out) return PS_ERR_INVALID; *out = c->params; return PS_OK; }
extern "C" uint64_t ps_num_particles(PsCtx *c) { return c ? c->n : 0; }
extern "C" uint64_t ps_num_owned(PsCtx *c) { return c ? c->n - c->n_ghost : 0; }
extern "C" uint64_t ps_num_cells(PsCtx *c) { return c ? c->num_cells : 0; }
extern "C" uint32_t ps_launches_per_step(PsCtx *c) { return c ? c->launches_per_step : 0; }
extern "C" void *ps_stream(PsCtx *c) { return c ? (void *)c->stream : nullptr; }

// ------------------------------------------------------------------ scene building ------------------------------------------------------------------
extern "C" int ps_append_particles(PsCtx *c, const float *pos4, const float *vel4, const float *inv_mass, const float *rest_density,
                                   const int32_t *phase, uint64_t n) {
    NEED(c);
    if (n == 0) return PS_OK;
    if (!pos4 || !vel4 || !inv_mass || !rest_density || !phase) { ps_set_error("ps_append_particles: null array"); return PS_ERR_INVALID; }
    if (c->n_ghost) { ps_set_error("ps_append_particles: drop ghosts first (ps_set_ghost_count(ctx,0))"); return PS_ERR_STATE; }
    if (c->limit && (uint64_t)c->n + n > c->limit) {
        ps_set_error("ps_append_particles: %llu + %llu exceeds max_particles %llu", (unsigned long long)c->n, (unsigned long long)n, (unsigned long long)c->limit);
        return PS_ERR_CAPACITY;
    }
    DeviceGuard dg(c->device);
    int r = ps_ctx_ensure_capacity(c, (uint64_t)c->n + n);
    if (r != PS_OK) return r;
    cudaStream_t s = c->stream;
    CU(cudaMemcpyAsync(c->pos + c->n, pos4, n * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->vel + c->n, vel4, n * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->w + c->n, inv_mass, n * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->ros + c->n, rest_density, n * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->phase + c->n, phase, n * 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));  // host buffers may be stack arrays of the caller (the reference's builders are)
    for (uint64_t k = 0; k < n; k++) { c->n_fluid += phase[k] == PH_FLUID; c->n_gas += phase[k] == 1; c->n_contact += phase[k] >= PH_CLOTH; c->contact_sources += phase[k] >= PH_CLOTH; c->nonfluid_sources += phase[k] != PH_FLUID; }
    c->n += (u32)n;
    c->h_occ.resize(c->n, 0u);
    c->constraints_dirty = true;
    c->grid_valid = false;
    return PS_OK;
}

extern "C" int ps_add_distance_constraints(PsCtx *c, const uint32_t *idx, const float *rest, uint64_t m) {
    NEED(c);
    if (m == 0) return PS_OK;
    if (!idx || !rest) { ps_set_error("ps_add_distance_constraints: null array"); return PS_ERR_INVALID; }
    if (c->n_ghost || c->slab_used) { ps_set_error("ps_add_distance_constraints: index-based constraints are not supported on slab contexts"); return PS_ERR_STATE; }
    for (uint64_t k = 0; k < 2 * m; k++)
        if (idx[k] >= c->n) { ps_set_error("distance constraint endpoint %u >= %u particles", idx[k], c->n); return PS_ERR_INVALID; }
    c->h_dist_idx.insert(c->h_dist_idx.end(), idx, idx + 2 * m);
    c->h_dist_rest.insert(c->h_dist_rest.end(), rest, rest + m);
    for (uint64_t k = 0; k < 2 * m; k++) c->h_occ[idx[k]]++;  // updateOccurences, solver.cu:72-106,149
    c->constraints_dirty = true;
    return PS_OK;
}

extern "C" int ps_add_point_constraints(PsCtx *c, const uint32_t *idx, const float *xyz, uint64_t p) {
    NEED(c);
    if (p == 0) return PS_OK;
    if (!idx || !xyz) { ps_set_error("ps_add_point_constraints: null array"); return PS_ERR_INVALID; }
    if (c->n_ghost || c->slab_used) { ps_set_error("ps_add_point_constraints: index-based constraints are not supported on slab contexts"); return PS_ERR_STATE; }
    for (uint64_t k = 0; k < p; k++)
        if (idx[k] >= c->n) { ps_set_error("point constraint index %u >= %u particles", idx[k], c->n); return PS_ERR_INVALID; }
    c->h_point_idx.insert(c->h_point_idx.end(), idx, idx + p);
    c->h_point_xyz.insert(c->h_point_xyz.end(), xyz, xyz + 3 * p);
    for (uint64_t k = 0; k < p; k++) c->h_occ[idx[k]]++;  // solver.cu:122
    c->constraints_dirty = true;
    return PS_OK;
}

extern "C" uint64_t ps_num_distance_constraints(PsCtx *c) { return c ? c->h_dist_rest.size() : 0; }
extern "C" uint64_t ps_num_point_constraints(PsCtx *c) { return c ? c->h_point_idx.size() : 0; }
extern "C" int ps_copy_distance_constraints(PsCtx *c, uint32_t *idx, float *rest) {
    NEED(c);
    if (idx) memcpy(idx, c->h_dist_idx.data(), c->h_dist_idx.size() * sizeof(u32));
    if (rest) memcpy(rest, c->h_dist_rest.data(), c->h_dist_rest.size() * sizeof(float));
    return PS_OK;
}
extern "C" int ps_copy_point_constraints(PsCtx *c, uint32_t *idx, float *xyz) {
    NEED(c);
    if (idx) memcpy(idx, c->h_point_idx.data(), c->h_point_idx.size() * sizeof(u32));
    if (xyz) memcpy(xyz, c->h_point_xyz.data(), c->h_point_xyz.size() * sizeof(float));
    return PS_OK;
}

template <class T>
static int upload_vec(T **d, const std::vector<T> &h, cudaStream_t s) {
    if (*d) { CU(cudaFree(*d)); *d = nullptr; }
    if (h.empty()) return PS_OK;
    CU(cudaMalloc((void **)d, h.size() * sizeof(T)));
    CU(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return PS_OK;
}

// Build the gather CSR of the distance constraints (see K9 note in ps_stream_kernels.cu) and upload pins / counts.
int ps_ctx_sync_constraints(PsCtx *c) {
    if (!c->constraints_dirty) return PS_OK;
    cudaStream_t s = c->stream;
    CU(cudaStreamSynchronize(s));
    const size_t m = c->h_dist_rest.size();
    std::vector<u32> deg(c->n, 0u);
    for (size_t k = 0; k < 2 * m; k++) deg[c->h_dist_idx[k]]++;
    std::vector<u32> particle, off, other;
    std::vector<float> rest;
    std::vector<u32> slot(c->n, 0xffffffffu);
    for (u32 i = 0; i < c->n; i++)
        if (deg[i]) { slot[i] = (u32)particle.size(); particle.push_back(i); }
    const u32 K = (u32)particle.size();
    c->dist_prefix_ok = true;
    for (u32 k = 0; k < K; k++) if (particle[k] != k) { c->dist_prefix_ok = false; break; }
    off.assign(K + 1, 0u);
    for (u32 k = 0; k < K; k++) off[k + 1] = off[k] + deg[particle[k]];
    other.assign(2 * m, 0u);
    rest.assign(2 * m, 0.f);
    std::vector<u32> fill(off.begin(), off.end() - (K ? 1 : 0));
    if (K == 0) fill.clear();
    // order inside a particle's list == order of its entries after the reference's stable sort_by_key of
    // [all first endpoints..., all second endpoints...] (solver.cu:203-219): "a" roles first, then "b" roles
    for (size_t k = 0; k < m; k++) { u32 a = c->h_dist_idx[2 * k], b = c->h_dist_idx[2 * k + 1]; u32 t = fill[slot[a]]++; other[t] = b; rest[t] = c->h_dist_rest[k]; }
    for (size_t k = 0; k < m; k++) { u32 a = c->h_dist_idx[2 * k], b = c->h_dist_idx[2 * k + 1]; u32 t = fill[slot[b]]++; other[t] = a | 0x80000000u; rest[t] = c->h_dist_rest[k]; }
    int r;
    if ((r = upload_vec(&c->csr_particle, particle, s)) != PS_OK) return r;
    if ((r = upload_vec(&c->csr_off, off, s)) != PS_OK) return r;
    if ((r = upload_vec(&c->csr_other, other, s)) != PS_OK) return r;
    if ((r = upload_vec(&c->csr_rest, rest, s)) != PS_OK) return r;
    // adjacency of the distance constraints by particle index, for the opt-in self-collision of K5 (PS_FLAG_SELF_COLLISION);
    // fluid-only scenes (m == 0) carry none
    std::vector<u32> adj_off, adj;
    if (m) {
        adj_off.assign((size_t)c->n + 1, 0u);
        for (u32 i = 0; i < c->n; i++) adj_off[i + 1] = adj_off[i] + deg[i];
        adj.resize(2 * m);
        for (u32 k = 0; k < K; k++)
            for (u32 t = off[k]; t < off[k + 1]; t++) adj[adj_off[particle[k]] + (t - off[k])] = other[t] & 0x7fffffffu;
    }
    if ((r = upload_vec(&c->adj_off, adj_off, s)) != PS_OK) return r;
    if ((r = upload_vec(&c->adj, adj, s)) != PS_OK) return r;
    if ((r = upload_vec(&c->d_point_idx, c->h_point_idx, s)) != PS_OK) return r;
    if ((r = upload_vec(&c->d_point_xyz, c->h_point_xyz, s)) != PS_OK) return r;
    if (c->dist_scratch) { CU(cudaFree(c->dist_scratch)); c->dist_scratch = nullptr; }
    if (K) CU(cudaMalloc((void **)&c->dist_scratch, (size_t)K * sizeof(float4)));
    if (c->n) CU(cudaMemcpyAsync(c->occ, c->h_occ.data(), (size_t)c->n * sizeof(u32), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    c->num_constrained = K;
    c->num_points = (u32)c->h_point_idx.size();
    c->constraints_dirty = false;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    return PS_OK;
}

// ------------------------------------------------------------------ stages ------------------------------------------------------------------
// K2 -> K3 -> K4 (+ dense table).  When the pass count is odd the unsorted keys are written straight into the
// scratch pair so that the last pass lands in (hash,index); values are implicit (identity) in the first pass.
u32 ps_issue_build_grid(PsCtx *c, const float4 *pos) {
    cudaStream_t s = c->stream;
    const u32 n = c->n;
    const bool odd = (c->sort_passes & 1) != 0;
    u32 *kA = odd ? c->hash_tmp : c->hash, *vA = odd ? c->index_tmp : c->index;
    u32 *kB = odd ? c->hash : c->hash_tmp, *vB = odd ? c->index : c->index_tmp;
    const SortScratch sc = ps_ctx_sort_scratch(c, n);
    ps_launch_sort_prepare(n, c->sort_passes, sc, s);
    ps_launch_calc_hash_hist(kA, pos, n, c->grid, c->sort_passes, sc.hist, s);  // K2 + the sort's digit histograms
    ps_launch_sort(kA, vA, kB, vB, n, c->sort_passes, true, sc, s, true);
    ps_launch_reorder(c->spos, c->sw, c->sphase, c->chunk_lb, c->hash, c->index, pos, c->w, c->phase, n, c->num_cells, s, (c->params.flags & PS_FLAG_GAS) != 0, true);
    ps_launch_cell_begin(c->cell_begin, c->hash, c->chunk_lb, n, c->num_cells, s);
    c->grid_valid = true;
    c->ref_tables_valid = false;
    // kernels only (the memset node of the sort is not counted): calc_hash + histograms 1, passes, reorder 1, cell_begin 1
    return 1 + (u32)c->sort_passes + 1 + 1;
}

// PS_FLAG_GAS and the scene holds (or may hold) GAS particles: the prediction reads the phases for their buoyancy
const int *ps_ctx_gas_phase(PsCtx *c) {
    return ((c->params.flags & PS_FLAG_GAS) && (!c->census_known || c->n_gas > 0)) ? c->phase : nullptr;
}

static int ready(PsCtx *c) {
    NEED(c);
    if (c->n == 0) return PS_OK;
    int r = ps_ctx_sync_constraints(c);
    if (r == PS_OK) r = ps_ext_sync_bodies(c);
    if (r == PS_OK) r = ps_ext_prepare_step(c);
    return r;
}
static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ps_set_error("%s: %s", what, cudaGetErrorString(e)); return PS_ERR_CUDA; }
    return PS_OK;
}

extern "C" int ps_begin_step(PsCtx *c) {
    NEED(c);
    DeviceGuard dg(c->device);
    const u32 iters = c->params.solver_iterations;
    // one curandGenerateUniform(gen, rands, 6) per solver iteration, exactly like collideWorld (integration.cu:326)
    for (u32 it = 0; it < iters; it++) CR(curandGenerateUniform(c->gen, c->rands + 6 * it, 6));
    c->rand_calls += iters;
    return PS_OK;
}

extern "C" int ps_predict(PsCtx *c, float dt) {
    PsNvtxRange nvtx("ps_predict");
    int r = ready(c); if (r != PS_OK) return r;
    DeviceGuard dg(c->device);
    dt = std::min(dt, .05f);
    const PsParams &p = c->params;
    ps_launch_predict(c->pos, c->vel, c->prev, c->n - c->n_ghost, dt, make_float3(p.gravity[0], p.gravity[1], p.gravity[2]), c->stream, ps_ctx_gas_phase(c));
    return check_launch("ps_predict");
}
extern "C" int ps_build_grid(PsCtx *c) {
    PsNvtxRange nvtx("ps_build_grid");
    int r = ready(c); if (r != PS_OK) return r;
    DeviceGuard dg(c->device);
    if (c->n) ps_issue_build_grid(c, c->pos);
    return check_launch("ps_build_grid");
}
// K5's same-phase rule: nullptr = the reference's (no self-collision); the constraint adjacency under PS_FLAG_SELF_COLLISION
static inline const u32 *self_collision_adj(const PsCtx *c) { return (c->params.flags & PS_FLAG_SELF_COLLISION) ? c->adj_off : nullptr; }

extern "C" int ps_solve_contacts(PsCtx *c) {
    PsNvtxRange nvtx("ps_solve_contacts");
    int r = ready(c); if (r != PS_OK) return r;
    if (c->n && !c->grid_valid) { ps_set_error("ps_solve_contacts: no grid (call ps_build_grid)"); return PS_ERR_STATE; }
    DeviceGuard dg(c->device);
    ps_ext_issue_sdf(c);
    ps_launch_collide(c->pos, c->prev, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->num_neighbors, c->n, c->n - c->n_ghost, c->grid,
                      c->params.particle_radius, c->params.omega, self_collision_adj(c), c->adj, c->has_sdf ? c->sdf_world : nullptr, c->stream);
    return check_launch("ps_solve_contacts");
}
static int issue_fluid(PsCtx *c, const char *what, bool do_lambda, bool do_delta) {
    PsNvtxRange nvtx(what);
    int r = ready(c); if (r != PS_OK) return r;
    if (c->n && !c->grid_valid) { ps_set_error("%s: no grid (call ps_build_grid)", what); return PS_ERR_STATE; }
    DeviceGuard dg(c->device);
    // lambda is needed for the ghosts next to a face too (their owners are on another GPU): computed here for the ghosts inside
    // ps_slab_set_lambda_range, or received from the owners between the two halves (ps_slab_pack_lambda / ps_slab_set_ghost_lambda)
    if (do_lambda) {
        LambdaSinks sinks;
        if (c->lam_sink[0] && c->slab_ranks_valid && c->slab_halo_counts[0] <= c->lam_sink_cap && c->slab_halo_counts[1] <= c->lam_sink_cap) {
            sinks.ranks = reinterpret_cast<const uint2 *>(c->slab_ranks); sinks.left = c->lam_sink[0]; sinks.right = c->lam_sink[1];
            sinks.cap = (u32)std::min<uint64_t>(c->lam_sink_cap, 0xfffffffeu);
            sinks.left_below = c->slab_left_below; sinks.right_from = c->slab_right_from;
        }
        ps_launch_find_lambdas(c->lambda, c->num_neighbors, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->ros, c->n, c->n - c->n_ghost,
                               c->lambda_xmin, c->lambda_xmax, c->grid, c->stencil, (c->params.flags & PS_FLAG_ZERO_NONFLUID_LAMBDA) != 0, c->nbr_list,
                               c->nbr_rows, c->nbr_max_rows, c->capacity, true, (c->params.flags & PS_FLAG_STAGED_LAMBDA) != 0, c->device, c->stream, sinks,
                               &c->lam_sinks_written);
    }
    if (do_delta)
        ps_launch_solve_fluids(c->pos, c->lambda, c->spos, c->sphase, c->index, c->cell_begin, c->ros, c->n, c->n - c->n_ghost, c->grid, c->stencil,
                               c->params.omega, c->nbr_list, c->nbr_rows, c->nbr_max_rows, c->num_neighbors, c->device, c->stream);
    return check_launch(what);
}
extern "C" int ps_solve_fluid(PsCtx *c) { return issue_fluid(c, "ps_solve_fluid", true, true); }
extern "C" int ps_solve_fluid_lambda(PsCtx *c) { return issue_fluid(c, "ps_solve_fluid_lambda", true, false); }
extern "C" int ps_solve_fluid_delta(PsCtx *c) { return issue_fluid(c, "ps_solve_fluid_delta", false, true); }
extern "C" int ps_collide_world(PsCtx *c, uint32_t iteration) {
    PsNvtxRange nvtx("ps_collide_world");
    int r = ready(c); if (r != PS_OK) return r;
    if (iteration >= c->rands_iters) { ps_set_error("ps_collide_world: iteration %u out of range", iteration); return PS_ERR_INVALID; }
    DeviceGuard dg(c->device);
    ps_launch_collide_world(c->pos, c->prev, c->phase, c->n - c->n_ghost, c->rands + 6 * iteration, c->world, c->stream);
    return check_launch("ps_collide_world");
}
extern "C" int ps_solve_distance(PsCtx *c) {
    PsNvtxRange nvtx("ps_solve_distance");
    int r = ready(c); if (r != PS_OK) return r;
    DeviceGuard dg(c->device);
    ps_launch_distance(c->pos, c->dist_scratch, c->csr_particle, c->csr_off, c->csr_other, c->csr_rest, c->occ, c->num_constrained,
                       c->params.omega, c->stream);
    return check_launch("ps_solve_distance");
}
extern "C" int ps_solve_point(PsCtx *c) {
    PsNvtxRange nvtx("ps_solve_point");
    int r = ready(c); if (r != PS_OK) return r;
    DeviceGuard dg(c->device);
    ps_launch_point(c->pos, c->d_point_idx, c->d_point_xyz, c->num_points, c->stream);
    return check_launch("ps_solve_point");
}
extern "C" int ps_update_velocity(PsCtx *c, float dt) {
    PsNvtxRange nvtx("ps_update_velocity");
    int r = ready(c); if (r != PS_OK) return r;
    DeviceGuard dg(c->device);
    dt = std::min(dt, .05f);
    ps_launch_velocity(c->pos, c->prev, c->vel, c->n - c->n_ghost, dt, c->stream);
    return check_launch("ps_update_velocity");
}

// issue the whole step on the context's stream (either eagerly or under stream capture); returns #launches
static u32 issue_step(PsCtx *c, float dt) {
    const PsParams &p = c->params;
    cudaStream_t s = c->stream;
    const u32 n = c->n, n_owned = c->n - c->n_ghost;
    u32 launches = 0;
    // the phase census lets an all-fluid scene skip the contact pass and a fluid-free scene the two PBF passes
    // (both kernels would only read every phase and return)
    const bool has_contact = !c->census_known || c->n_contact > 0, has_fluid = !c->census_known || c->n_fluid > 0 || ((c->params.flags & PS_FLAG_GAS) && c->n_gas > 0);
    ps_launch_predict(c->pos, c->vel, c->prev, n_owned, dt, make_float3(p.gravity[0], p.gravity[1], p.gravity[2]), s, ps_ctx_gas_phase(c));
    launches++;
    for (u32 it = 0; it < p.solver_iterations; it++) {
        launches += ps_issue_build_grid(c, c->pos);
        if (has_contact) {
            launches += ps_ext_issue_sdf(c);
            ps_launch_collide(c->pos, c->prev, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->num_neighbors, n, n_owned, c->grid,
                              p.particle_radius, p.omega, self_collision_adj(c), c->adj, c->has_sdf ? c->sdf_world : nullptr, s);
            launches++;
        }
        if (has_fluid) {
            launches += ps_launch_find_lambdas(c->lambda, c->num_neighbors, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->ros, n, n_owned,
                                               c->lambda_xmin, c->lambda_xmax, c->grid, c->stencil, (p.flags & PS_FLAG_ZERO_NONFLUID_LAMBDA) != 0, c->nbr_list,
                                               c->nbr_rows, c->nbr_max_rows, c->capacity, true, (c->params.flags & PS_FLAG_STAGED_LAMBDA) != 0, c->device, s);
            launches += ps_launch_solve_fluids(c->pos, c->lambda, c->spos, c->sphase, c->index, c->cell_begin, c->ros, n, n_owned, c->grid,
                                               c->stencil, p.omega, c->nbr_list, c->nbr_rows, c->nbr_max_rows, c->num_neighbors, c->device, s);
        }
        ps_launch_collide_world(c->pos, c->prev, c->phase, n_owned, c->rands + 6 * it, c->world, s);
        launches++;
        if (c->num_constrained) { ps_launch_distance(c->pos, c->dist_scratch, c->csr_particle, c->csr_off, c->csr_other, c->csr_rest, c->occ, c->num_constrained, p.omega, s); launches += 2; }
        launches += ps_ext_issue_shapes(c);
        if (c->num_points) { ps_launch_point(c->pos, c->d_point_idx, c->d_point_xyz, c->num_points, s); launches++; }
    }
    ps_launch_velocity(c->pos, c->prev, c->vel, n_owned, dt, s);
    launches++;
    launches += ps_ext_issue_viscosity(c, dt);
    return launches;
}

extern "C" int ps_step(PsCtx *c, float dt) {
    PsNvtxRange nvtx("ps_step");
    int r = ready(c); if (r != PS_OK) return r;
    if (c->n == 0) return PS_OK;  // the reference returns early too (particlesystem.cpp:151-155)
    if (c->n_ghost) { ps_set_error("ps_step: slab contexts with ghosts are stepped stage by stage (halo refresh between stages)"); return PS_ERR_STATE; }
    DeviceGuard dg(c->device);
    dt = std::min(dt, .05f);  // particlesystem.cpp:149
    if ((r = ps_begin_step(c)) != PS_OK) return r;
    static const bool no_graph = getenv("PS_NO_GRAPH") != nullptr;
    cudaStream_t s = c->stream;
    CU(cudaEventRecord(c->ev0, s));
    if (no_graph) {
        c->launches_per_step = issue_step(c, dt);
    } else {
        PsCtx::GraphKey key{c->n, c->n_ghost, (u32)c->h_dist_rest.size(), c->num_points, c->params.solver_iterations, c->params.flags, dt, c->params.omega,
                            c->num_bodies, c->xsph_c, c->vorticity_eps,
                            (c->census_known ? 1u : 0u) | (c->n_contact ? 2u : 0u) | (c->n_fluid ? 4u : 0u) | (c->n_gas ? 8u : 0u), c->lambda_xmin, c->lambda_xmax};
        if (!c->graph_exec || !(key == c->graph_key)) {
            if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            u32 launches = issue_step(c, dt);
            cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (e != cudaSuccess) { ps_set_error("stream capture of the step failed: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
            e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { ps_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); c->graph_exec = nullptr; return PS_ERR_CUDA; }
            c->graph_key = key;
            c->launches_per_step = launches;
        }
        CU(cudaGraphLaunch(c->graph_exec, s));
    }
    CU(cudaEventRecord(c->ev1, s));
    return check_launch("ps_step");
}

// ---- instrumented step: same launches as ps_step, issued eagerly with a CUDA event after every stage ----
static const char *const kStageNames[PS_NUM_STAGES] = {"predict", "hash", "sort", "reorder", "cell_table", "contacts", "lambda", "delta_p",
                                                       "world", "distance", "point", "velocity"};
extern "C" const char *ps_stage_name(int stage) { return (stage >= 0 && stage < PS_NUM_STAGES) ? kStageNames[stage] : ""; }

extern "C" int ps_step_profiled(PsCtx *c, float dt, float *stage_ms, uint32_t *stage_launches) {
    int r = ready(c); if (r != PS_OK) return r;
    if (!stage_ms) { ps_set_error("ps_step_profiled: null output"); return PS_ERR_INVALID; }
    for (int k = 0; k < PS_NUM_STAGES; k++) { stage_ms[k] = 0.f; if (stage_launches) stage_launches[k] = 0; }
    if (c->n == 0) return PS_OK;
    if (c->n_ghost) { ps_set_error("ps_step_profiled: not for slab contexts"); return PS_ERR_STATE; }
    DeviceGuard dg(c->device);
    dt = std::min(dt, .05f);
    if ((r = ps_begin_step(c)) != PS_OK) return r;
    const PsParams &p = c->params;
    cudaStream_t s = c->stream;
    const u32 n = c->n;
    const u32 iters = p.solver_iterations;
    const bool has_contact = !c->census_known || c->n_contact > 0, has_fluid = !c->census_known || c->n_fluid > 0 || ((c->params.flags & PS_FLAG_GAS) && c->n_gas > 0);
    const int max_marks = 4 + (int)iters * 17;
    std::vector<cudaEvent_t> ev(max_marks);
    std::vector<int> tag(max_marks, -1);
    for (auto &e : ev) CU(cudaEventCreate(&e));
    int m = 0;
    // one NVTX range per stage: opened after the previous stage's mark, closed at this stage's mark
    nvtxRangePushA("predict");
    auto mark = [&](int stage, u32 launches) {
        cudaEventRecord(ev[m], s); tag[m] = stage; if (stage >= 0 && stage_launches) stage_launches[stage] += launches; m++;
        if (stage >= 0) { nvtxRangePop(); nvtxRangePushA("solver stage"); }
    };
    mark(-1, 0);
    ps_launch_predict(c->pos, c->vel, c->prev, n, dt, make_float3(p.gravity[0], p.gravity[1], p.gravity[2]), s, ps_ctx_gas_phase(c)); mark(0, 1);
    const bool odd = (c->sort_passes & 1) != 0;
    u32 *kA = odd ? c->hash_tmp : c->hash, *vA = odd ? c->index_tmp : c->index;
    u32 *kB = odd ? c->hash : c->hash_tmp, *vB = odd ? c->index : c->index_tmp;
    for (u32 it = 0; it < iters; it++) {
        const SortScratch sc = ps_ctx_sort_scratch(c, n);
        ps_launch_sort_prepare(n, c->sort_passes, sc, s);
        ps_launch_calc_hash_hist(kA, c->pos, n, c->grid, c->sort_passes, sc.hist, s); mark(1, 1);
        ps_launch_sort(kA, vA, kB, vB, n, c->sort_passes, true, sc, s, true); mark(2, c->sort_passes);
        ps_launch_reorder(c->spos, c->sw, c->sphase, c->chunk_lb, c->hash, c->index, c->pos, c->w, c->phase, n, c->num_cells, s, (p.flags & PS_FLAG_GAS) != 0, true); mark(3, 1);
        ps_launch_cell_begin(c->cell_begin, c->hash, c->chunk_lb, n, c->num_cells, s); mark(4, 1);
        c->grid_valid = true;
        c->ref_tables_valid = false;
        if (has_contact) { const u32 lsdf = ps_ext_issue_sdf(c); ps_launch_collide(c->pos, c->prev, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->num_neighbors, n, n, c->grid, p.particle_radius, p.omega, self_collision_adj(c), c->adj, c->has_sdf ? c->sdf_world : nullptr, s); mark(5, 1 + lsdf); }
        if (has_fluid) {
            const u32 k6 = ps_launch_find_lambdas(c->lambda, c->num_neighbors, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->ros, n, n, c->lambda_xmin,
                                                  c->lambda_xmax, c->grid, c->stencil, (p.flags & PS_FLAG_ZERO_NONFLUID_LAMBDA) != 0, c->nbr_list, c->nbr_rows,
                                                  c->nbr_max_rows, c->capacity, true, (c->params.flags & PS_FLAG_STAGED_LAMBDA) != 0, c->device, s); mark(6, k6);
            const u32 k7 = ps_launch_solve_fluids(c->pos, c->lambda, c->spos, c->sphase, c->index, c->cell_begin, c->ros, n, n, c->grid, c->stencil, p.omega,
                                                  c->nbr_list, c->nbr_rows, c->nbr_max_rows, c->num_neighbors, c->device, s); mark(7, k7);
        }
        ps_launch_collide_world(c->pos, c->prev, c->phase, n, c->rands + 6 * it, c->world, s); mark(8, 1);
        if (c->num_constrained) { ps_launch_distance(c->pos, c->dist_scratch, c->csr_particle, c->csr_off, c->csr_other, c->csr_rest, c->occ, c->num_constrained, p.omega, s); mark(9, 2); }
        if (c->num_bodies) { const u32 l = ps_ext_issue_shapes(c); mark(9, l); }
        if (c->num_points) { ps_launch_point(c->pos, c->d_point_idx, c->d_point_xyz, c->num_points, s); mark(10, 1); }
    }
    ps_launch_velocity(c->pos, c->prev, c->vel, n, dt, s); mark(11, 1);
    if (c->xsph_c != 0.f || c->vorticity_eps != 0.f) { const u32 l = ps_ext_issue_viscosity(c, dt); mark(11, l); }
    nvtxRangePop();
    CU(cudaStreamSynchronize(s));
    for (int k = 1; k < m; k++) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ev[k - 1], ev[k]));
        stage_ms[tag[k]] += ms;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return check_launch("ps_step_profiled");
}

extern "C" int ps_timer_start(PsCtx *c) {
    NEED(c);
    DeviceGuard dg(c->device);
    if (!c->tm0) { CU(cudaEventCreate(&c->tm0)); CU(cudaEventCreate(&c->tm1)); }
    CU(cudaEventRecord(c->tm0, c->stream));
    return PS_OK;
}
extern "C" int ps_timer_stop(PsCtx *c, float *ms) {
    NEED(c);
    if (!ms || !c->tm0) { ps_set_error("ps_timer_stop without ps_timer_start"); return PS_ERR_STATE; }
    DeviceGuard dg(c->device);
    CU(cudaEventRecord(c->tm1, c->stream));
    CU(cudaEventSynchronize(c->tm1));
    CU(cudaEventElapsedTime(ms, c->tm0, c->tm1));
    return PS_OK;
}

extern "C" int ps_sync(PsCtx *c) {
    NEED(c);
    DeviceGuard dg(c->device);
    CU(cudaStreamSynchronize(c->stream));
    return PS_OK;
}
extern "C" int ps_last_step_ms(PsCtx *c, float *ms) {
    NEED(c);
    if (!ms) return PS_ERR_INVALID;
    DeviceGuard dg(c->device);
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return PS_OK;
}

// ------------------------------------------------------------------ data access ------------------------------------------------------------------
struct ArrInfo { void *ptr; size_t esz; uint64_t count; };
static ArrInfo arr_info(PsCtx *c, int which) {
    const uint64_t n = c->n, cells = c->num_cells;
    switch (which) {
        case PS_ARR_POS: return {c->pos, 4, 4 * n};
        case PS_ARR_VEL: return {c->vel, 4, 4 * n};
        case PS_ARR_PREV: return {c->prev, 4, 4 * n};
        case PS_ARR_INV_MASS: return {c->w, 4, n};
        case PS_ARR_PHASE: return {c->phase, 4, n};
        case PS_ARR_REST_DENSITY: return {c->ros, 4, n};
        case PS_ARR_HASH: return {c->hash, 4, n};
        case PS_ARR_INDEX: return {c->index, 4, n};
        case PS_ARR_CELL_START: return {c->cell_start, 4, cells};
        case PS_ARR_CELL_END: return {c->cell_end, 4, cells};
        case PS_ARR_SORTED_POS: return {c->spos, 4, 4 * n};
        case PS_ARR_SORTED_INV_MASS: return {c->sw, 4, n};
        case PS_ARR_SORTED_PHASE: return {c->sphase, 4, n};
        case PS_ARR_LAMBDA: return {c->lambda, 4, n};
        case PS_ARR_NUM_NEIGHBORS: return {c->num_neighbors, 4, n};
        case PS_ARR_RANDS: return {c->rands, 4, (uint64_t)c->params.solver_iterations * 6};
        case PS_ARR_OCCURRENCES: return {c->occ, 4, n};
        case PS_ARR_CELL_BEGIN: return {c->cell_begin, 4, cells + 1};
        case PS_ARR_NEIGHBOR_ROWS: return {c->nbr_rows, 4, c->nbr_rows ? ((n + 31) / 32) * PS_LIST_RECORD_WORDS : 0};
        default: return {nullptr, 0, 0};
    }
}
static int copy_common(PsCtx *c, int which, void *host, uint64_t off, uint64_t cnt, bool to_host, bool sync) {
    NEED(c);
    ArrInfo a = arr_info(c, which);
    if (!a.ptr && a.count == 0 && cnt == 0) return PS_OK;
    if (!a.ptr && !(which == PS_ARR_CELL_START || which == PS_ARR_CELL_END)) { ps_set_error("unknown array selector %d", which); return PS_ERR_INVALID; }
    if (!host && cnt) { ps_set_error("null host pointer"); return PS_ERR_INVALID; }
    if (off + cnt > a.count) { ps_set_error("range [%llu,%llu) exceeds array %d of %llu elements", (unsigned long long)off, (unsigned long long)(off + cnt), which, (unsigned long long)a.count); return PS_ERR_INVALID; }
    if (cnt == 0) return PS_OK;
    DeviceGuard dg(c->device);
    if (which == PS_ARR_OCCURRENCES) { int r = ps_ctx_sync_constraints(c); if (r != PS_OK) return r; }
    if (which == PS_ARR_CELL_START || which == PS_ARR_CELL_END) {
        if (!to_host) { ps_set_error("cellStart/cellEnd are derived outputs"); return PS_ERR_INVALID; }
        int r = ps_ctx_emit_reference_tables(c); if (r != PS_OK) return r;
        a = arr_info(c, which);
    }
    if (which == PS_ARR_SORTED_POS) {
        // derived outputs of the grid build; .w of the resident array carries the sorted slot (ps_fluid_staged.cu), the reference's
        // sortedPos carries pos.w: hand out the reference's (a diagnostic path: temporary buffer, synchronous)
        if (!to_host) { ps_set_error("the sorted arrays are derived outputs"); return PS_ERR_INVALID; }
        float4 *tmp = nullptr;
        CU(cudaMalloc((void **)&tmp, (size_t)c->n * sizeof(float4)));
        ps_launch_export_sorted_pos(tmp, c->spos, c->pos, c->index, c->n, c->stream);
        cudaError_t e = cudaMemcpyAsync(host, (char *)tmp + off * a.esz, cnt * a.esz, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        if (e != cudaSuccess) { ps_set_error("download of the sorted positions failed: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
        return PS_OK;
    }
    char *d = (char *)a.ptr + off * a.esz;
    if (to_host) CU(cudaMemcpyAsync(host, d, cnt * a.esz, cudaMemcpyDeviceToHost, c->stream));
    else CU(cudaMemcpyAsync(d, host, cnt * a.esz, cudaMemcpyHostToDevice, c->stream));
    if (sync) CU(cudaStreamSynchronize(c->stream));
    if (!to_host && (which == PS_ARR_POS || which == PS_ARR_INV_MASS || which == PS_ARR_PHASE)) c->grid_valid = false;
    if (!to_host && which == PS_ARR_PHASE) { c->census_known = false; c->contact_sources++; c->nonfluid_sources++; if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; } }
    return PS_OK;
}
extern "C" int ps_download(PsCtx *c, int which, void *host, uint64_t off, uint64_t cnt) { return copy_common(c, which, host, off, cnt, true, true); }
extern "C" int ps_upload(PsCtx *c, int which, const void *host, uint64_t off, uint64_t cnt) { return copy_common(c, which, (void *)host, off, cnt, false, true); }
extern "C" int ps_download_async(PsCtx *c, int which, void *host, uint64_t off, uint64_t cnt) { return copy_common(c, which, host, off, cnt, true, false); }
extern "C" int ps_upload_async(PsCtx *c, int which, const void *host, uint64_t off, uint64_t cnt) { return copy_common(c, which, (void *)host, off, cnt, false, false); }
extern "C" void *ps_device_ptr(PsCtx *c, int which) {
    if (!c) return nullptr;
    if (which == PS_ARR_CELL_START || which == PS_ARR_CELL_END) {  // derived tables: filled for the grid of the last build
        DeviceGuard dg(c->device);
        if (ps_ctx_emit_reference_tables(c) != PS_OK) return nullptr;
    }
    return arr_info(c, which).ptr;
}

// ------------------------------------------------------------------ slabs ------------------------------------------------------------------
extern "C" int ps_set_ghost_count(PsCtx *c, uint64_t ghosts) {
    NEED(c);
    const uint64_t owned = c->n - c->n_ghost;
    if (c->limit && owned + ghosts > c->limit) { ps_set_error("ps_set_ghost_count: %llu owned + %llu ghosts exceeds max_particles", (unsigned long long)owned, (unsigned long long)ghosts); return PS_ERR_CAPACITY; }
    DeviceGuard dg(c->device);
    int r = ps_ctx_ensure_capacity(c, owned + ghosts);
    if (r != PS_OK) return r;
    c->n = (u32)(owned + ghosts);
    c->n_ghost = (u32)ghosts;
    c->h_occ.resize(c->n, 0u);
    c->grid_valid = false;
    return PS_OK;
}

static int slab_ready(PsCtx *c, const char *what) {
    NEED(c);
    if (!c->h_dist_rest.empty() || !c->h_point_idx.empty() || c->num_bodies || !c->h_body_idx.empty()) {
        ps_set_error("%s: slab contexts cannot hold distance / point constraints or rigid bodies (migration changes particle indices)", what);
        return PS_ERR_STATE;
    }
    c->slab_used = true;
    const size_t need = ps_slab_scratch_elems((u32)c->capacity);
    if (need > c->slab_scratch_elems) {
        if (c->slab_scratch) CU(cudaFree(c->slab_scratch));
        CU(cudaMalloc((void **)&c->slab_scratch, need * sizeof(u32)));
        c->slab_scratch_elems = need;
    }
    if (!c->slab_counts_host) CU(cudaMallocHost((void **)&c->slab_counts_host, 2 * sizeof(u32)));
    if (c->slab_ranks_cap < c->capacity) {  // the arrays grew (ghosts arriving): keep the ranks of the last halo pack
        u32 *grown = nullptr;
        CU(cudaMalloc((void **)&grown, 2 * c->capacity * sizeof(u32)));
        if (c->slab_ranks) {
            CU(cudaMemcpyAsync(grown, c->slab_ranks, 2 * c->slab_ranks_cap * sizeof(u32), cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaFree(c->slab_ranks));
        }
        c->slab_ranks = grown;
        c->slab_ranks_cap = c->capacity;
    }
    return PS_OK;
}
static int slab_fetch_counts(PsCtx *c, u32 n, uint32_t counts[2]) {
    const size_t tiles = (ps_slab_scratch_elems(n) - 2) / 2;
    ps_launch_copy_words(c->slab_counts_host, c->slab_scratch + 2 * tiles, 2, c->stream);  // (pinned host memory; not a copy-engine transfer)
    CU(cudaStreamSynchronize(c->stream));
    counts[0] = c->slab_counts_host[0];
    counts[1] = c->slab_counts_host[1];
    return PS_OK;
}

extern "C" int ps_slab_pack_halo(PsCtx *c, float x_lo, float x_hi, float width, void *left_buf, void *right_buf, uint64_t cap, uint32_t counts[2]) {
    int r = slab_ready(c, "ps_slab_pack_halo"); if (r != PS_OK) return r;
    if (!counts || ((!left_buf || !right_buf) && cap)) { ps_set_error("ps_slab_pack_halo: null argument"); return PS_ERR_INVALID; }
    DeviceGuard dg(c->device);
    const u32 owned = c->n - c->n_ghost;
    const float left_below = x_lo + width, right_from = x_hi - width;
    ps_launch_slab_select(c->pos, owned, left_below, right_from, c->slab_scratch, c->stream);
    ps_launch_slab_pack_halo(c->pos, c->w, c->ros, c->phase, owned, left_below, right_from, c->slab_scratch, left_buf, right_buf, (u32)std::min<uint64_t>(cap, 0xffffffffu), c->stream,
                             c->slab_ranks);
    if ((r = check_launch("ps_slab_pack_halo")) != PS_OK) return r;
    if ((r = slab_fetch_counts(c, owned, counts)) != PS_OK) return r;
    c->slab_halo_counts[0] = counts[0];
    c->slab_halo_counts[1] = counts[1];
    c->slab_ranks_valid = true;
    c->slab_left_below = left_below; c->slab_right_from = right_from;
    if (counts[0] > cap || counts[1] > cap) { ps_set_error("ps_slab_pack_halo: %u / %u records exceed the buffer capacity %llu", counts[0], counts[1], (unsigned long long)cap); return PS_ERR_CAPACITY; }
    return PS_OK;
}

extern "C" int ps_slab_set_ghosts(PsCtx *c, const void *from_left, uint64_t n_left, const void *from_right, uint64_t n_right) {
    int r = slab_ready(c, "ps_slab_set_ghosts"); if (r != PS_OK) return r;
    if ((n_left && !from_left) || (n_right && !from_right)) { ps_set_error("ps_slab_set_ghosts: null buffer"); return PS_ERR_INVALID; }
    if ((r = ps_set_ghost_count(c, n_left + n_right)) != PS_OK) return r;
    DeviceGuard dg(c->device);
    ps_launch_slab_unpack_halo(c->pos, c->w, c->ros, c->phase, c->n - c->n_ghost, from_left, (u32)n_left, from_right, (u32)n_right, c->stream);
    c->census_known = false;  // ghosts may carry any phase
    return check_launch("ps_slab_set_ghosts");
}

// lambda exchange between K6 and K7 (ps_solve_fluid_lambda / ps_solve_fluid_delta): one float per record of the last halo pack,
// in the order of those records
extern "C" int ps_slab_pack_lambda(PsCtx *c, void *left_buf, void *right_buf, uint64_t cap, uint32_t counts[2]) {
    int r = slab_ready(c, "ps_slab_pack_lambda"); if (r != PS_OK) return r;
    if (!counts || ((!left_buf || !right_buf) && cap)) { ps_set_error("ps_slab_pack_lambda: null argument"); return PS_ERR_INVALID; }
    if (!c->slab_ranks_valid) { ps_set_error("ps_slab_pack_lambda: no halo pack to answer (call ps_slab_pack_halo first)"); return PS_ERR_STATE; }
    if (c->n && !c->grid_valid) { ps_set_error("ps_slab_pack_lambda: no grid (lambda lives by sorted slot)"); return PS_ERR_STATE; }
    counts[0] = c->slab_halo_counts[0];
    counts[1] = c->slab_halo_counts[1];
    if (counts[0] > cap || counts[1] > cap) { ps_set_error("ps_slab_pack_lambda: %u / %u values exceed the buffer capacity %llu", counts[0], counts[1], (unsigned long long)cap); return PS_ERR_CAPACITY; }
    if (c->lam_sinks_written && left_buf == c->lam_sink[0] && right_buf == c->lam_sink[1]) {  // K6 has filled these very buffers already
        c->lam_sinks_written = false;
        return PS_OK;
    }
    DeviceGuard dg(c->device);
    ps_launch_slab_pack_lambda(c->lambda, c->index, c->slab_ranks, c->n, c->n - c->n_ghost, (float *)left_buf, (float *)right_buf, (u32)std::min<uint64_t>(cap, 0xffffffffu), c->stream);
    return check_launch("ps_slab_pack_lambda");
}
// The lambda messages of ps_slab_pack_lambda as sinks of the lambda pass itself: with them set, ps_solve_fluid_lambda (the default, fused
// K6) writes the lambda of every particle of the last halo pack into left_buf / right_buf as it computes it, and a following
// ps_slab_pack_lambda on the same buffers only reports the counts.  For contexts whose particles are all FLUID (K6 leaves the lambda
// of other phases untouched, and a message must carry those too).  NULL buffers: off.
extern "C" int ps_slab_set_lambda_sinks(PsCtx *c, void *left_buf, void *right_buf, uint64_t cap) {
    NEED(c);
    if ((left_buf == nullptr) != (right_buf == nullptr)) { ps_set_error("ps_slab_set_lambda_sinks: both buffers or none"); return PS_ERR_INVALID; }
    c->lam_sink[0] = (float *)left_buf; c->lam_sink[1] = (float *)right_buf;
    c->lam_sink_cap = left_buf ? cap : 0;
    c->lam_sinks_written = false;
    return PS_OK;
}
extern "C" int ps_slab_set_ghost_lambda(PsCtx *c, const void *from_left, uint64_t n_left, const void *from_right, uint64_t n_right) {
    int r = slab_ready(c, "ps_slab_set_ghost_lambda"); if (r != PS_OK) return r;
    if ((n_left && !from_left) || (n_right && !from_right)) { ps_set_error("ps_slab_set_ghost_lambda: null buffer"); return PS_ERR_INVALID; }
    if (n_left + n_right != c->n_ghost) { ps_set_error("ps_slab_set_ghost_lambda: %llu + %llu values for %u ghosts", (unsigned long long)n_left, (unsigned long long)n_right, c->n_ghost); return PS_ERR_INVALID; }
    if (c->n && !c->grid_valid) { ps_set_error("ps_slab_set_ghost_lambda: no grid (lambda lives by sorted slot)"); return PS_ERR_STATE; }
    DeviceGuard dg(c->device);
    ps_launch_slab_unpack_lambda(c->lambda, c->index, c->n, c->n - c->n_ghost, (const float *)from_left, (u32)n_left, (const float *)from_right, c->stream);
    return check_launch("ps_slab_set_ghost_lambda");
}

extern "C" int ps_slab_pack_migrants(PsCtx *c, float x_lo, float x_hi, void *left_buf, void *right_buf, uint64_t cap, uint32_t counts[2]) {
    int r = slab_ready(c, "ps_slab_pack_migrants"); if (r != PS_OK) return r;
    c->slab_ranks_valid = false;  // particle indices change
    if (!counts || ((!left_buf || !right_buf) && cap)) { ps_set_error("ps_slab_pack_migrants: null argument"); return PS_ERR_INVALID; }
    if (!(x_lo < x_hi)) { ps_set_error("ps_slab_pack_migrants: empty slab [%g, %g)", x_lo, x_hi); return PS_ERR_INVALID; }
    if ((r = ps_set_ghost_count(c, 0)) != PS_OK) return r;
    DeviceGuard dg(c->device);
    cudaStream_t s = c->stream;
    const u32 n = c->n;
    ps_launch_slab_select(c->pos, n, x_lo, x_hi, c->slab_scratch, s);
    if ((r = check_launch("ps_slab_pack_migrants")) != PS_OK) return r;
    if ((r = slab_fetch_counts(c, n, counts)) != PS_OK) return r;
    if (counts[0] > cap || counts[1] > cap) { ps_set_error("ps_slab_pack_migrants: %u / %u records exceed the buffer capacity %llu", counts[0], counts[1], (unsigned long long)cap); return PS_ERR_CAPACITY; }
    const u32 gone = counts[0] + counts[1];
    if (!gone) return PS_OK;
    ps_launch_slab_pack_migrants(c->pos, c->prev, c->vel, c->w, c->ros, c->phase, n, x_lo, x_hi, c->slab_scratch, left_buf, right_buf, (u32)std::min<uint64_t>(cap, 0xffffffffu), s);
    // stable compaction of the stayers, array by array through the sorted-copy scratch (rebuilt by the next grid build);
    // pos classifies every pass, so it goes last
    const size_t keep = n - gone;
    auto c4 = [&](float4 *arr) { ps_launch_slab_compact4(c->pos, arr, c->spos, n, x_lo, x_hi, c->slab_scratch, s); return cudaMemcpyAsync(arr, c->spos, keep * sizeof(float4), cudaMemcpyDeviceToDevice, s); };
    auto c1 = [&](void *arr) { ps_launch_slab_compact1(c->pos, (const u32 *)arr, c->hash_tmp, n, x_lo, x_hi, c->slab_scratch, s); return cudaMemcpyAsync(arr, c->hash_tmp, keep * sizeof(u32), cudaMemcpyDeviceToDevice, s); };
    CU(c4(c->prev)); CU(c4(c->vel)); CU(c1(c->w)); CU(c1(c->ros)); CU(c1(c->phase)); CU(c4(c->pos));
    c->n = (u32)keep;
    c->h_occ.resize(c->n, 0u);
    c->census_known = false;
    c->grid_valid = false;
    return check_launch("ps_slab_pack_migrants");
}

extern "C" int ps_slab_append_migrants(PsCtx *c, const void *from_left, uint64_t n_left, const void *from_right, uint64_t n_right) {
    int r = slab_ready(c, "ps_slab_append_migrants"); if (r != PS_OK) return r;
    if ((n_left && !from_left) || (n_right && !from_right)) { ps_set_error("ps_slab_append_migrants: null buffer"); return PS_ERR_INVALID; }
    if ((r = ps_set_ghost_count(c, 0)) != PS_OK) return r;
    const uint64_t add = n_left + n_right;
    if (!add) return PS_OK;
    if (c->limit && (uint64_t)c->n + add > c->limit) { ps_set_error("ps_slab_append_migrants: %llu + %llu exceeds max_particles", (unsigned long long)c->n, (unsigned long long)add); return PS_ERR_CAPACITY; }
    DeviceGuard dg(c->device);
    if ((r = ps_ctx_ensure_capacity(c, (uint64_t)c->n + add)) != PS_OK) return r;
    ps_launch_slab_append_migrants(c->pos, c->prev, c->vel, c->w, c->ros, c->phase, c->n, from_left, (u32)n_left, from_right, (u32)n_right, c->stream);
    c->n += (u32)add;
    c->h_occ.resize(c->n, 0u);
    c->census_known = false;
    c->grid_valid = false;
    return check_launch("ps_slab_append_migrants");
}

// histogram of the OWNED particles' x (current positions) over `bins` equal bins of [x_min, x_max), out-of-range values in
// the end bins: every rank's histogram summed gives the global distribution the slabs are re-cut from (slab.balanced_cuts)
extern "C" int ps_slab_x_histogram(PsCtx *c, float x_min, float x_max, uint32_t bins, uint64_t *host_counts) {
    NEED(c);
    if (!host_counts || !bins || bins > 65536 || !(x_min < x_max)) { ps_set_error("ps_slab_x_histogram: bad argument"); return PS_ERR_INVALID; }
    DeviceGuard dg(c->device);
    const u32 n_owned = c->n - c->n_ghost;
    const size_t need = bins;
    if (c->slab_scratch_elems < need) {
        if (c->slab_scratch) CU(cudaFree(c->slab_scratch));
        c->slab_scratch = nullptr;
        CU(cudaMalloc((void **)&c->slab_scratch, need * sizeof(u32)));
        c->slab_scratch_elems = need;
    }
    ps_launch_slab_x_histogram(c->pos, n_owned, x_min, x_max, bins, c->slab_scratch, c->stream);
    std::vector<u32> h(bins);
    CU(cudaMemcpyAsync(h.data(), c->slab_scratch, bins * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (u32 b = 0; b < bins; b++) host_counts[b] = h[b];
    return check_launch("ps_slab_x_histogram");
}

extern "C" int ps_slab_set_lambda_range(PsCtx *c, float x_min, float x_max) {
    NEED(c);
    c->lambda_xmin = x_min;
    c->lambda_xmax = x_max;
    return PS_OK;
}
