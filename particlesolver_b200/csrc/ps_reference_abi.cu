// particlesolver_b200/csrc/ps_reference_abi.cu — the reference's extern "C" wrapper names on our kernels
// (include/ps_reference_abi.h).  One library-global context on the legacy default stream, caller-owned device
// arrays, exit-on-error: the contract of gpu/src/cuda/{integration,solver,shared_variables}.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "ps_context.h"
#include "../../include/ps_reference_abi.h"

namespace {
PsCtx *g = nullptr;            // the one live system (the reference keeps file-scope thrust vectors)
u32 g_n_integ = 0;             // sizes of the three independently appended groups
u32 g_n_shared = 0;            //   V/ros (integration.cu:50-72), W/phase (shared_variables.cu:32-50),
u32 g_n_solver = 0;            //   occurences (solver.cu:64-70)
const u32 *g_dense_for = nullptr;  // cellStart pointer handed out by the last reorderDataAndFindCellStart

[[noreturn]] void die(const char *where) {
    // reference behaviour: checkCudaErrors prints and exit(EXIT_FAILURE)s (helper_cuda.h:981-1008)
    fprintf(stderr, "libpsolver (reference ABI) %s: %s\n", where, ps_last_error());
    exit(EXIT_FAILURE);
}
void ck(int r, const char *where) { if (r != PS_OK) die(where); }
void ck_cuda(cudaError_t e, const char *where) {
    if (e != cudaSuccess) { ps_set_error("%s", cudaGetErrorString(e)); die(where); }
}
void ck_launch(const char *where) { ck_cuda(cudaGetLastError(), where); }

PsCtx *ctx() {
    if (!g) {
        PsParams p;
        ps_default_params(&p);
        int dev = 0;
        cudaGetDevice(&dev);
        ck(ps_create_internal(dev, &p, 0 /* unlimited, grows */, true /* legacy default stream */, &g), "initIntegration");
        g_n_integ = g_n_shared = g_n_solver = 0;
    }
    return g;
}
void maybe_destroy() {
    // the reference frees its three groups of vectors with three calls (particlesystem.cpp:131-133)
    if (g && g_n_integ == 0 && g_n_shared == 0 && g_n_solver == 0) { ps_destroy(g); g = nullptr; g_dense_for = nullptr; }
}
// collide / solveFluids receive the caller's cellStart/cellEnd; the kernels walk the dense table that
// reorderDataAndFindCellStart built beside them, so the two must come from the same call.
void need_dense(const u32 *cell_start, u32 num_cells, const char *where) {
    PsCtx *c = ctx();
    if (!g_dense_for || g_dense_for != cell_start || num_cells != c->num_cells) {
        ps_set_error("cell tables were not produced by the preceding reorderDataAndFindCellStart call");
        die(where);
    }
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- integration.cu ----------------------------------------------------------------
void initIntegration(void) { ctx(); }

void freeIntegrationVectors(void) { g_n_integ = 0; maybe_destroy(); }

void appendIntegrationParticle(float *v, float *ro, uint n) {
    PsCtx *c = ctx();
    ck(ps_ctx_ensure_capacity(c, (uint64_t)g_n_integ + n), "appendIntegrationParticle");
    ck_cuda(cudaMemcpy(c->vel + g_n_integ, v, (size_t)n * 16, cudaMemcpyHostToDevice), "appendIntegrationParticle");
    ck_cuda(cudaMemcpy(c->ros + g_n_integ, ro, (size_t)n * 4, cudaMemcpyHostToDevice), "appendIntegrationParticle");
    g_n_integ += n;
}

void setParameters(PsRefSimParams *h) {
    PsCtx *c = ctx();
    PsParams p = c->params;
    memcpy(p.gravity, h->gravity, 12);
    p.global_damping = h->globalDamping;
    p.particle_radius = h->particleRadius;
    memcpy(p.grid_size, h->gridSize, 12);
    memcpy(p.world_origin, h->worldOrigin, 12);
    memcpy(p.cell_size, h->cellSize, 12);
    if (memcmp(&p, &c->params, sizeof p) == 0) return;  // the reference re-uploads every frame (particlesystem.cpp:163)
    ck(ps_set_params(c, &p), "setParameters");
    g_dense_for = nullptr;
}

void integrateSystem(float *pos, float dt, uint n) {
    PsCtx *c = ctx();
    const PsParams &p = c->params;
    // copyToXstar + integrate_functor in one pass (integration.cu:122-135)
    ps_launch_predict((float4 *)pos, c->vel, c->prev, n, dt, make_float3(p.gravity[0], p.gravity[1], p.gravity[2]), c->stream);
    ck_launch("integrateSystem");
}

void calcHash(uint *hash, uint *index, float *pos, int n) {
    PsCtx *c = ctx();
    ps_launch_calc_hash(hash, index, (const float4 *)pos, (u32)n, c->grid, c->stream);
    ck_launch("calcHash");
}

void sortParticles(uint *hash, uint *index, uint n) {
    PsCtx *c = ctx();
    ck(ps_ctx_ensure_capacity(c, n), "sortParticles");
    ps_launch_sort(hash, index, c->hash_tmp, c->index_tmp, n, c->sort_passes, false, ps_ctx_sort_scratch(c, n), c->stream);
    if (c->sort_passes & 1) {  // in-place contract of thrust::sort_by_key
        ck_cuda(cudaMemcpyAsync(hash, c->hash_tmp, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream), "sortParticles");
        ck_cuda(cudaMemcpyAsync(index, c->index_tmp, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream), "sortParticles");
    }
    ck_launch("sortParticles");
}

void reorderDataAndFindCellStart(uint *cellStart, uint *cellEnd, float *sortedPos, float *sortedW, int *sortedPhase, uint *hash,
                                 uint *index, float *oldPos, uint n, uint numCells) {
    PsCtx *c = ctx();
    if (numCells != c->num_cells) { ps_set_error("numCells %u does not match setParameters' grid (%u)", numCells, c->num_cells); die("reorderDataAndFindCellStart"); }
    // the caller's sortedPos receives the reference's exact float4 copies; the fluid kernels work from the context's own sorted
    // positions, whose .w carries the sorted slot (what the staged K6 wants, ps_fluid_staged.cu)
    ck(ps_ctx_ensure_capacity(c, n), "reorderDataAndFindCellStart");
    ps_launch_reorder(c->spos, sortedW, sortedPhase, c->chunk_lb, hash, index, (const float4 *)oldPos, c->w, c->phase, n, numCells,
                      c->stream, false, true, (float4 *)sortedPos);
    if (n) {
        ps_launch_cell_begin(c->cell_begin, hash, c->chunk_lb, n, numCells, c->stream);
        // the caller's tables in the reference's format (cellEnd of empty cells: 0; the reference leaves them stale)
        ps_launch_emit_reference_tables(cellStart, cellEnd, c->cell_begin, numCells, c->stream);
    } else {
        ck_cuda(cudaMemsetAsync(cellStart, 0xff, (size_t)numCells * sizeof(uint), c->stream), "reorderDataAndFindCellStart");
    }
    ck_launch("reorderDataAndFindCellStart");
    g_dense_for = cellStart;
}

void collide(float *particles, float *sortedPos, float *sortedW, int *sortedPhase, uint *index, uint *cellStart, uint *cellEnd, uint n,
             uint numCells) {
    (void)cellEnd;
    PsCtx *c = ctx();
    need_dense(cellStart, numCells, "collide");
    ps_launch_collide((float4 *)particles, c->prev, (const float4 *)sortedPos, sortedW, sortedPhase, index, c->cell_begin, c->num_neighbors, n,
                      n, c->grid, c->params.particle_radius, 1.0f, nullptr, nullptr, nullptr, c->stream);  // the reference ABI: the reference same-phase rule
    ck_launch("collide");
}

void solveFluids(float *sortedPos, float *sortedW, int *sortedPhase, uint *index, uint *cellStart, uint *cellEnd, float *particles, uint n,
                 uint numCells) {
    (void)cellEnd;
    PsCtx *c = ctx();
    need_dense(cellStart, numCells, "solveFluids");
    (void)sortedPos;  // == the context's sorted positions up to .w (see reorderDataAndFindCellStart)
    ps_launch_find_lambdas(c->lambda, c->num_neighbors, c->spos, sortedW, sortedPhase, index, c->cell_begin, c->ros, n, n,
                           -3.0e38f, 3.0e38f, c->grid, c->stencil, false, c->nbr_list, c->nbr_rows, c->nbr_max_rows, c->capacity, true, false, c->device, c->stream);
    ps_launch_solve_fluids((float4 *)particles, c->lambda, c->spos, sortedPhase, index, c->cell_begin, c->ros, n, n, c->grid,
                           c->stencil, 1.0f, c->nbr_list, c->nbr_rows, c->nbr_max_rows, c->num_neighbors, c->device, c->stream);
    ck_launch("solveFluids");
}

void collideWorld(float *pos, float *sortedPos, uint n, PsRefInt3 minB, PsRefInt3 maxB) {
    (void)sortedPos;
    PsCtx *c = ctx();
    if (curandGenerateUniform(c->gen, c->rands, 6) != CURAND_STATUS_SUCCESS) { ps_set_error("curandGenerateUniform failed"); die("collideWorld"); }
    c->rand_calls++;
    WorldDesc w;
    w.radius = c->params.particle_radius;
    w.min_x = minB.x; w.min_y = minB.y; w.min_z = minB.z;
    w.max_x = maxB.x; w.max_y = maxB.y; w.max_z = maxB.z;
    ps_launch_collide_world((float4 *)pos, c->prev, c->phase, n, c->rands, w, c->stream);
    ck_launch("collideWorld");
}

void sortByType(float *, uint) {}  // empty in the reference as well (integration.cu:314-317)

void calcVelocity(float *pos, float dt, uint n) {
    PsCtx *c = ctx();
    ps_launch_velocity((const float4 *)pos, c->prev, c->vel, n, dt, c->stream);
    ck_launch("calcVelocity");
}

// ---------------------------------------------------------------- solver.cu ----------------------------------------------------------------
void appendSolverParticle(uint n) {
    PsCtx *c = ctx();
    g_n_solver += n;
    if (c->h_occ.size() < g_n_solver) c->h_occ.resize(g_n_solver, 0u);
    c->n = std::max(c->n, g_n_solver);
    ck(ps_ctx_ensure_capacity(c, c->n), "appendSolverParticle");
    c->constraints_dirty = true;
}
void addPointConstraint(uint *index, float *point, uint n) { ck(ps_add_point_constraints(ctx(), index, point, n), "addPointConstraint"); }
void addDistanceConstraint(uint *index, float *distance, uint n) { ck(ps_add_distance_constraints(ctx(), index, distance, n), "addDistanceConstraint"); }
void freeSolverVectors(void) {
    if (g) {
        g->h_dist_idx.clear(); g->h_dist_rest.clear(); g->h_point_idx.clear(); g->h_point_xyz.clear(); g->h_occ.clear();
        g->constraints_dirty = true;
    }
    g_n_solver = 0;
    maybe_destroy();
}
void solvePointConstraints(float *particles) {
    PsCtx *c = ctx();
    ck(ps_ctx_sync_constraints(c), "solvePointConstraints");
    ps_launch_point((float4 *)particles, c->d_point_idx, c->d_point_xyz, c->num_points, c->stream);
    ck_launch("solvePointConstraints");
}
void solveDistanceConstraints(float *particles) {
    PsCtx *c = ctx();
    ck(ps_ctx_sync_constraints(c), "solveDistanceConstraints");
    ps_launch_distance((float4 *)particles, c->dist_scratch, c->csr_particle, c->csr_off, c->csr_other, c->csr_rest, c->occ, c->num_constrained,
                       1.0f, c->stream);
    ck_launch("solveDistanceConstraints");
}

// ---------------------------------------------------------------- shared_variables.cu ----------------------------------------------------------------
void freeSharedVectors(void) { g_n_shared = 0; maybe_destroy(); }
void appendPhaseAndMass(int *fase, float *w, uint n) {
    PsCtx *c = ctx();
    ck(ps_ctx_ensure_capacity(c, (uint64_t)g_n_shared + n), "appendPhaseAndMass");
    ck_cuda(cudaMemcpy(c->phase + g_n_shared, fase, (size_t)n * 4, cudaMemcpyHostToDevice), "appendPhaseAndMass");
    ck_cuda(cudaMemcpy(c->w + g_n_shared, w, (size_t)n * 4, cudaMemcpyHostToDevice), "appendPhaseAndMass");
    g_n_shared += n;
}
void copyToXstar(float *pos, uint n) {
    PsCtx *c = ctx();
    ck_cuda(cudaMemcpyAsync(c->prev, pos, (size_t)n * 16, cudaMemcpyDeviceToDevice, c->stream), "copyToXstar");
}
int *getPhaseRawPtr(void) { return ctx()->phase; }
float *getXstarRawPtr(void) { return (float *)ctx()->prev; }
float *getWRawPtr(void) { return ctx()->w; }
void printXstar(void) {
    PsCtx *c = ctx();
    std::vector<float> h((size_t)g_n_shared * 4);
    cudaMemcpy(h.data(), c->prev, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("Xstar: size: %u\n", (uint)h.size());
    for (u32 i = 0; i < g_n_shared; i++) printf("i: %u: %.2f, %.2f, %.2f\n", i, h[4 * i], h[4 * i + 1], h[4 * i + 2]);
    printf("\n");
}

// ---------------------------------------------------------------- util.cu (the part without OpenGL) ----------------------------------------------------------------
// gpu/src/cuda/util.cuh:6-25 / util.cu: device selection, raw device arrays and blocking copies for the host class
// (particlesystem.cpp:95-142), launch-geometry helpers.  The four GL-interop entry points (register / unregister / map / unmap of the
// position VBO) stay with the viewer: a headless host maps an ordinary device allocation (oracle/ref_gpu_glue.cu shows one).
void cudaInit(void) {
    // the reference picks the device with the most GFLOP/s (findCudaDevice); here: PS_DEVICE, else the current device
    int dev = 0, count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { printf("No CUDA Capable devices found, exiting...\n"); exit(EXIT_SUCCESS); }
    if (const char *e = getenv("PS_DEVICE")) dev = atoi(e);
    else cudaGetDevice(&dev);
    ck_cuda(cudaSetDevice(dev), "cudaInit");
}
void allocateArray(void **devPtr, int size) { ck_cuda(cudaMalloc(devPtr, (size_t)(unsigned int)size), "allocateArray"); }
void freeArray(void *devPtr) { ck_cuda(cudaFree(devPtr), "freeArray"); }
void copyArrayToDevice(void *device, const void *host, int offset, int size) {
    ck_cuda(cudaMemcpy((char *)device + offset, host, (size_t)(unsigned int)size, cudaMemcpyHostToDevice), "copyArrayToDevice");
}
void copyArrayFromDevice(void *host, const void *device, int size) {
    ck_cuda(cudaMemcpy(host, device, (size_t)(unsigned int)size, cudaMemcpyDeviceToHost), "copyArrayFromDevice");
}
uint iDivUp(uint a, uint b) { return (a % b != 0) ? (a / b + 1) : (a / b); }
void computeGridSize(uint n, uint blockSize, uint &numBlocks, uint &numThreads) {
    numThreads = std::min(blockSize, n);
    numBlocks = iDivUp(n, numThreads);
}

// ---------------------------------------------------------------- additions ----------------------------------------------------------------
float *psRefVelocityPtr(void) { return (float *)ctx()->vel; }
float *psRefLambdaPtr(void) { return ctx()->lambda; }
float *psRefRestDensityPtr(void) { return ctx()->ros; }
uint *psRefNumNeighborsPtr(void) { return ctx()->num_neighbors; }
float *psRefRandsPtr(void) { return ctx()->rands; }
uint *psRefOccurrencesPtr(void) { PsCtx *c = ctx(); ck(ps_ctx_sync_constraints(c), "psRefOccurrencesPtr"); return c->occ; }
uint psRefNumDistanceConstraints(void) { return g ? (uint)g->h_dist_rest.size() : 0; }
uint psRefNumPointConstraints(void) { return g ? (uint)g->h_point_idx.size() : 0; }
void psRefCopyDistanceConstraints(uint *idx, float *rest) {
    if (!g) return;
    if (idx) memcpy(idx, g->h_dist_idx.data(), g->h_dist_idx.size() * 4);
    if (rest) memcpy(rest, g->h_dist_rest.data(), g->h_dist_rest.size() * 4);
}
void psRefCopyPointConstraints(uint *idx, float *xyz) {
    if (!g) return;
    if (idx) memcpy(idx, g->h_point_idx.data(), g->h_point_idx.size() * 4);
    if (xyz) memcpy(xyz, g->h_point_xyz.data(), g->h_point_xyz.size() * 4);
}
}  // extern "C"
