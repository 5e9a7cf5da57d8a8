// particlesolver_b200/csrc/ps_extensions.cu — the parts of the unified solver step that north_star names but the
// reference does not contain, behind the same context API:
//   rigid bodies with 3-D shape matching (K12, ps_shape_kernels.cu)       ps_add_rigid_body, ps_solve_shapes
//   XSPH viscosity + vorticity confinement (K13, ps_neighbor_kernels.cu)   ps_set_viscosity, ps_find_neighbors, ps_apply_viscosity
// Both are off unless asked for (no body added, coefficients 0), so every parity run of the reference's scenes is unaffected.
// Parity unpinned: the reference has no implementation to compare with; tests check them against float64 restatements.
#include <algorithm>
#include <cmath>
#include <vector>
#include "ps_context.h"

#define XCU(x)                                                                                       \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            ps_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return PS_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

namespace {
struct DevGuard {
    int prev = 0;
    explicit DevGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
constexpr int kShapeIters = 64;  // rotation-extraction iterations per call at most: convergence is linear (~0.7 per iteration for an
                                 // anisotropic body), a cold start after a large rotation needs ~45; warm-started calls stop after 1-3
}  // namespace

void ps_ext_free(PsCtx *c) {
    void *ptrs[] = {c->body_off, c->body_idx, c->body_rest, c->body_quat, c->body_stiff, c->visc_scratch, c->body_sdf, c->sdf_world, c->member_body};
    for (void *p : ptrs) if (p) cudaFree(p);
}

// A rigid body over existing particles; its rest shape is their CURRENT configuration (rest offsets from the centre of
// mass, masses 1 / inverse mass).  Members should share one phase >= PS_PHASE_RIGID so that they do not collide with each
// other (the contact pass skips equal phases above SOLID, integration_kernel.cuh:336-337).
extern "C" int ps_add_rigid_body(PsCtx *c, const uint32_t *indices, uint64_t n, float stiffness, uint32_t *body) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!indices || n < 2) { ps_set_error("ps_add_rigid_body: a rigid body needs at least 2 particles"); return PS_ERR_INVALID; }
    if (!(stiffness > 0.f && stiffness <= 1.f)) { ps_set_error("ps_add_rigid_body: stiffness must be in (0, 1]"); return PS_ERR_INVALID; }
    if (c->n_ghost || c->slab_used) { ps_set_error("ps_add_rigid_body: index-based constraints are not supported on slab contexts"); return PS_ERR_STATE; }
    for (uint64_t k = 0; k < n; k++)
        if (indices[k] >= c->n) { ps_set_error("ps_add_rigid_body: particle index %u >= %u particles", indices[k], c->n); return PS_ERR_INVALID; }
    DevGuard dg(c->device);
    XCU(cudaStreamSynchronize(c->stream));
    std::vector<float> pos(4 * n), w(n);
    for (uint64_t k = 0; k < n; k++) {  // bodies are small; setup path
        XCU(cudaMemcpy(&pos[4 * k], c->pos + indices[k], sizeof(float4), cudaMemcpyDeviceToHost));
        XCU(cudaMemcpy(&w[k], c->w + indices[k], sizeof(float), cudaMemcpyDeviceToHost));
        if (!(w[k] > 0.f)) { ps_set_error("ps_add_rigid_body: particle %u has infinite mass", indices[k]); return PS_ERR_INVALID; }
    }
    double cx = 0, cy = 0, cz = 0, mt = 0;
    for (uint64_t k = 0; k < n; k++) { const double m = 1.0 / w[k]; cx += m * pos[4 * k]; cy += m * pos[4 * k + 1]; cz += m * pos[4 * k + 2]; mt += m; }
    cx /= mt; cy /= mt; cz /= mt;
    for (uint64_t k = 0; k < n; k++) {
        c->h_body_idx.push_back(indices[k]);
        c->h_body_rest.push_back((float)(pos[4 * k] - cx)); c->h_body_rest.push_back((float)(pos[4 * k + 1] - cy));
        c->h_body_rest.push_back((float)(pos[4 * k + 2] - cz)); c->h_body_rest.push_back(1.0f / w[k]);
        const float none[4] = {0.f, 0.f, 0.f, -1.f};
        c->h_body_sdf.insert(c->h_body_sdf.end(), none, none + 4);
    }
    c->h_body_off.push_back((u32)c->h_body_idx.size());
    c->h_body_stiff.push_back(stiffness);
    if (body) *body = c->num_bodies;
    c->num_bodies++;
    return PS_OK;
}
extern "C" uint64_t ps_num_rigid_bodies(PsCtx *c) { return c ? c->num_bodies : 0; }

// Signed-distance data of a body's particles, in the frame the body was added in: per member (gx, gy, gz, depth) — the outward
// surface normal nearest to the particle and how deep the particle sits below the surface (the reference CPU app's SDFData,
// cpu/src/solver/particle.h:82-93: its box builders use depth = radius for face particles, radius * sqrt(2) at corners,
// simulation.cpp:666-672).  depth < 0 = this member has none.  Contacts between two particles that both carry SDF data then
// follow RigidContactConstraint (rigidcontactconstraint.cpp:13-96) instead of the centre-to-centre rule.
extern "C" int ps_set_rigid_body_sdf(PsCtx *c, uint32_t body, const float *sdf4) {
    if (!c || !sdf4) { ps_set_error("ps_set_rigid_body_sdf: null argument"); return PS_ERR_INVALID; }
    if (body >= c->num_bodies) { ps_set_error("ps_set_rigid_body_sdf: body %u of %u", body, c->num_bodies); return PS_ERR_INVALID; }
    const u32 b0 = c->h_body_off[body], b1 = c->h_body_off[body + 1];
    for (u32 k = b0; k < b1; k++) {
        const float *g = sdf4 + 4 * (size_t)(k - b0);
        if (!(g[3] >= 0.f)) continue;
        const float len = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        if (!(len > 1e-6f) || !std::isfinite(len) || !std::isfinite(g[3])) { ps_set_error("ps_set_rigid_body_sdf: member %u has depth %g but no gradient", k - b0, (double)g[3]); return PS_ERR_INVALID; }
    }
    for (u32 k = b0; k < b1; k++) {
        const float *g = sdf4 + 4 * (size_t)(k - b0);
        float *o = &c->h_body_sdf[4 * (size_t)k];
        if (!(g[3] >= 0.f)) { o[0] = o[1] = o[2] = 0.f; o[3] = -1.f; continue; }
        const float inv = 1.f / std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        o[0] = g[0] * inv; o[1] = g[1] * inv; o[2] = g[2] * inv; o[3] = g[3];
    }
    c->has_sdf = false;
    for (size_t k = 3; k < c->h_body_sdf.size(); k += 4) if (c->h_body_sdf[k] >= 0.f) { c->has_sdf = true; break; }
    c->sdf_dirty = true;
    return PS_OK;
}

int ps_ext_sync_bodies(PsCtx *c) {
    if (c->bodies_uploaded == c->num_bodies && !c->sdf_dirty && (!c->has_sdf || c->sdf_capacity >= c->capacity)) return PS_OK;
    cudaStream_t s = c->stream;
    XCU(cudaStreamSynchronize(s));
    // keep the rotations of the bodies that already existed (warm start), identity for the new ones
    std::vector<float> quat(4 * (size_t)c->num_bodies, 0.f);
    for (u32 b = 0; b < c->num_bodies; b++) quat[4 * b + 3] = 1.f;
    if (c->bodies_uploaded) XCU(cudaMemcpy(quat.data(), c->body_quat, 16 * (size_t)c->bodies_uploaded, cudaMemcpyDeviceToHost));
    void *old[] = {c->body_off, c->body_idx, c->body_rest, c->body_quat, c->body_stiff, c->body_sdf, c->sdf_world, c->member_body};
    for (void *p : old) if (p) cudaFree(p);
    c->body_off = c->body_idx = c->member_body = nullptr; c->body_rest = c->body_quat = c->body_sdf = c->sdf_world = nullptr; c->body_stiff = nullptr;
    c->sdf_capacity = 0;
    const size_t m = c->h_body_idx.size();
    XCU(cudaMalloc((void **)&c->body_off, (c->num_bodies + 1) * sizeof(u32)));
    XCU(cudaMalloc((void **)&c->body_idx, m * sizeof(u32)));
    XCU(cudaMalloc((void **)&c->body_rest, m * sizeof(float4)));
    XCU(cudaMalloc((void **)&c->body_quat, c->num_bodies * sizeof(float4)));
    XCU(cudaMalloc((void **)&c->body_stiff, c->num_bodies * sizeof(float)));
    XCU(cudaMemcpy(c->body_off, c->h_body_off.data(), (c->num_bodies + 1) * sizeof(u32), cudaMemcpyHostToDevice));
    XCU(cudaMemcpy(c->body_idx, c->h_body_idx.data(), m * sizeof(u32), cudaMemcpyHostToDevice));
    XCU(cudaMemcpy(c->body_rest, c->h_body_rest.data(), m * sizeof(float4), cudaMemcpyHostToDevice));
    XCU(cudaMemcpy(c->body_quat, quat.data(), c->num_bodies * sizeof(float4), cudaMemcpyHostToDevice));
    XCU(cudaMemcpy(c->body_stiff, c->h_body_stiff.data(), c->num_bodies * sizeof(float), cudaMemcpyHostToDevice));
    if (c->has_sdf) {
        std::vector<u32> member_body(m);
        for (u32 b = 0; b < c->num_bodies; b++)
            for (u32 k = c->h_body_off[b]; k < c->h_body_off[b + 1]; k++) member_body[k] = b;
        XCU(cudaMalloc((void **)&c->body_sdf, m * sizeof(float4)));
        XCU(cudaMalloc((void **)&c->member_body, m * sizeof(u32)));
        XCU(cudaMalloc((void **)&c->sdf_world, (size_t)c->capacity * sizeof(float4)));
        XCU(cudaMemcpy(c->body_sdf, c->h_body_sdf.data(), m * sizeof(float4), cudaMemcpyHostToDevice));
        XCU(cudaMemcpy(c->member_body, member_body.data(), m * sizeof(u32), cudaMemcpyHostToDevice));
        XCU(cudaMemset(c->sdf_world, 0xff, (size_t)c->capacity * sizeof(float4)));  // NaN depth = no SDF
        c->sdf_capacity = c->capacity;
    }
    c->sdf_dirty = false;
    c->bodies_uploaded = c->num_bodies;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    return PS_OK;
}

u32 ps_ext_issue_sdf(PsCtx *c) {
    if (!c->has_sdf) return 0;
    ps_launch_sdf_world(c->sdf_world, c->body_sdf, c->body_idx, c->member_body, c->body_quat, (u32)c->h_body_idx.size(), c->stream);
    return 1;
}

u32 ps_ext_issue_shapes(PsCtx *c) {
    if (!c->num_bodies) return 0;
    ps_launch_shape_match(c->pos, c->body_off, c->body_idx, c->body_rest, c->body_quat, c->body_stiff, c->num_bodies, kShapeIters, c->stream);
    return 1;
}

extern "C" int ps_solve_shapes(PsCtx *c) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    DevGuard dg(c->device);
    int r = ps_ext_sync_bodies(c);
    if (r != PS_OK) return r;
    ps_ext_issue_shapes(c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ps_set_error("ps_solve_shapes: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
    return PS_OK;
}
extern "C" int ps_rigid_body_rotation(PsCtx *c, uint32_t body, float *quat_xyzw) {
    if (!c || !quat_xyzw || body >= c->num_bodies) { ps_set_error("ps_rigid_body_rotation: bad argument"); return PS_ERR_INVALID; }
    DevGuard dg(c->device);
    int r = ps_ext_sync_bodies(c);
    if (r != PS_OK) return r;
    XCU(cudaStreamSynchronize(c->stream));
    XCU(cudaMemcpy(quat_xyzw, c->body_quat + body, sizeof(float4), cudaMemcpyDeviceToHost));
    return PS_OK;
}

extern "C" int ps_set_viscosity(PsCtx *c, float xsph_c, float vorticity_eps) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!(xsph_c >= 0.f && xsph_c <= 1.f) || !(vorticity_eps >= 0.f)) { ps_set_error("ps_set_viscosity: xsph_c in [0,1], vorticity_eps >= 0"); return PS_ERR_INVALID; }
    c->xsph_c = xsph_c;
    c->vorticity_eps = vorticity_eps;
    return PS_OK;
}

static int ensure_visc_scratch(PsCtx *c) {
    if (c->visc_scratch_cap >= c->capacity && c->visc_scratch) return PS_OK;
    if (c->visc_scratch) cudaFree(c->visc_scratch);
    c->visc_scratch = nullptr;
    XCU(cudaMalloc((void **)&c->visc_scratch, 2 * (size_t)c->capacity * sizeof(float4)));
    c->visc_scratch_cap = c->capacity;
    return PS_OK;
}

u32 ps_ext_issue_viscosity(PsCtx *c, float dt) {
    if (c->xsph_c == 0.f && c->vorticity_eps == 0.f) return 0;
    if (!c->n || !c->visc_scratch || !c->grid_valid) return 0;
    return ps_launch_viscosity(c->vel, c->visc_scratch, c->spos, c->sphase, c->index, c->cell_begin, c->n, c->grid, c->stencil, c->xsph_c, c->vorticity_eps,
                               dt, c->nbr_list, c->nbr_rows, c->nbr_max_rows, c->num_neighbors, c->device, c->stream);
}

// K6 alone on the current grid: fills lambda, the neighbour counts and the neighbour lists without moving anything
extern "C" int ps_find_neighbors(PsCtx *c) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!c->n) return PS_OK;
    if (!c->grid_valid) { ps_set_error("ps_find_neighbors before ps_build_grid"); return PS_ERR_STATE; }
    DevGuard dg(c->device);
    ps_launch_find_lambdas(c->lambda, c->num_neighbors, c->spos, c->sw, c->sphase, c->index, c->cell_begin, c->ros, c->n, c->n - c->n_ghost, c->lambda_xmin,
                           c->lambda_xmax, c->grid, c->stencil, (c->params.flags & PS_FLAG_ZERO_NONFLUID_LAMBDA) != 0, c->nbr_list, c->nbr_rows,
                           c->nbr_max_rows, c->capacity, true, (c->params.flags & PS_FLAG_STAGED_LAMBDA) != 0, c->device, c->stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ps_set_error("ps_find_neighbors: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
    return PS_OK;
}

// velocity post-pass on the neighbour structure of the last grid build + K6 (ps_step runs it after the velocity update when
// a coefficient is non-zero; standalone: ps_build_grid, ps_find_neighbors, ps_apply_viscosity)
extern "C" int ps_apply_viscosity(PsCtx *c, float dt) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!c->n) return PS_OK;
    if (c->n_ghost) { ps_set_error("ps_apply_viscosity: ghost particles carry no velocity (slab contexts)"); return PS_ERR_STATE; }
    if (!c->grid_valid) { ps_set_error("ps_apply_viscosity before ps_build_grid / ps_find_neighbors"); return PS_ERR_STATE; }
    DevGuard dg(c->device);
    int r = ensure_visc_scratch(c);
    if (r != PS_OK) return r;
    ps_ext_issue_viscosity(c, dt);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ps_set_error("ps_apply_viscosity: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
    return PS_OK;
}

// Diagnostics of the current state (north_star's long-run parity bar is stated in these): the mean and the largest density
// error |rho_i / rho0_i - 1| over the fluid particles, rho_i being the K6 estimate on a freshly built grid
// (integration_kernel.cuh:565-589), and the kinetic energy sum 1/2 m v^2.  Rebuilds the grid and the neighbour lists from the
// current positions; positions and velocities are not touched.
extern "C" int ps_fluid_stats(PsCtx *c, double *mean_density_error, double *max_density_error, double *kinetic_energy) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (mean_density_error) *mean_density_error = 0.;
    if (max_density_error) *max_density_error = 0.;
    if (kinetic_energy) *kinetic_energy = 0.;
    if (!c->n) return PS_OK;
    // a slab context: statistics over the OWNED particles, with the ghosts of the last halo refresh as their neighbours
    const u32 n_owned = c->n - c->n_ghost;
    DevGuard dg(c->device);
    int r = ensure_visc_scratch(c);
    if (r != PS_OK) return r;
    ps_issue_build_grid(c, c->pos);
    if ((r = ps_find_neighbors(c)) != PS_OK) return r;
    const u32 n = c->n;
    XCU(cudaMemsetAsync(c->visc_scratch, 0, (size_t)n * sizeof(float4), c->stream));
    ps_launch_density_error(c->visc_scratch, c->spos, c->sw, c->sphase, c->index, c->ros, c->cell_begin, n, c->grid, c->stencil, c->nbr_list, c->nbr_rows,
                            c->nbr_max_rows, c->num_neighbors, c->device, c->stream);
    std::vector<float> err(4 * (size_t)n), vel(4 * (size_t)n), w(n);
    std::vector<int> sph(n);
    std::vector<u32> idx(n);
    XCU(cudaMemcpyAsync(err.data(), c->visc_scratch, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    XCU(cudaMemcpyAsync(sph.data(), c->sphase, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    XCU(cudaMemcpyAsync(idx.data(), c->index, (size_t)n * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    XCU(cudaMemcpyAsync(vel.data(), c->vel, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    XCU(cudaMemcpyAsync(w.data(), c->w, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    XCU(cudaStreamSynchronize(c->stream));
    double sum = 0., mx = 0., ke = 0.;
    uint64_t cnt = 0;
    for (u32 i = 0; i < n; i++) {
        if (sph[i] == PH_FLUID && idx[i] < n_owned) { const double e = err[4 * (size_t)i]; sum += e; mx = std::max(mx, e); cnt++; }
        if (i < n_owned && w[i] != 0.f) ke += 0.5 * ((double)vel[4 * (size_t)i] * vel[4 * (size_t)i] + (double)vel[4 * (size_t)i + 1] * vel[4 * (size_t)i + 1] +
                                       (double)vel[4 * (size_t)i + 2] * vel[4 * (size_t)i + 2]) / w[i];
    }
    if (mean_density_error) *mean_density_error = cnt ? sum / (double)cnt : 0.;
    if (max_density_error) *max_density_error = mx;
    if (kinetic_energy) *kinetic_energy = ke;
    return PS_OK;
}

// called by ps_step's ready(): scratch must exist before the step is captured into a graph
int ps_ext_prepare_step(PsCtx *c) {
    if (c->xsph_c == 0.f && c->vorticity_eps == 0.f) return PS_OK;
    return ensure_visc_scratch(c);
}
