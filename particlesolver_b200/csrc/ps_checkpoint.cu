// particlesolver_b200/csrc/ps_checkpoint.cu — binary checkpoints of a solver context (scene / state
// I/O).  The reference has no persistence at all (scenes exist only as code, particleapp.cpp:141-215); this is the
// restartable state of a PsCtx: parameters, the SoA particle arrays, the constraint lists in insertion order, rigid
// bodies with their rotations, viscosity coefficients and the position of the wall-jitter stream (cuRAND XORWOW, seed
// 1234: repositioned by replaying the same number of 6-value draws) and — version 3 — the lambda array by sorted slot: K6 leaves
// the lambda of non-fluid slots untouched (the reference's behaviour; PS_FLAG_ZERO_NONFLUID_LAMBDA off) and K7 reads lambda_j of
// every neighbour whatever its phase, so in a scene whose fluid touches solids those stale values are state that crosses steps.
// A run continued from a checkpoint is bit-identical to the uninterrupted run (tests/test_gpu_checkpoint.py).  Little-endian,
// fixed-width fields, no pointers.
#include <cstdio>
#include <cstring>
#include <vector>
#include "ps_context.h"

#define KCU(x)                                                                                       \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            ps_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return PS_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

namespace {
const char kMagic[8] = {'P', 'S', 'B', '2', '0', '0', 'C', '1'};
struct Header {
    char magic[8];
    uint32_t version, params_bytes;
    uint64_t n, limit, num_distance, num_point, num_bodies, body_members, rand_calls;
    float xsph_c, vorticity_eps;
};
struct File {
    FILE *f = nullptr;
    ~File() { if (f) fclose(f); }
};
template <class T>
bool put(FILE *f, const T *p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <class T>
bool get(FILE *f, T *p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
}  // namespace

extern "C" int ps_save(PsCtx *c, const char *path) {
    if (!c || !path) { ps_set_error("ps_save: null argument"); return PS_ERR_INVALID; }
    if (c->n_ghost) { ps_set_error("ps_save: slab contexts hold ghost copies; save after ps_set_ghost_count(ctx, 0)"); return PS_ERR_STATE; }
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != c->device) cudaSetDevice(c->device);
    int r = ps_ext_sync_bodies(c);
    if (r != PS_OK) return r;
    KCU(cudaStreamSynchronize(c->stream));
    const size_t n = c->n;
    std::vector<float> pos(4 * n), vel(4 * n), prev(4 * n), w(n), ros(n), quat(4 * (size_t)c->num_bodies), lam(n);
    std::vector<int> phase(n);
    KCU(cudaMemcpy(lam.data(), c->lambda, 4 * n, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(pos.data(), c->pos, 16 * n, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(vel.data(), c->vel, 16 * n, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(prev.data(), c->prev, 16 * n, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(w.data(), c->w, 4 * n, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(ros.data(), c->ros, 4 * n, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(phase.data(), c->phase, 4 * n, cudaMemcpyDeviceToHost));
    if (c->num_bodies) KCU(cudaMemcpy(quat.data(), c->body_quat, 16 * (size_t)c->num_bodies, cudaMemcpyDeviceToHost));
    Header h{};
    memcpy(h.magic, kMagic, 8);
    h.version = 3; h.params_bytes = (uint32_t)sizeof(PsParams);
    h.n = n; h.limit = c->limit; h.num_distance = c->h_dist_rest.size(); h.num_point = c->h_point_idx.size();
    h.num_bodies = c->num_bodies; h.body_members = c->h_body_idx.size(); h.rand_calls = c->rand_calls;
    h.xsph_c = c->xsph_c; h.vorticity_eps = c->vorticity_eps;
    File F;
    F.f = fopen(path, "wb");
    if (!F.f) { ps_set_error("ps_save: cannot open %s for writing", path); return PS_ERR_INVALID; }
    bool ok = put(F.f, &h, 1) && put(F.f, &c->params, 1) && put(F.f, pos.data(), pos.size()) && put(F.f, vel.data(), vel.size()) &&
              put(F.f, prev.data(), prev.size()) && put(F.f, w.data(), n) && put(F.f, ros.data(), n) && put(F.f, phase.data(), n) &&
              put(F.f, c->h_dist_idx.data(), c->h_dist_idx.size()) && put(F.f, c->h_dist_rest.data(), c->h_dist_rest.size()) &&
              put(F.f, c->h_point_idx.data(), c->h_point_idx.size()) && put(F.f, c->h_point_xyz.data(), c->h_point_xyz.size()) &&
              put(F.f, c->h_body_off.data(), c->h_body_off.size()) && put(F.f, c->h_body_idx.data(), c->h_body_idx.size()) &&
              put(F.f, c->h_body_rest.data(), c->h_body_rest.size()) && put(F.f, c->h_body_stiff.data(), c->h_body_stiff.size()) &&
              put(F.f, quat.data(), quat.size());
    const uint64_t sdf_members = c->has_sdf ? c->h_body_idx.size() : 0;  // version 2: SDF data of the rigid bodies (4 floats per member) or none
    ok = ok && put(F.f, &sdf_members, 1) && put(F.f, c->h_body_sdf.data(), (size_t)(4 * sdf_members));
    ok = ok && put(F.f, lam.data(), n);  // version 3: lambda by sorted slot
    if (!ok || fflush(F.f) != 0) { ps_set_error("ps_save: short write to %s", path); return PS_ERR_INVALID; }
    return PS_OK;
}

extern "C" int ps_load(const char *path, int device, PsCtx **out) {
    if (!path || !out) { ps_set_error("ps_load: null argument"); return PS_ERR_INVALID; }
    *out = nullptr;
    File F;
    F.f = fopen(path, "rb");
    if (!F.f) { ps_set_error("ps_load: cannot open %s", path); return PS_ERR_INVALID; }
    Header h{};
    if (!get(F.f, &h, 1) || memcmp(h.magic, kMagic, 8) != 0) { ps_set_error("ps_load: %s is not a libpsolver checkpoint", path); return PS_ERR_INVALID; }
    if (h.version < 1 || h.version > 3 || h.params_bytes != sizeof(PsParams)) { ps_set_error("ps_load: checkpoint version %u / parameter block of %u bytes not understood", h.version, h.params_bytes); return PS_ERR_INVALID; }
    if (h.n > h.limit || h.limit > (1ull << 31) || h.num_distance > (1ull << 32) || h.num_point > (1ull << 32) || h.body_members > h.n * 64 + 64 || h.num_bodies > h.body_members) {
        ps_set_error("ps_load: implausible sizes in %s", path); return PS_ERR_INVALID;
    }
    PsParams p;
    if (!get(F.f, &p, 1)) { ps_set_error("ps_load: truncated file"); return PS_ERR_INVALID; }
    const size_t n = h.n;
    std::vector<float> pos(4 * n), vel(4 * n), prev(4 * n), w(n), ros(n), drest(h.num_distance), pxyz(3 * h.num_point), brest(4 * h.body_members),
        bstiff(h.num_bodies), quat(4 * h.num_bodies);
    std::vector<int> phase(n);
    std::vector<u32> didx(2 * h.num_distance), pidx(h.num_point), boff(h.num_bodies + 1), bidx(h.body_members);
    bool ok = get(F.f, pos.data(), pos.size()) && get(F.f, vel.data(), vel.size()) && get(F.f, prev.data(), prev.size()) && get(F.f, w.data(), n) &&
              get(F.f, ros.data(), n) && get(F.f, phase.data(), n) && get(F.f, didx.data(), didx.size()) && get(F.f, drest.data(), drest.size()) &&
              get(F.f, pidx.data(), pidx.size()) && get(F.f, pxyz.data(), pxyz.size()) && get(F.f, boff.data(), boff.size()) &&
              get(F.f, bidx.data(), bidx.size()) && get(F.f, brest.data(), brest.size()) && get(F.f, bstiff.data(), bstiff.size()) &&
              get(F.f, quat.data(), quat.size());
    std::vector<float> bsdf(4 * h.body_members, 0.f);
    for (size_t k = 3; k < bsdf.size(); k += 4) bsdf[k] = -1.f;
    bool has_sdf = false;
    if (ok && h.version >= 2) {
        uint64_t sdf_members = 0;
        ok = get(F.f, &sdf_members, 1) && (sdf_members == 0 || sdf_members == h.body_members) && get(F.f, bsdf.data(), (size_t)(4 * sdf_members));
        has_sdf = ok && sdf_members != 0;
    }
    std::vector<float> lam;
    if (ok && h.version >= 3) { lam.resize(n); ok = get(F.f, lam.data(), n); }
    if (!ok) { ps_set_error("ps_load: truncated file %s", path); return PS_ERR_INVALID; }
    for (u32 v : bidx) if (v >= n) { ps_set_error("ps_load: body member index out of range"); return PS_ERR_INVALID; }
    if (boff[0] != 0 || boff[h.num_bodies] != h.body_members) { ps_set_error("ps_load: inconsistent body table"); return PS_ERR_INVALID; }
    for (size_t b = 0; b < h.num_bodies; b++)
        if (boff[b + 1] < boff[b]) { ps_set_error("ps_load: body offsets are not ascending"); return PS_ERR_INVALID; }
    for (size_t k = 0; k < didx.size(); k++) if (didx[k] >= n) { ps_set_error("ps_load: distance constraint index out of range"); return PS_ERR_INVALID; }
    for (u32 v : pidx) if (v >= n) { ps_set_error("ps_load: point constraint index out of range"); return PS_ERR_INVALID; }
    PsCtx *c = nullptr;
    int r = ps_create(device, &p, h.limit, &c);
    if (r != PS_OK) return r;
    auto fail = [&](int code) { ps_destroy(c); return code; };
    if (n && (r = ps_append_particles(c, pos.data(), vel.data(), w.data(), ros.data(), phase.data(), n)) != PS_OK) return fail(r);
    if (n && (r = ps_upload(c, PS_ARR_PREV, prev.data(), 0, 4 * n)) != PS_OK) return fail(r);
    if (n && !lam.empty() && cudaMemcpy(c->lambda, lam.data(), 4 * n, cudaMemcpyHostToDevice) != cudaSuccess) { ps_set_error("ps_load: lambda upload failed"); return fail(PS_ERR_CUDA); }
    if ((r = ps_add_distance_constraints(c, didx.data(), drest.data(), h.num_distance)) != PS_OK) return fail(r);
    if ((r = ps_add_point_constraints(c, pidx.data(), pxyz.data(), h.num_point)) != PS_OK) return fail(r);
    // bodies: the stored rest shape, not the current configuration
    c->h_body_off = boff; c->h_body_idx = bidx; c->h_body_rest = brest; c->h_body_stiff = bstiff;
    c->h_body_sdf = bsdf; c->has_sdf = has_sdf; c->sdf_dirty = has_sdf;
    c->num_bodies = (u32)h.num_bodies;
    c->bodies_uploaded = 0;
    if ((r = ps_ext_sync_bodies(c)) != PS_OK) return fail(r);
    if (h.num_bodies && cudaMemcpy(c->body_quat, quat.data(), 16 * h.num_bodies, cudaMemcpyHostToDevice) != cudaSuccess) { ps_set_error("ps_load: body upload failed"); return fail(PS_ERR_CUDA); }
    c->xsph_c = h.xsph_c; c->vorticity_eps = h.vorticity_eps;
    // reposition the wall-jitter stream: the same draws in the same call pattern
    for (uint64_t k = 0; k < h.rand_calls; k++)
        if (curandGenerateUniform(c->gen, c->rands, 6) != CURAND_STATUS_SUCCESS) { ps_set_error("ps_load: curandGenerateUniform failed"); return fail(PS_ERR_CUDA); }
    c->rand_calls = h.rand_calls;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { ps_set_error("ps_load: stream error"); return fail(PS_ERR_CUDA); }
    *out = c;
    return PS_OK;
}
