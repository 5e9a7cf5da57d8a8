// particlesolver_b200/csrc/ps_comm.cu — the slab-decomposed step behind the C ABI (ps_comm_init(ncclUniqueId, rank,
// nranks); §8e): one context per GPU, one process per context, neighbour exchange over NCCL point-to-point calls issued on the
// context's own stream.  The host logic is particlesolver_b200/slab.py's SlabDomain.step, restated in C++ so that the C++ host
// (psolver_cli --scene c5) runs the multi-GPU configuration without Python:
//     predict -> migrate (owned particles whose predicted x left [x_lo, x_hi) move to the neighbour with their full state)
//     iterations x [ refresh halo -> grid build, contacts, K6 -> ghost lambdas from their owners -> K7, world bounds ]
//     velocity update
// The reference is single-GPU (nothing to cite); the stage calls are the same ps_* entry points a single context uses.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 in ps_comm_init): libpsolver.so carries no link-time dependency on it, a
// single-GPU host never loads it, and inside a process that already holds an NCCL (torch's) that one is used.
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "ps_context.h"

namespace {
// the slice of nccl.h this file uses (NCCL's C ABI is stable across 2.x: opaque comm handle, 128-byte unique id, enum values below)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
enum { kNcclUint8 = 1, kNcclFloat64 = 8, kNcclSum = 0 };  // ncclDataType_t / ncclRedOp_t values of nccl.h
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
    if (g_nccl.lib) return PS_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) { ps_set_error("ps_comm: cannot load libnccl.so.2 (%s)", dlerror()); return PS_ERR_STATE; }
    auto sym = [&](const char *n) { return dlsym(h, n); };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd ||
        !g_nccl.AllReduce) {
        ps_set_error("ps_comm: libnccl lacks an expected symbol");
        return PS_ERR_STATE;
    }
    g_nccl.lib = h;
    return PS_OK;
}

#define NC(call)                                                                                                               \
    do {                                                                                                                       \
        ncclResult_t r_ = (call);                                                                                              \
        if (r_ != 0) { ps_set_error("%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error"); return PS_ERR_CUDA; } \
    } while (0)
#define CC(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) { ps_set_error("%s: %s", #call, cudaGetErrorString(e_)); return PS_ERR_CUDA; } \
    } while (0)
#define OK(call)                    \
    do {                            \
        int r_ = (call);            \
        if (r_ != PS_OK) return r_; \
    } while (0)

struct DevGuard {
    int prev = 0;
    explicit DevGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

// state of a context's communicator and slab (PsCtx::comm)
struct PsComm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    float x_lo = -INFINITY, x_hi = INFINITY, halo = 2.25f, lambda_ext = 2.25f;
    bool exchange_lambda = true, slab_set = false;
    uint64_t halo_cap = 0, migr_cap = 0;
    // record buffers: [0] towards / from the left neighbour, [1] the right
    void *halo_send[2] = {nullptr, nullptr}, *halo_recv[2] = {nullptr, nullptr};
    void *migr_send[2] = {nullptr, nullptr}, *migr_recv[2] = {nullptr, nullptr};
    void *lam_send[2] = {nullptr, nullptr}, *lam_recv[2] = {nullptr, nullptr};
    uint32_t *counts_dev = nullptr;   // [0..1] my counts to (left, right), [2..3] counts from (left, right)
    uint32_t *counts_host = nullptr;  // pinned mirror
    double *reduce_dev = nullptr;     // 8 doubles for ps_comm_allreduce_sum
    uint32_t ghost_counts[2] = {0, 0};
    uint64_t migrated_out = 0, ghosts = 0, bytes_sent = 0, steps = 0;
    // the global phase census taken by ps_comm_set_slab: no rank was ever handed a contact-phase particle => the contact pass is skipped
    bool any_contact = true;
    uint64_t contact_sources_seen = 0, nonfluid_sources_seen = 0;
    // load balancing (ps_comm_set_recut): every `recut_every` steps the cut planes move to the equal-count quantiles of the particles' x
    std::vector<double> cuts;         // nranks + 1 planes, outer ones +-inf; empty: fixed slab (ps_comm_set_slab alone)
    uint32_t recut_every = 0, recut_bins = 0, recuts = 0;
    float recut_xmin = 0.f, recut_xmax = 0.f;
    double *hist_dev = nullptr;       // recut_bins doubles for the all-reduce of the x-histograms
    uint32_t hist_cap = 0;
};

static void comm_free_buffers(PsComm *m) {
    void **all[] = {&m->halo_send[0], &m->halo_send[1], &m->halo_recv[0], &m->halo_recv[1], &m->migr_send[0], &m->migr_send[1], &m->migr_recv[0],
                    &m->migr_recv[1], &m->lam_send[0], &m->lam_send[1], &m->lam_recv[0], &m->lam_recv[1]};
    for (void **p : all) { if (*p) cudaFree(*p); *p = nullptr; }
}

void ps_comm_free(PsCtx *c) {
    PsComm *m = c->comm;
    if (!m) return;
    cudaStreamSynchronize(c->stream);
    if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
    comm_free_buffers(m);
    if (m->counts_dev) cudaFree(m->counts_dev);
    if (m->counts_host) cudaFreeHost(m->counts_host);
    if (m->reduce_dev) cudaFree(m->reduce_dev);
    if (m->hist_dev) cudaFree(m->hist_dev);
    delete m;
    c->comm = nullptr;
}

extern "C" int ps_comm_get_unique_id(void *id128) {
    if (!id128) { ps_set_error("ps_comm_get_unique_id: null buffer"); return PS_ERR_INVALID; }
    OK(load_nccl());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return PS_OK;
}

extern "C" int ps_comm_init(PsCtx *c, const void *id128, int rank, int nranks) {
    if (!c || !id128) { ps_set_error("ps_comm_init: null argument"); return PS_ERR_INVALID; }
    if (nranks < 1 || rank < 0 || rank >= nranks) { ps_set_error("ps_comm_init: rank %d of %d", rank, nranks); return PS_ERR_INVALID; }
    if (c->comm) { ps_set_error("ps_comm_init: the context already has a communicator"); return PS_ERR_STATE; }
    OK(load_nccl());
    DevGuard dg(c->device);
    PsComm *m = new PsComm;
    m->rank = rank; m->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&m->comm, nranks, id, rank);
    if (r != 0) { ps_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"); delete m; return PS_ERR_CUDA; }
    if (cudaMalloc((void **)&m->counts_dev, 4 * sizeof(uint32_t)) != cudaSuccess || cudaMallocHost((void **)&m->counts_host, 4 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc((void **)&m->reduce_dev, 8 * sizeof(double)) != cudaSuccess) {
        ps_set_error("ps_comm_init: allocation failed");
        c->comm = m; ps_comm_free(c);
        return PS_ERR_CUDA;
    }
    c->comm = m;
    return PS_OK;
}

extern "C" int ps_comm_destroy(PsCtx *c) {
    if (!c) return PS_OK;
    DevGuard dg(c->device);
    ps_comm_free(c);
    return PS_OK;
}

// This rank's slab [x_lo, x_hi) (the outer faces of the first / last rank at -inf / +inf), the drift bound of one step's solver
// iterations, whether ghost lambdas are exchanged (halo = H + drift) or computed locally (halo = 2 H + 2 drift), buffer capacities.
extern "C" int ps_comm_set_slab(PsCtx *c, float x_lo, float x_hi, float drift, int exchange_lambda, uint64_t halo_capacity, uint64_t migrant_capacity) {
    if (!c || !c->comm) { ps_set_error("ps_comm_set_slab: no communicator (ps_comm_init)"); return PS_ERR_STATE; }
    if (!(x_lo < x_hi) || !(drift >= 0.f) || !halo_capacity || !migrant_capacity) { ps_set_error("ps_comm_set_slab: bad arguments"); return PS_ERR_INVALID; }
    PsComm *m = c->comm;
    DevGuard dg(c->device);
    CC(cudaStreamSynchronize(c->stream));
    m->x_lo = x_lo; m->x_hi = x_hi;
    m->exchange_lambda = exchange_lambda != 0;
    m->lambda_ext = PS_H + drift;
    m->halo = m->exchange_lambda ? PS_H + drift : 2.f * PS_H + 2.f * drift;
    if (halo_capacity != m->halo_cap || migrant_capacity != m->migr_cap) {
        comm_free_buffers(m);
        for (int k = 0; k < 2; k++) {
            CC(cudaMalloc(&m->halo_send[k], halo_capacity * PS_HALO_RECORD_BYTES));
            CC(cudaMalloc(&m->halo_recv[k], halo_capacity * PS_HALO_RECORD_BYTES));
            CC(cudaMalloc(&m->migr_send[k], migrant_capacity * PS_MIGRANT_RECORD_BYTES));
            CC(cudaMalloc(&m->migr_recv[k], migrant_capacity * PS_MIGRANT_RECORD_BYTES));
            CC(cudaMalloc(&m->lam_send[k], halo_capacity * sizeof(float)));
            CC(cudaMalloc(&m->lam_recv[k], halo_capacity * sizeof(float)));
        }
        m->halo_cap = halo_capacity; m->migr_cap = migrant_capacity;
    }
    // exchanged lambdas: no ghost computes one (empty range); local: ghosts within H + drift of a face do
    if (m->exchange_lambda) OK(ps_slab_set_lambda_range(c, 1.f, -1.f));
    else OK(ps_slab_set_lambda_range(c, x_lo - m->lambda_ext, x_hi + m->lambda_ext));
    // one global agreement on the phases in play (particles only ever move between ranks, phases never change): a fluid-only run
    // skips the contact pass, which would otherwise launch over all owned + ghost slots five times a step just to find nothing
    double census[2] = {(double)c->contact_sources, (double)c->nonfluid_sources};
    CC(cudaMemcpyAsync(m->reduce_dev, census, sizeof census, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(m->reduce_dev, m->reduce_dev, 2, kNcclFloat64, kNcclSum, m->comm, c->stream));
    CC(cudaMemcpyAsync(census, m->reduce_dev, sizeof census, cudaMemcpyDeviceToHost, c->stream));
    CC(cudaStreamSynchronize(c->stream));
    m->any_contact = census[0] > 0.;
    m->contact_sources_seen = c->contact_sources;
    m->nonfluid_sources_seen = c->nonfluid_sources;
    // an all-fluid run: K6 fills the outgoing lambda messages itself (ps_slab_set_lambda_sinks), no pack pass over the sorted slots
    static const bool no_sinks = getenv("PS_NO_LAMBDA_SINKS") != nullptr;  // (measurements: the separate pack pass instead)
    if (m->exchange_lambda && census[1] == 0. && !no_sinks) OK(ps_slab_set_lambda_sinks(c, m->lam_send[0], m->lam_send[1], m->halo_cap));
    else OK(ps_slab_set_lambda_sinks(c, nullptr, nullptr, 0));
    m->slab_set = true;
    return PS_OK;
}

// ---- load balancing: re-cutting the slabs (slab.py: SlabDomain.apply_recut / balanced_cuts, restated) ----
// New cut planes from the global histogram of the particles' x (hist[b] = count in bin b of [x_min, x_max)): the equal-count
// quantiles, linearly interpolated inside a bin, then limited — a cut moves by at most max_shift per re-cut (particles change owner
// by hopping to the NEIGHBOURING rank, one hop per step, so a cut must not jump over a slab) and slabs stay at least min_width wide
// (a rank's ghosts all come from its two neighbours: the halo must fit inside a slab).  A pure function of its arguments: every rank
// derives the same planes from the all-reduced histogram.
static std::vector<double> balanced_cuts(const std::vector<double> &hist, double x_min, double x_max, const std::vector<double> &old_cuts, double min_width,
                                         double max_shift) {
    const int nranks = (int)old_cuts.size() - 1, bins = (int)hist.size();
    std::vector<double> out(nranks + 1);
    out[0] = -INFINITY; out[nranks] = INFINITY;
    if (nranks == 1) return out;
    std::vector<double> cum(bins + 1, 0.0);
    for (int b = 0; b < bins; b++) cum[b + 1] = cum[b] + hist[b];
    const double total = cum[bins];
    for (int r = 1; r < nranks; r++) {
        const double target = total * r / nranks;
        int b = (int)(std::upper_bound(cum.begin(), cum.end(), target) - cum.begin()) - 1;   // np.searchsorted(cum, target, side="right") - 1
        b = std::min(std::max(b, 0), bins - 1);
        const double frac = hist[b] > 0 ? (target - cum[b]) / hist[b] : 0.5;
        const double e0 = x_min + (x_max - x_min) * b / bins, e1 = x_min + (x_max - x_min) * (b + 1) / bins;
        double cnew = e0 + frac * (e1 - e0);
        const double old = old_cuts[r];
        if (std::isfinite(old)) cnew = std::min(std::max(cnew, old - max_shift), old + max_shift);
        if (std::isfinite(out[r - 1])) cnew = std::max(cnew, out[r - 1] + min_width);
        out[r] = cnew;
    }
    for (int r = nranks - 1; r > 1; r--)  // keep the minimum width from the right as well
        if (out[r] - out[r - 1] < min_width) out[r - 1] = out[r] - min_width;
    return out;
}

static int apply_cuts(PsCtx *c, const std::vector<double> &cuts) {
    PsComm *m = c->comm;
    m->cuts = cuts;
    m->x_lo = (float)cuts[m->rank]; m->x_hi = (float)cuts[m->rank + 1];
    if (!m->exchange_lambda) OK(ps_slab_set_lambda_range(c, m->x_lo - m->lambda_ext, m->x_hi + m->lambda_ext));
    return PS_OK;
}

// All the cut planes (nranks + 1 values, cuts[0] = -INFINITY, cuts[nranks] = INFINITY; this rank owns [cuts[rank], cuts[rank + 1])) and the
// re-cut schedule: every `every` steps (0: never), right after the predict, the planes move to the equal-count quantiles of the owned
// particles' x — per-rank histograms over `bins` bins of [x_min, x_max), all-reduced — and the migration that follows hands the
// particles over.  Call after ps_comm_set_slab, with the same arguments on every rank.
extern "C" int ps_comm_set_recut(PsCtx *c, const float *cuts, uint32_t every, float x_min, float x_max, uint32_t bins) {
    if (!c || !c->comm || !c->comm->slab_set) { ps_set_error("ps_comm_set_recut: call ps_comm_init and ps_comm_set_slab first"); return PS_ERR_STATE; }
    PsComm *m = c->comm;
    if (!cuts) { ps_set_error("ps_comm_set_recut: null cut planes"); return PS_ERR_INVALID; }
    if (every && (!(x_min < x_max) || bins < 2 || bins > 65536)) { ps_set_error("ps_comm_set_recut: need x_min < x_max and 2..65536 bins"); return PS_ERR_INVALID; }
    std::vector<double> cv(m->nranks + 1);
    for (int r = 0; r <= m->nranks; r++) cv[r] = cuts[r];
    cv[0] = -INFINITY; cv[m->nranks] = INFINITY;
    for (int r = 0; r < m->nranks; r++)
        if (!(cv[r] < cv[r + 1])) { ps_set_error("ps_comm_set_recut: cut planes must ascend"); return PS_ERR_INVALID; }
    DevGuard dg(c->device);
    if (every && bins > m->hist_cap) {
        if (m->hist_dev) CC(cudaFree(m->hist_dev));
        CC(cudaMalloc((void **)&m->hist_dev, bins * sizeof(double)));
        m->hist_cap = bins;
    }
    m->recut_every = every; m->recut_bins = bins; m->recut_xmin = x_min; m->recut_xmax = x_max;
    return apply_cuts(c, cv);
}

static int recut(PsCtx *c) {
    PsComm *m = c->comm;
    const uint32_t bins = m->recut_bins;
    std::vector<uint64_t> mine(bins);
    OK(ps_slab_x_histogram(c, m->recut_xmin, m->recut_xmax, bins, mine.data()));
    std::vector<double> h(bins);
    for (uint32_t b = 0; b < bins; b++) h[b] = (double)mine[b];  // (exact: counts are far below 2^53)
    CC(cudaMemcpyAsync(m->hist_dev, h.data(), bins * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(m->hist_dev, m->hist_dev, bins, kNcclFloat64, kNcclSum, m->comm, c->stream));
    CC(cudaMemcpyAsync(h.data(), m->hist_dev, bins * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CC(cudaStreamSynchronize(c->stream));
    double width = INFINITY;  // narrowest interior slab
    for (int r = 1; r + 2 <= m->nranks - 1 + 1 && r + 1 < m->nranks; r++) width = std::min(width, m->cuts[r + 1] - m->cuts[r]);
    const double halo = m->halo;
    const double max_shift = std::isfinite(width) ? 0.5 * std::min(width, 4.0 * halo) : 2.0 * halo;
    OK(apply_cuts(c, balanced_cuts(h, m->recut_xmin, m->recut_xmax, m->cuts, halo + 0.5, max_shift)));
    m->recuts++;
    return PS_OK;
}

// Sends `counts[0]` records of to[0] to rank - 1 and `counts[1]` of to[1] to rank + 1, receives the neighbours' into from[0] /
// from[1].  recv_known: the receive counts are already known (the lambda exchange answers the halo exchange record by record), so
// nothing but the payload travels and the host does not wait; otherwise the counts are exchanged first (one host synchronisation).
static int exchange(PsCtx *c, void *const to[2], const uint32_t counts[2], void *const from[2], uint32_t recv[2], size_t record_bytes, uint64_t cap,
                    bool recv_known) {
    PsComm *m = c->comm;
    const int r = m->rank, w = m->nranks;
    const bool has_l = r > 0, has_r = r < w - 1;
    cudaStream_t s = c->stream;
    if (!recv_known) {
        m->counts_host[0] = counts[0]; m->counts_host[1] = counts[1]; m->counts_host[2] = 0; m->counts_host[3] = 0;
        // (kernel copies between pinned host memory and the device: a copy-engine transfer would queue behind a streamed step's bulk I/O)
        ps_launch_copy_words(m->counts_dev, m->counts_host, 4, s);
        NC(g_nccl.GroupStart());
        if (has_l) { NC(g_nccl.Send(m->counts_dev + 0, 4, kNcclUint8, r - 1, m->comm, s)); NC(g_nccl.Recv(m->counts_dev + 2, 4, kNcclUint8, r - 1, m->comm, s)); }
        if (has_r) { NC(g_nccl.Send(m->counts_dev + 1, 4, kNcclUint8, r + 1, m->comm, s)); NC(g_nccl.Recv(m->counts_dev + 3, 4, kNcclUint8, r + 1, m->comm, s)); }
        NC(g_nccl.GroupEnd());
        ps_launch_copy_words(m->counts_host + 2, m->counts_dev + 2, 2, s);
        CC(cudaStreamSynchronize(s));
        recv[0] = has_l ? m->counts_host[2] : 0;
        recv[1] = has_r ? m->counts_host[3] : 0;
    }
    if (!has_l) recv[0] = 0;
    if (!has_r) recv[1] = 0;
    if (recv[0] > cap || recv[1] > cap) { ps_set_error("ps_comm: a neighbour sends %u / %u records, the receive buffers hold %llu", recv[0], recv[1], (unsigned long long)cap); return PS_ERR_CAPACITY; }
    const bool any = (has_l && (counts[0] || recv[0])) || (has_r && (counts[1] || recv[1]));
    if (any) {
        NC(g_nccl.GroupStart());
        if (has_l && counts[0]) NC(g_nccl.Send(to[0], (size_t)counts[0] * record_bytes, kNcclUint8, r - 1, m->comm, s));
        if (has_r && counts[1]) NC(g_nccl.Send(to[1], (size_t)counts[1] * record_bytes, kNcclUint8, r + 1, m->comm, s));
        if (recv[0]) NC(g_nccl.Recv(from[0], (size_t)recv[0] * record_bytes, kNcclUint8, r - 1, m->comm, s));
        if (recv[1]) NC(g_nccl.Recv(from[1], (size_t)recv[1] * record_bytes, kNcclUint8, r + 1, m->comm, s));
        NC(g_nccl.GroupEnd());
    }
    m->bytes_sent += ((uint64_t)(has_l ? counts[0] : 0) + (has_r ? counts[1] : 0)) * record_bytes;
    return PS_OK;
}

// One whole step of this rank's slab == SlabDomain.step (particlesolver_b200/slab.py); every rank of the communicator calls it.
extern "C" int ps_comm_step(PsCtx *c, float dt) {
    if (!c || !c->comm || !c->comm->slab_set) { ps_set_error("ps_comm_step: no communicator / slab (ps_comm_init, ps_comm_set_slab)"); return PS_ERR_STATE; }
    PsComm *m = c->comm;
    if (c->contact_sources != m->contact_sources_seen || c->nonfluid_sources != m->nonfluid_sources_seen) {
        ps_set_error("ps_comm_step: particles or phases were added after ps_comm_set_slab; call it again on every rank (it takes the global phase census)");
        return PS_ERR_STATE;
    }
    DevGuard dg(c->device);
    uint32_t cnt[2], rcv[2];
    OK(ps_begin_step(c));
    OK(ps_predict(c, dt));
    if (m->recut_every && m->nranks > 1 && m->steps > 0 && m->steps % m->recut_every == 0) OK(recut(c));
    // ---- migration: owned particles whose predicted x left the slab go to the neighbour with pos / prev / vel / w / rho0 / phase ----
    OK(ps_slab_pack_migrants(c, m->x_lo, m->x_hi, m->migr_send[0], m->migr_send[1], m->migr_cap, cnt));
    m->migrated_out += (uint64_t)cnt[0] + cnt[1];
    OK(exchange(c, m->migr_send, cnt, m->migr_recv, rcv, PS_MIGRANT_RECORD_BYTES, m->migr_cap, false));
    OK(ps_slab_append_migrants(c, rcv[0] ? m->migr_recv[0] : nullptr, rcv[0], rcv[1] ? m->migr_recv[1] : nullptr, rcv[1]));
    PsParams p;
    OK(ps_get_params(c, &p));
    for (uint32_t it = 0; it < p.solver_iterations; it++) {
        // ---- ghost halo: positions change every solver iteration ----
        OK(ps_slab_pack_halo(c, m->x_lo, m->x_hi, m->halo, m->halo_send[0], m->halo_send[1], m->halo_cap, cnt));
        OK(exchange(c, m->halo_send, cnt, m->halo_recv, rcv, PS_HALO_RECORD_BYTES, m->halo_cap, false));
        m->ghost_counts[0] = rcv[0]; m->ghost_counts[1] = rcv[1];
        m->ghosts = (uint64_t)rcv[0] + rcv[1];
        OK(ps_slab_set_ghosts(c, rcv[0] ? m->halo_recv[0] : nullptr, rcv[0], rcv[1] ? m->halo_recv[1] : nullptr, rcv[1]));
        OK(ps_build_grid(c));
        if (m->any_contact) OK(ps_solve_contacts(c));
        if (m->exchange_lambda) {
            OK(ps_solve_fluid_lambda(c));
            // ---- ghost lambdas from their owners: one float per halo record, same order, counts known on both sides ----
            OK(ps_slab_pack_lambda(c, m->lam_send[0], m->lam_send[1], m->halo_cap, cnt));
            rcv[0] = m->ghost_counts[0]; rcv[1] = m->ghost_counts[1];
            OK(exchange(c, m->lam_send, cnt, m->lam_recv, rcv, sizeof(float), m->halo_cap, true));
            OK(ps_slab_set_ghost_lambda(c, rcv[0] ? m->lam_recv[0] : nullptr, rcv[0], rcv[1] ? m->lam_recv[1] : nullptr, rcv[1]));
            OK(ps_solve_fluid_delta(c));
        } else {
            OK(ps_solve_fluid(c));
        }
        OK(ps_collide_world(c, it));
    }
    OK(ps_update_velocity(c, dt));
    m->steps++;
    return PS_OK;
}

// the cut planes in force (nranks + 1 floats) and the number of re-cuts so far
extern "C" int ps_comm_get_cuts(PsCtx *c, float *cuts, uint32_t *recuts) {
    if (!c || !c->comm) { ps_set_error("ps_comm_get_cuts: no communicator"); return PS_ERR_STATE; }
    PsComm *m = c->comm;
    if (cuts) {
        if (m->cuts.empty()) { for (int r = 0; r <= m->nranks; r++) cuts[r] = r == m->rank ? m->x_lo : (r == m->rank + 1 ? m->x_hi : NAN); }
        else for (int r = 0; r <= m->nranks; r++) cuts[r] = (float)m->cuts[r];
    }
    if (recuts) *recuts = m->recuts;
    return PS_OK;
}

// out[4] = particles this rank handed to its neighbours so far, ghosts it held in the last iteration, payload bytes it sent, steps
extern "C" int ps_comm_stats(PsCtx *c, uint64_t out[4]) {
    if (!c || !c->comm || !out) { ps_set_error("ps_comm_stats: no communicator"); return PS_ERR_STATE; }
    out[0] = c->comm->migrated_out; out[1] = c->comm->ghosts; out[2] = c->comm->bytes_sent; out[3] = c->comm->steps;
    return PS_OK;
}

// element-wise sum over the ranks of up to 8 doubles (particle counts, energies: the global figures of a decomposed run); blocking
extern "C" int ps_comm_allreduce_sum(PsCtx *c, double *values, uint32_t count) {
    if (!c || !c->comm || !values || count > 8) { ps_set_error("ps_comm_allreduce_sum: bad arguments"); return PS_ERR_INVALID; }
    if (!count) return PS_OK;
    PsComm *m = c->comm;
    DevGuard dg(c->device);
    CC(cudaMemcpyAsync(m->reduce_dev, values, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(m->reduce_dev, m->reduce_dev, count, kNcclFloat64, kNcclSum, m->comm, c->stream));
    CC(cudaMemcpyAsync(values, m->reduce_dev, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CC(cudaStreamSynchronize(c->stream));
    return PS_OK;
}
