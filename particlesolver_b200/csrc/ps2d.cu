// particlesolver_b200/csrc/ps2d.cu — the 2-D double-precision path (include/psolver2d.h): one Simulation::tick of the
// reference's CPU application on the GPU, for all-fluid scenes (config C1).  Compiled WITHOUT -use_fast_math and with
// -fmad=false: the reference is plain x86-64 double arithmetic without contraction, and a tick is compared with it
// to 1e-9.  Sums over neighbours run sequentially in ascending particle index, like the reference's O(N^2) loops
// (cpu/src/constraint/totalfluidconstraint.cpp:52-76), with the other particles staged through shared memory in tiles.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/psolver.h"
#include "../../include/psolver2d.h"

void ps_set_error(const char *fmt, ...);

namespace {
typedef uint32_t u32;
constexpr int kBlock = 128;
constexpr int kTile = 128;
constexpr double kRad = 0.25;    // PARTICLE_RAD, cpu/src/particle.h:6
constexpr double kEps = 1e-4;    // EPSILON, cpu/src/includes.h:34
constexpr double kH = 2., kH2 = 4., kH6 = 64., kH9 = 512.;  // totalfluidconstraint.h:16-19
constexpr double kRelax = .01, kKP = .1, kEP = 4., kDQ = .2;   // totalfluidconstraint.h:22-27
constexpr double kPi = 3.14159265358979323846;

__device__ __forceinline__ double poly6(double r2) {  // totalfluidconstraint.cpp:121-127
    if (r2 >= kH2) return 0.;
    const double term2 = kH2 - r2;
    return (315. / (64. * kPi * kH9)) * (term2 * term2 * term2);
}
// spikyGrad(r, rlen) = -normalize(r) * (45 / (pi H^6)) * (H - rlen)^2, zero outside the support and at r = 0 (:129-135);
// glm::normalize(v) = v * (1 / sqrt(dot(v, v)))
__device__ __forceinline__ double2 spiky_grad(double rx, double ry, double rlen) {
    if (rlen >= kH || rlen == 0.) return make_double2(0., 0.);
    const double inv = 1. / rlen;
    const double c = 45. / (kPi * kH6), hm = kH - rlen;
    return make_double2((-(rx * inv) * c) * hm * hm, (-(ry * inv) * c) * hm * hm);
}

// (1)-(4): v += dt g; ep = p + dt v (fixed particles stay); simulation.cpp:139-161, particle.h:56-58
__global__ void k2d_predict(double2 *__restrict__ v, double2 *__restrict__ ep, const double2 *__restrict__ p, const double *__restrict__ imass, u32 n,
                            double dt, double gx, double gy) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    double2 vi = v[i];
    vi.x = vi.x + dt * gx;
    vi.y = vi.y + dt * gy;
    v[i] = vi;
    const double2 pi = p[i];
    ep[i] = imass[i] == 0. ? pi : make_double2(pi.x + dt * vi.x, pi.y + dt * vi.y);
}

// (8): which walls a predicted position violates — at most one constraint per axis, x before y (simulation.cpp:202-224).
// flags: bit0 x-low, bit1 x-high, bit2 y-low, bit3 y-high; counts[i] = number of constraints (BoundaryConstraint::updateCounts).
__global__ void k2d_boundary_flags(const double2 *__restrict__ ep, u32 n, double x0, double x1, double y0, double y1, u32 *__restrict__ flags,
                                   u32 *__restrict__ counts) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const double2 e = ep[i];
    u32 f = 0;
    if (e.x < x0 + kRad) f |= 1u; else if (e.x > x1 - kRad) f |= 2u;
    if (e.y < y0 + kRad) f |= 4u; else if (e.y > y1 - kRad) f |= 8u;
    flags[i] = f;
    counts[i] = __popc(f);
}

// rank[i] = number of boundary constraints of particles before i = position of i's first constraint in the reference's
// constraint list (and so in the rand() stream of an iteration); total -> *num.  One CTA, sequential over chunks.
__global__ void __launch_bounds__(1024) k2d_scan_counts(const u32 *__restrict__ counts, u32 *__restrict__ rank, u32 n, u32 *__restrict__ num) {
    __shared__ u32 sm[32];
    __shared__ u32 carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        const u32 i = base + threadIdx.x;
        const u32 c = i < n ? counts[i] : 0u;
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm[wid] = incl;
        __syncthreads();
        u32 before = carry, tot = 0;
        for (int w = 0; w < 32; w++) {
            if (w < wid) before += sm[w];
            tot += sm[w];
        }
        if (i < n) rank[i] = before + incl - c;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *num = carry;
}

// BoundaryConstraint::project for every constraint of one solver iteration (boundaryconstraint.cpp:14-93).  A constraint
// touches one coordinate of one particle, so the list order only decides which draw of the stream it gets:
// draw = raw[iteration * num + position in the list], extra = (double)(float)(draw / RAND_MAX) * .003 (frand() is
// float-typed, includes.h:25), consumed before the early-out.  Friction is a no-op for fluids (sFriction = kFriction = 0).
__global__ void k2d_boundary_project(double2 *__restrict__ ep, const u32 *__restrict__ flags, const u32 *__restrict__ rank, u32 n, const int *__restrict__ raw,
                                     u32 draw_base, double x0, double x1, double y0, double y1) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const u32 f = flags[i];
    if (!f) return;
    double2 e = ep[i];
    u32 k = draw_base + rank[i];
    auto extra = [&](u32 idx) { return (double)(float)((double)raw[idx] / 2147483647.0) * .003; };
    if (f & 3u) {
        const double d = kRad + extra(k++);
        if (f & 1u) { if (!(e.x >= x0 + kRad)) e.x = x0 + d; }
        else { if (!(e.x <= x1 - kRad)) e.x = x1 - d; }
    }
    if (f & 12u) {
        const double d = kRad + extra(k++);
        if (f & 4u) { if (!(e.y >= y0 + kRad)) e.y = y0 + d; }
        else { if (!(e.y <= y1 - kRad)) e.y = y1 - d; }
    }
    ep[i] = e;
}

// TotalFluidConstraint::project, first loop (totalfluidconstraint.cpp:45-93): lambda of every particle of fluid `f`,
// 0 for everybody else (the constraint's lambdas is a QHash cleared per call: the other fluid reads 0, :106).
__global__ void __launch_bounds__(kBlock) k2d_fluid_lambda(const double2 *__restrict__ ep, const double *__restrict__ imass, const int *__restrict__ fluid,
                                                           u32 n, int f, double p0, double *__restrict__ lambda, u32 *__restrict__ nbcount) {
    __shared__ double2 s_ep[kTile];
    __shared__ double s_im[kTile];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    const bool mine = i < n && fluid[i] == f;
    const double2 pi = i < n ? ep[i] : make_double2(0., 0.);
    double rho = 0., denom = 0., ox = 0., oy = 0.;
    u32 nb = 0;
    for (u32 base = 0; base < n; base += kTile) {
        __syncthreads();
        if (base + threadIdx.x < n && threadIdx.x < kTile) { s_ep[threadIdx.x] = ep[base + threadIdx.x]; s_im[threadIdx.x] = imass[base + threadIdx.x]; }
        __syncthreads();
        if (!mine) continue;
        const u32 cnt = min((u32)kTile, n - base);
        for (u32 t = 0; t < cnt; t++) {
            const u32 j = base + t;
            if (j == i) {  // the particle itself, at its place in the index order (:78-81)
                nb++;
                rho += poly6(0.) / s_im[t];
                continue;
            }
            if (s_im[t] == 0.) continue;  // fixed particles are ignored
            const double rx = pi.x - s_ep[t].x, ry = pi.y - s_ep[t].y;
            const double r2 = rx * rx + ry * ry;
            if (r2 < kH2) {
                nb++;
                rho += poly6(r2) / s_im[t];
                const double2 sg = spiky_grad(rx, ry, sqrt(r2));
                const double gx = -sg.x / p0, gy = -sg.y / p0;  // grad(k, j) = -spikyGrad / p0 (:137-144)
                denom += gx * gx + gy * gy;
                ox += sg.x; oy += sg.y;                         // grad(k, i) = sum_j spikyGrad / p0 (:146-157)
            }
        }
    }
    if (i >= n) return;
    if (!mine) { lambda[i] = 0.; return; }
    ox = ox / p0; oy = oy / p0;
    denom += ox * ox + oy * oy;
    lambda[i] = -((rho / p0) - 1.) / (denom + kRelax);
    nbcount[i] = nb;
}

// second loop (:95-111): delta_i = sum_j (lambda_i + lambda_j + s_corr) spikyGrad / p0, divided by (#neighbours incl. self +
// boundary count) (:113-115).  Written to `delta`, applied by k2d_fluid_apply: all deltas of a fluid come from the same ep.
__global__ void __launch_bounds__(kBlock) k2d_fluid_delta(const double2 *__restrict__ ep, const double *__restrict__ imass, const int *__restrict__ fluid,
                                                          u32 n, int f, double p0, const double *__restrict__ lambda, const u32 *__restrict__ nbcount,
                                                          const u32 *__restrict__ counts, double2 *__restrict__ delta) {
    __shared__ double2 s_ep[kTile];
    __shared__ double s_im[kTile], s_lam[kTile];
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    const bool mine = i < n && fluid[i] == f;
    const double2 pi = i < n ? ep[i] : make_double2(0., 0.);
    const double li = mine ? lambda[i] : 0.;
    const double base6 = poly6(kDQ * kDQ * kH * kH);
    double dx = 0., dy = 0.;
    for (u32 base = 0; base < n; base += kTile) {
        __syncthreads();
        if (base + threadIdx.x < n && threadIdx.x < kTile) {
            s_ep[threadIdx.x] = ep[base + threadIdx.x]; s_im[threadIdx.x] = imass[base + threadIdx.x]; s_lam[threadIdx.x] = lambda[base + threadIdx.x];
        }
        __syncthreads();
        if (!mine) continue;
        const u32 cnt = min((u32)kTile, n - base);
        for (u32 t = 0; t < cnt; t++) {
            const u32 j = base + t;
            if (j == i || s_im[t] == 0.) continue;
            const double rx = pi.x - s_ep[t].x, ry = pi.y - s_ep[t].y;
            const double r2 = rx * rx + ry * ry;
            if (r2 < kH2) {
                const double rlen = sqrt(r2);
                const double2 sg = spiky_grad(rx, ry, rlen);
                const double corr = -kKP * pow(poly6(rlen * rlen) / base6, kEP);
                const double s = (li + s_lam[t]) + corr;
                dx += s * sg.x; dy += s * sg.y;
            }
        }
    }
    if (!mine) return;
    const double div = (double)nbcount[i] + (double)counts[i];
    delta[i] = make_double2((dx / p0) / div, (dy / p0) / div);
}
__global__ void k2d_fluid_apply(double2 *__restrict__ ep, const double2 *__restrict__ delta, const int *__restrict__ fluid, u32 n, int f) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n || fluid[i] != f) return;
    double2 e = ep[i];
    const double2 d = delta[i];
    e.x += d.x; e.y += d.y;
    ep[i] = e;
}

// (23)-(27): v = (ep - p) / dt; particles that moved less than EPSILON sleep (particle.h:60-65)
__global__ void k2d_finish(double2 *__restrict__ p, double2 *__restrict__ v, const double2 *__restrict__ ep, u32 n, double dt) {
    const u32 i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const double2 pi = p[i], e = ep[i];
    const double dx = e.x - pi.x, dy = e.y - pi.y;
    if (sqrt(dx * dx + dy * dy) < kEps) { v[i] = make_double2(0., 0.); return; }
    v[i] = make_double2(dx / dt, dy / dt);
    p[i] = e;
}

// glibc rand() = random(), TYPE_3: r[i] = r[i-31] + r[i-3], output r[i] >> 1 (after 310 discarded words)
struct GlibcRand {
    std::vector<uint32_t> r;
    size_t pos = 0;
    uint64_t calls = 0;
    void seed(uint32_t s) {
        r.assign(344, 0);
        r[0] = s ? s : 1;
        for (int i = 1; i < 31; i++) {
            const long hi = (long)r[i - 1] / 127773, lo = (long)r[i - 1] % 127773;
            long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            r[i] = (uint32_t)w;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
        r.erase(r.begin(), r.end() - 31);
        calls = 0;
    }
    int next() {
        const uint32_t v = r[0] + r[28];
        r.erase(r.begin());
        r.push_back(v);
        calls++;
        return (int)(v >> 1);
    }
};
}  // namespace

struct Ps2dCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Ps2dParams params{};
    uint64_t cap = 0;
    u32 n = 0;
    double2 *p = nullptr, *v = nullptr, *ep = nullptr, *delta = nullptr;
    double *imass = nullptr, *lambda = nullptr;
    int *fluid = nullptr, *raw = nullptr;
    u32 *flags = nullptr, *counts = nullptr, *rank = nullptr, *nbcount = nullptr, *num_dev = nullptr, *num_host = nullptr;
    size_t raw_cap = 0;
    std::vector<double> rho0;
    std::vector<int> h_raw;
    GlibcRand rng;
    u32 last_num_boundary = 0, launches = 0;
};

#define CU2(x)                                                                                        \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            ps_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return PS_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

extern "C" void ps2d_default_params(Ps2dParams *p) {
    if (!p) return;
    p->x_bounds[0] = -8.; p->x_bounds[1] = 8.;    // scene 6, cpu/src/simulation.cpp:898-899
    p->y_bounds[0] = -8.; p->y_bounds[1] = 40.;
    p->gravity[0] = 0.; p->gravity[1] = -9.8;
    p->solver_iterations = 3;                     // simulation.h:11
}

extern "C" int ps2d_create(int device, const Ps2dParams *params, uint64_t max_particles, Ps2dCtx **out) {
    if (!params || !out || !max_particles) { ps_set_error("ps2d_create: bad argument"); return PS_ERR_INVALID; }
    *out = nullptr;
    if (!(params->x_bounds[0] < params->x_bounds[1]) || !(params->y_bounds[0] < params->y_bounds[1]) || params->solver_iterations > 64) {
        ps_set_error("ps2d_create: bad bounds or iteration count"); return PS_ERR_INVALID;
    }
    int ndev = 0;
    CU2(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { ps_set_error("ps2d_create: device %d of %d", device, ndev); return PS_ERR_INVALID; }
    cudaDeviceProp prop;
    CU2(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { ps_set_error("ps2d_create: device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor); return PS_ERR_CUDA; }
    CU2(cudaSetDevice(device));
    Ps2dCtx *c = new Ps2dCtx();
    c->device = device; c->params = *params; c->cap = max_particles;
    c->rng.seed(1);
    const size_t n = max_particles;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&c->p, n * 16) != cudaSuccess ||
        cudaMalloc(&c->v, n * 16) != cudaSuccess || cudaMalloc(&c->ep, n * 16) != cudaSuccess || cudaMalloc(&c->delta, n * 16) != cudaSuccess ||
        cudaMalloc(&c->imass, n * 8) != cudaSuccess || cudaMalloc(&c->lambda, n * 8) != cudaSuccess || cudaMalloc(&c->fluid, n * 4) != cudaSuccess ||
        cudaMalloc(&c->flags, n * 4) != cudaSuccess || cudaMalloc(&c->counts, n * 4) != cudaSuccess || cudaMalloc(&c->rank, n * 4) != cudaSuccess ||
        cudaMalloc(&c->nbcount, n * 4) != cudaSuccess || cudaMalloc(&c->num_dev, 4) != cudaSuccess || cudaMallocHost(&c->num_host, 4) != cudaSuccess) {
        ps_set_error("ps2d_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        ps2d_destroy(c);
        return PS_ERR_CUDA;
    }
    *out = c;
    return PS_OK;
}

extern "C" int ps2d_destroy(Ps2dCtx *c) {
    if (!c) return PS_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void *ptrs[] = {c->p, c->v, c->ep, c->delta, c->imass, c->lambda, c->fluid, c->raw, c->flags, c->counts, c->rank, c->nbcount, c->num_dev};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (c->num_host) cudaFreeHost(c->num_host);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PS_OK;
}

extern "C" int ps2d_create_fluid(Ps2dCtx *c, const double *p2, const double *v2, const double *inv_mass, uint64_t n, double density) {
    if (!c || !p2 || !v2 || !inv_mass) { ps_set_error("ps2d_create_fluid: null argument"); return PS_ERR_INVALID; }
    if (c->n + n > c->cap) { ps_set_error("ps2d_create_fluid: %llu + %llu exceeds max_particles", (unsigned long long)c->n, (unsigned long long)n); return PS_ERR_CAPACITY; }
    if (!(density > 0.)) { ps_set_error("ps2d_create_fluid: density must be positive"); return PS_ERR_INVALID; }
    for (uint64_t k = 0; k < n; k++)
        if (inv_mass[k] == 0.0) { ps_set_error("A fluid cannot have a point of infinite mass."); return PS_ERR_INVALID; }  // simulation.cpp:441-444
    CU2(cudaSetDevice(c->device));
    std::vector<int> fl(n, (int)c->rho0.size());
    CU2(cudaMemcpyAsync(c->p + c->n, p2, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU2(cudaMemcpyAsync(c->v + c->n, v2, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU2(cudaMemcpyAsync(c->ep + c->n, p2, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU2(cudaMemcpyAsync(c->imass + c->n, inv_mass, n * 8, cudaMemcpyHostToDevice, c->stream));
    CU2(cudaMemcpyAsync(c->fluid + c->n, fl.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    CU2(cudaStreamSynchronize(c->stream));
    c->n += (u32)n;
    c->rho0.push_back(density);
    return PS_OK;
}

extern "C" int ps2d_seed_rand(Ps2dCtx *c, uint32_t seed, uint64_t skip) {
    if (!c) return PS_ERR_INVALID;
    c->rng.seed(seed);
    for (uint64_t k = 0; k < skip; k++) c->rng.next();
    return PS_OK;
}
extern "C" uint64_t ps2d_rand_calls(Ps2dCtx *c) { return c ? c->rng.calls : 0; }
extern "C" uint64_t ps2d_num_particles(Ps2dCtx *c) { return c ? c->n : 0; }
extern "C" uint32_t ps2d_last_num_boundary_constraints(Ps2dCtx *c) { return c ? c->last_num_boundary : 0; }
extern "C" uint32_t ps2d_launches_per_tick(Ps2dCtx *c) { return c ? c->launches : 0; }

extern "C" int ps2d_tick(Ps2dCtx *c, double dt) {
    if (!c) { ps_set_error("null context"); return PS_ERR_INVALID; }
    if (!c->n) return PS_OK;
    CU2(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const u32 n = c->n, blocks = (n + kBlock - 1) / kBlock;
    const Ps2dParams &P = c->params;
    u32 launches = 0;
    k2d_predict<<<blocks, kBlock, 0, s>>>(c->v, c->ep, c->p, c->imass, n, dt, P.gravity[0], P.gravity[1]);
    k2d_boundary_flags<<<blocks, kBlock, 0, s>>>(c->ep, n, P.x_bounds[0], P.x_bounds[1], P.y_bounds[0], P.y_bounds[1], c->flags, c->counts);
    k2d_scan_counts<<<1, 1024, 0, s>>>(c->counts, c->rank, n, c->num_dev);
    launches += 3;
    // the number of wall constraints decides how many draws of the rand() stream this tick consumes: one per constraint
    // per solver iteration, in list order (the only host round trip of a tick)
    CU2(cudaMemcpyAsync(c->num_host, c->num_dev, 4, cudaMemcpyDeviceToHost, s));
    CU2(cudaStreamSynchronize(s));
    const u32 num = *c->num_host;
    c->last_num_boundary = num;
    const size_t draws = (size_t)num * P.solver_iterations;
    if (draws) {
        c->h_raw.resize(draws);
        for (size_t k = 0; k < draws; k++) c->h_raw[k] = c->rng.next();
        if (draws > c->raw_cap) {
            if (c->raw) CU2(cudaFree(c->raw));
            CU2(cudaMalloc(&c->raw, draws * 2 * sizeof(int)));
            c->raw_cap = draws * 2;
        }
        CU2(cudaMemcpyAsync(c->raw, c->h_raw.data(), draws * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    for (u32 it = 0; it < P.solver_iterations; it++) {
        if (num) {
            k2d_boundary_project<<<blocks, kBlock, 0, s>>>(c->ep, c->flags, c->rank, n, c->raw, it * num, P.x_bounds[0], P.x_bounds[1], P.y_bounds[0], P.y_bounds[1]);
            launches++;
        }
        for (size_t f = 0; f < c->rho0.size(); f++) {
            k2d_fluid_lambda<<<blocks, kBlock, 0, s>>>(c->ep, c->imass, c->fluid, n, (int)f, c->rho0[f], c->lambda, c->nbcount);
            k2d_fluid_delta<<<blocks, kBlock, 0, s>>>(c->ep, c->imass, c->fluid, n, (int)f, c->rho0[f], c->lambda, c->nbcount, c->counts, c->delta);
            k2d_fluid_apply<<<blocks, kBlock, 0, s>>>(c->ep, c->delta, c->fluid, n, (int)f);
            launches += 3;
        }
    }
    k2d_finish<<<blocks, kBlock, 0, s>>>(c->p, c->v, c->ep, n, dt);
    launches++;
    c->launches = launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ps_set_error("ps2d_tick: %s", cudaGetErrorString(e)); return PS_ERR_CUDA; }
    CU2(cudaStreamSynchronize(s));  // h_raw is reused by the next tick
    return PS_OK;
}

extern "C" int ps2d_download(Ps2dCtx *c, int which, double *host) {
    if (!c || !host) { ps_set_error("ps2d_download: null argument"); return PS_ERR_INVALID; }
    CU2(cudaSetDevice(c->device));
    const void *src = nullptr;
    size_t bytes = (size_t)c->n * 16;
    switch (which) {
        case PS2D_ARR_P: src = c->p; break;
        case PS2D_ARR_V: src = c->v; break;
        case PS2D_ARR_EP: src = c->ep; break;
        case PS2D_ARR_LAMBDA: src = c->lambda; bytes = (size_t)c->n * 8; break;
        default: ps_set_error("ps2d_download: unknown array %d", which); return PS_ERR_INVALID;
    }
    CU2(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU2(cudaStreamSynchronize(c->stream));
    return PS_OK;
}

extern "C" int ps2d_kinetic_energy(Ps2dCtx *c, double *out) {
    if (!c || !out) return PS_ERR_INVALID;
    std::vector<double> v((size_t)c->n * 2), im(c->n);
    int r = ps2d_download(c, PS2D_ARR_V, v.data());
    if (r != PS_OK) return r;
    CU2(cudaMemcpy(im.data(), c->imass, (size_t)c->n * 8, cudaMemcpyDeviceToHost));
    double e = 0;  // same order as Simulation::getKineticEnergy
    for (u32 i = 0; i < c->n; i++)
        if (im[i] != 0.) e += .5 * (v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]) / im[i];
    *out = e;
    return PS_OK;
}
