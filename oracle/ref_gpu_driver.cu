// Headless driver around the reference's UNMODIFIED GPU solver (gpu/src/particlesystem.cpp +
// gpu/src/cuda/{integration,solver,shared_variables}.cu, compiled from /root/reference by
// oracle/Makefile into oracle/_ref/ref_gpu).  TEST INFRASTRUCTURE ONLY.
//
// It builds one of the reference's scenes through the reference's own ParticleSystem builders
// (scene table: gpu/src/particleapp.cpp:141-215), then either
//   --mode staged : replays ParticleSystem::update (gpu/src/particlesystem.cpp:144-246) call by call
//                   through the reference's extern "C" wrappers, dumping every array after every
//                   wrapper call  -> golden vectors for the per-stage parity tests, or
//   --mode whole  : calls ParticleSystem::update(dt) itself N times, dumping positions/velocities
//                   at chosen steps and timing the step with CUDA events.
// Output: raw little-endian arrays <out>/<name>.bin + <out>/manifest.txt ("name dtype count").
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <sys/resource.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#define private public
#include "particlesystem.h"
#undef private
#include "wrappers.cuh"
#include "util.cuh"
#include "shared_variables.cuh"
#include "../include/ps_scenes.h"

#ifdef PS_DROPIN
// Drop-in build (oracle/_ref/ref_host_on_psolver): the reference's UNMODIFIED particlesystem.cpp and this driver
// link against libpsolver.so instead of the reference's integration.cu / solver.cu / shared_variables.cu.  The
// state those files keep in file-scope thrust vectors is reached through the shim's accessor functions.
#define PS_REFERENCE_ABI_NO_PROTOTYPES
#include "../include/ps_reference_abi.h"
static float *st_vel() { return psRefVelocityPtr(); }
static float *st_lambda() { return psRefLambdaPtr(); }
static float *st_ros() { return psRefRestDensityPtr(); }
static uint *st_nn() { return psRefNumNeighborsPtr(); }
static float *st_rands() { return psRefRandsPtr(); }
static uint *st_occ() { return psRefOccurrencesPtr(); }
static size_t st_num_dist() { return psRefNumDistanceConstraints(); }
static size_t st_num_points() { return psRefNumPointConstraints(); }
#else
// file-scope state of the reference's translation units (integration.cu:23-34, solver.cu:32-41)
extern thrust::device_vector<float> V, lambda, ros;
extern thrust::device_vector<uint> numNeighbors;
extern thrust::device_vector<uint> distsI, pointsI, occurences;
extern thrust::device_vector<float> dists, points;
extern float *rands;
template <class T> static T *raw(thrust::device_vector<T> &v) { return thrust::raw_pointer_cast(v.data()); }
static float *st_vel() { return raw(V); }
static float *st_lambda() { return raw(lambda); }
static float *st_ros() { return raw(ros); }
static uint *st_nn() { return raw(numNeighbors); }
static float *st_rands() { return rands; }
static uint *st_occ() { return raw(occurences); }
static size_t st_num_dist() { return dists.size(); }
static size_t st_num_points() { return pointsI.size(); }
#endif

#ifdef PS_DROPIN
#define IMPL_NAME "reference_host_on_libpsolver"
#else
#define IMPL_NAME "reference_gpu_unmodified"
#endif
static std::string g_out;
static FILE *g_manifest = nullptr;

static void dump_dev(const std::string &name, const void *dptr, size_t count, const char *dtype, size_t esz) {
    std::vector<char> h(count * esz);
    if (count) cudaMemcpy(h.data(), dptr, count * esz, cudaMemcpyDeviceToHost);
    FILE *f = fopen((g_out + "/" + name + ".bin").c_str(), "wb");
    if (!f) { perror("fopen"); exit(1); }
    fwrite(h.data(), 1, h.size(), f);
    fclose(f);
    fprintf(g_manifest, "%s %s %zu\n", name.c_str(), dtype, count);
}
static void dump_host(const std::string &name, const void *h, size_t count, const char *dtype, size_t esz) {
    FILE *f = fopen((g_out + "/" + name + ".bin").c_str(), "wb");
    if (!f) { perror("fopen"); exit(1); }
    fwrite(h, 1, count * esz, f);
    fclose(f);
    fprintf(g_manifest, "%s %s %zu\n", name.c_str(), dtype, count);
}
static void dump_f(const std::string &n, const float *d, size_t c) { dump_dev(n, d, c, "f32", 4); }
static void dump_u(const std::string &n, const uint *d, size_t c) { dump_dev(n, d, c, "u32", 4); }
static void dump_i(const std::string &n, const int *d, size_t c) { dump_dev(n, d, c, "i32", 4); }
struct Args {
    std::string scene = "7", mode = "staged", out = "gpurun_out/ref_gpu";
    int grid = 64, steps = 1, dump_every = 1, iters = 5, side = 100;
    int dump_iters = 1 << 30;  // staged mode: dump the arrays of the first dump_iters solver iterations of a step only
    int light = 0;             // staged mode: only what the full-size parity test reads (1M particles: 0.25 GB instead of 1.3 GB per step)
    int warmup = -1;           // whole mode: untimed leading steps (default: the first step when more than one is run)
    unsigned max_particles = 15000;
    float dt = 1.0f / 60.0f;
};

// the scene scripts are shared with the product's host class (include/ps_scenes.h) so that both sides build
// literally the same scene through their own builders
static ParticleSystem *build_scene(const Args &a) {
    ps_scenes::SceneSpec s;
    s.scene = a.scene; s.grid = a.grid; s.max_particles = a.max_particles; s.iterations = a.iters; s.side = a.side;
    ParticleSystem *ps = ps_scenes::build<ParticleSystem>(s, colors, numColors);
    if (!ps) { fprintf(stderr, "unknown scene %s\n", a.scene.c_str()); exit(2); }
    return ps;
}

static void dump_scene(ParticleSystem *ps) {
    uint n = ps->m_numParticles;
    float *dPos = (float *)mapGLBufferObject(&ps->m_cuda_posvbo_resource);
    dump_f("init_pos", dPos, 4 * (size_t)n);
    dump_f("init_vel", st_vel(), 4 * (size_t)n);
    dump_f("init_w", getWRawPtr(), n);
    dump_i("init_phase", getPhaseRawPtr(), n);
    dump_f("init_ros", st_ros(), n);
#ifdef PS_DROPIN
    {
        std::vector<uint> di(2 * st_num_dist()), pi(st_num_points());
        std::vector<float> dr(st_num_dist()), px(3 * st_num_points());
        psRefCopyDistanceConstraints(di.data(), dr.data());
        psRefCopyPointConstraints(pi.data(), px.data());
        dump_host("dist_idx", di.data(), di.size(), "u32", 4);
        dump_host("dist_rest", dr.data(), dr.size(), "f32", 4);
        dump_host("point_idx", pi.data(), pi.size(), "u32", 4);
        dump_host("point_xyz", px.data(), px.size(), "f32", 4);
    }
#else
    dump_u("dist_idx", raw(distsI), distsI.size());
    dump_f("dist_rest", raw(dists), dists.size());
    dump_u("point_idx", raw(pointsI), pointsI.size());
    dump_f("point_xyz", raw(points), points.size());
#endif
    dump_u("occurences", st_occ(), n);
    FILE *f = fopen((g_out + "/scene.txt").c_str(), "w");
    fprintf(f, "n %u\nradius %g\ngrid %u %u %u\nmin %d %d %d\nmax %d %d %d\niters %u\n", n, ps->m_particleRadius,
            ps->m_gridSize.x, ps->m_gridSize.y, ps->m_gridSize.z, ps->m_minBounds.x, ps->m_minBounds.y, ps->m_minBounds.z,
            ps->m_maxBounds.x, ps->m_maxBounds.y, ps->m_maxBounds.z, ps->m_solverIterations);
    fclose(f);
}

// call-by-call replay of ParticleSystem::update (particlesystem.cpp:144-246) with dumps in between
static void staged_step(ParticleSystem *ps, float dt, int step, bool dump_step, int dump_iters, bool light) {
    bool dump = dump_step;
    dt = std::min(dt, .05f);
    uint n = ps->m_numParticles, cells = ps->m_numGridCells;
    float *dPos = (float *)mapGLBufferObject(&ps->m_cuda_posvbo_resource);
    setParameters(&ps->m_params);
    integrateSystem(dPos, dt, n);
    char tag[64];
    snprintf(tag, sizeof tag, "s%d_", step);
    std::string S(tag);
    if (dump) { dump_f(S + "predict_pos", dPos, 4 * (size_t)n); if (!light) dump_f(S + "prev", getXstarRawPtr(), 4 * (size_t)n); }
    for (uint i = 0; i < ps->m_solverIterations; i++) {
        dump = dump_step && (int)i < dump_iters;
        snprintf(tag, sizeof tag, "s%d_i%u_", step, i);
        std::string T(tag);
        calcHash(ps->m_dGridParticleHash, ps->m_dGridParticleIndex, dPos, n);
        if (dump && !light) dump_u(T + "hash_unsorted", ps->m_dGridParticleHash, n);
        sortParticles(ps->m_dGridParticleHash, ps->m_dGridParticleIndex, n);
        if (dump) { dump_u(T + "hash", ps->m_dGridParticleHash, n); dump_u(T + "index", ps->m_dGridParticleIndex, n); }
        reorderDataAndFindCellStart(ps->m_dCellStart, ps->m_dCellEnd, ps->m_dSortedPos, ps->m_dSortedW, ps->m_dSortedPhase,
                                    ps->m_dGridParticleHash, ps->m_dGridParticleIndex, dPos, n, cells);
        if (dump) {
            dump_u(T + "cell_start", ps->m_dCellStart, cells);
            dump_u(T + "cell_end", ps->m_dCellEnd, cells);
            dump_f(T + "sorted_pos", ps->m_dSortedPos, 4 * (size_t)n);
            if (!light) { dump_f(T + "sorted_w", ps->m_dSortedW, n); dump_i(T + "sorted_phase", ps->m_dSortedPhase, n); }
        }
        collide(dPos, ps->m_dSortedPos, ps->m_dSortedW, ps->m_dSortedPhase, ps->m_dGridParticleIndex, ps->m_dCellStart,
                ps->m_dCellEnd, n, cells);
        if (dump && !light) { dump_f(T + "collide_pos", dPos, 4 * (size_t)n); dump_u(T + "collide_nn", st_nn(), n); }
        solveFluids(ps->m_dSortedPos, ps->m_dSortedW, ps->m_dSortedPhase, ps->m_dGridParticleIndex, ps->m_dCellStart,
                    ps->m_dCellEnd, dPos, n, cells);
        if (dump) {
            dump_f(T + "lambda", st_lambda(), n);
            dump_u(T + "fluid_nn", st_nn(), n);
            dump_f(T + "fluid_pos", dPos, 4 * (size_t)n);
        }
        collideWorld(dPos, ps->m_dSortedPos, n, ps->m_minBounds, ps->m_maxBounds);
        if (dump && !light) { dump_f(T + "rands", st_rands(), 6); dump_f(T + "world_pos", dPos, 4 * (size_t)n); }
        solveDistanceConstraints(dPos);
        if (dump && !light) dump_f(T + "dist_pos", dPos, 4 * (size_t)n);
        solvePointConstraints(dPos);
        if (dump && !light) dump_f(T + "point_pos", dPos, 4 * (size_t)n);
    }
    calcVelocity(dPos, dt, n);
    dump = dump_step;
    if (dump) { dump_f(S + "final_pos", dPos, 4 * (size_t)n); dump_f(S + "final_vel", st_vel(), 4 * (size_t)n); }
    unmapGLBufferObject(ps->m_cuda_posvbo_resource);
}

int main(int argc, char **argv) {
    // the reference's scene builders keep whole scenes in stack VLAs (particlesystem.cpp:366-370)
    struct rlimit rl;
    getrlimit(RLIMIT_STACK, &rl);
    if (rl.rlim_cur != RLIM_INFINITY && rl.rlim_cur < (rlim_t)1 << 31 && !getenv("REF_GPU_REEXEC")) {
        rl.rlim_cur = rl.rlim_max;
        setrlimit(RLIMIT_STACK, &rl);
        setenv("REF_GPU_REEXEC", "1", 1);
        execv("/proc/self/exe", argv);
    }
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i];
        auto val = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", k.c_str()); exit(2); } return std::string(argv[++i]); };
        if (k == "--scene") a.scene = val();
        else if (k == "--mode") a.mode = val();
        else if (k == "--out") a.out = val();
        else if (k == "--grid") a.grid = atoi(val().c_str());
        else if (k == "--steps") a.steps = atoi(val().c_str());
        else if (k == "--dump-every") a.dump_every = atoi(val().c_str());
        else if (k == "--iters") a.iters = atoi(val().c_str());
        else if (k == "--side") a.side = atoi(val().c_str());
        else if (k == "--max") a.max_particles = (unsigned)atol(val().c_str());
        else if (k == "--dt") a.dt = (float)atof(val().c_str());
        else if (k == "--dump-iters") a.dump_iters = atoi(val().c_str());
        else if (k == "--light") a.light = atoi(val().c_str());
        else if (k == "--warmup") a.warmup = atoi(val().c_str());
        else { fprintf(stderr, "unknown arg %s\n", k.c_str()); return 2; }
    }
    g_out = a.out;
    std::string cmd = "mkdir -p '" + g_out + "'";
    if (system(cmd.c_str()) != 0) return 1;
    g_manifest = fopen((g_out + "/manifest.txt").c_str(), "w");
    cudaInit();
    ParticleSystem *ps = build_scene(a);
    uint n = ps->getNumParticles();
    printf("scene %s: %u particles, %zu distance, %zu point constraints, grid %d^3\n", a.scene.c_str(), n, st_num_dist(),
           st_num_points(), a.grid);
    if (n == 0) { fprintf(stderr, "scene is empty (maxParticles too small?)\n"); return 3; }
    dump_scene(ps);
    if (a.mode == "staged") {
        for (int s = 0; s < a.steps; s++) staged_step(ps, a.dt, s, true, a.dump_iters, a.light != 0);
    } else {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        double total_ms = 0;
        std::vector<float> per_step;
        const int warm = a.warmup >= 0 ? (a.warmup < a.steps ? a.warmup : a.steps - 1) : (a.steps > 1 ? 1 : 0);
        for (int s = 0; s < a.steps; s++) {
            cudaEventRecord(e0);
            ps->update(a.dt);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (s >= warm) { total_ms += ms; per_step.push_back(ms); }
            if (getenv("REF_GPU_VERBOSE")) fprintf(stderr, "step %d: %.3f ms\n", s, ms);
            if (a.dump_every > 0 && ((s + 1) % a.dump_every == 0 || s + 1 == a.steps)) {
                char tag[64]; snprintf(tag, sizeof tag, "w%d_", s + 1);
                float *dPos = (float *)mapGLBufferObject(&ps->m_cuda_posvbo_resource);
                dump_f(std::string(tag) + "pos", dPos, 4 * (size_t)n);
                dump_f(std::string(tag) + "vel", st_vel(), 4 * (size_t)n);
            }
        }
        int timed = a.steps - warm;
        double ms = total_ms / timed;
        // the reference's step time has a heavy tail on any box (per-call cudaMalloc / thrust temporaries, first touch of its 8 KB per
        // particle of neighbour lists): the median is the figure that is fair to it
        std::sort(per_step.begin(), per_step.end());
        double med = per_step.empty() ? ms : (per_step.size() & 1 ? per_step[per_step.size() / 2]
                                                                 : 0.5 * (per_step[per_step.size() / 2 - 1] + per_step[per_step.size() / 2]));
        printf("{\"impl\": \"" IMPL_NAME "\", \"scene\": \"%s\", \"n\": %u, \"grid\": %d, \"steps_timed\": %d, "
               "\"ms_per_step\": %.4f, \"particle_steps_per_s\": %.1f, \"ms_per_step_median\": %.4f, \"ms_per_step_min\": %.4f, \"ms_per_step_max\": %.4f}\n",
               a.scene.c_str(), n, a.grid, timed, ms, n / (ms * 1e-3), med, per_step.empty() ? ms : per_step.front(), per_step.empty() ? ms : per_step.back());
        FILE *f = fopen((g_out + "/timing.json").c_str(), "w");
        fprintf(f, "{\"scene\": \"%s\", \"n\": %u, \"grid\": %d, \"steps_timed\": %d, \"ms_per_step\": %.4f, \"particle_steps_per_s\": %.1f}\n",
                a.scene.c_str(), n, a.grid, timed, ms, n / (ms * 1e-3));
        fclose(f);
    }
    fclose(g_manifest);
    delete ps;
    return 0;
}
