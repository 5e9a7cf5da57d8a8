// particlesolver_b200/csrc/psolver_cli.cpp — headless runner for both of the reference's applications.
// The reference's only front ends are two Qt viewers whose scenes are bound to keys (gpu/src/particleapp.cpp:141-215,
// cpu/src/view.cpp:129-177); this runs the same scenes without a display, over libpsolver.so's public C / C++ API only:
//   psolver_cli --app gpu --scene 7 --steps 600 [--dt 0.016667] [--grid 64] [--max-particles 15000] [--side 100]
//               [--iterations 5] [--xsph 0.01 --vorticity 0.3] [--self-collision] [--gas] [--staged-lambda]
//               [--emit]      the GPU app's fluid emitter switched on (key handling of particleapp.cpp:74-79: every 0.1 s of
//                             simulated time addFluid((-1,0,-1),(1,1,1), mass 1, density 1) until the system is full)
//               [--shoot K]   a left mouse click every K steps (particleapp.cpp:91-96): setParticleToAdd(eye, dir * 30, mass 2) —
//                             the viewer takes eye and ray from its camera; headless: eye (0,10,30) looking at (0,5,0)
//   psolver_cli --app gpu --scene c5 --planes 32 --ranks 2 --rank R --id-file F --steps 20     the synthetic PBF dam break
//               (BASELINE config C5), slab-decomposed in x over `ranks` processes, one GPU each (rank R uses device R unless
//               --device is given): ps_comm_init / ps_comm_set_slab / ps_comm_step (NCCL behind the C ABI).  Rank 0 writes the
//               128-byte NCCL id to F, the others wait for it.  --planes = lattice planes in x (100,000 particles each at the
//               default --ny 250 --nz 400; --vx V gives the block an initial x velocity); --ranks 1 runs the same scene undecomposed (ps_step).  [--dump-final FILE] writes the
//               owned particles' positions + velocities (8 floats each)
//   psolver_cli --app cpu --scene 6 --steps 1000 [--dt 0.01] [--stabilization 2]
//   psolver_cli --app session --script "6:100,1:50,w:20"     the CPU app as a user drives it: psb200::Simulation (constructor
//               builds WRECKING_BALL), then key presses and ticks; the rand() stream runs on across scenes like the reference's
// common: [--load FILE] [--save FILE] [--dump-every K --out DIR] [--json] [--device D]
// --load continues a checkpoint written by --save (ps_save / ps2d_save) instead of building a scene; --dump-every
// writes raw little-endian positions (header: "PSDUMP1\0", uint32 dims per particle, uint32 bytes per scalar, uint64 n).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <vector>
#include "../../include/particle_system.h"
#include "../../include/ps_scenes.h"
#include "../../include/ps_scenes2d.h"
#include "../../include/simulation2d.h"
#include "../../include/psolver.h"

namespace {
struct Args {
    std::string app = "gpu", scene = "7", load, save, out = "psolver_out", script;
    int steps = 100, grid = 64, side = 100, iterations = 5, dump_every = 0, device = 0;
    unsigned max_particles = 15000;
    double dt = -1;
    float xsph = 0.f, vorticity = 0.f;
    bool json = false;
    unsigned flags = 0;  // PS_FLAG_* switched on from the command line
    int ranks = 1, rank = 0, planes = 16, ny = 250, nz = 400;   // --scene c5
    bool device_given = false;
    std::string id_file, dump_final;
    float vx = 0.f;          // --scene c5: initial x velocity of the block (makes particles change owner in a decomposed run)
    bool emit = false;       // --app gpu: fluid emitter on
    int shoot = 0;           // --app gpu: shoot a particle every `shoot` steps (0 = never)
    int stabilization = -1;  // --app cpu: stabilization passes per tick (the reference's USE_STABILIZATION build: 2); -1 = leave as is
};
[[noreturn]] void die(const std::string &m) { fprintf(stderr, "psolver_cli: %s\n", m.c_str()); exit(1); }
void check(int r, const char *what) { if (r != PS_OK) die(std::string(what) + ": " + ps_last_error()); }

void dump(const std::string &dir, int step, const void *data, uint32_t dims, uint32_t scalar_bytes, uint64_t n) {
    mkdir(dir.c_str(), 0755);
    char name[512];
    snprintf(name, sizeof name, "%s/step%06d.bin", dir.c_str(), step);
    FILE *f = fopen(name, "wb");
    if (!f) die(std::string("cannot write ") + name);
    const char magic[8] = {'P', 'S', 'D', 'U', 'M', 'P', '1', 0};
    fwrite(magic, 1, 8, f); fwrite(&dims, 4, 1, f); fwrite(&scalar_bytes, 4, 1, f); fwrite(&n, 8, 1, f);
    fwrite(data, (size_t)dims * scalar_bytes, n, f);
    fclose(f);
}

int run_gpu(const Args &a) {
    psb200::ParticleSystem *ps = nullptr;
    PsCtx *ctx = nullptr;
    if (!a.load.empty()) {
        check(ps_load(a.load.c_str(), a.device, &ctx), "ps_load");
    } else {
        ps_scenes::SceneSpec s;
        s.scene = a.scene; s.grid = a.grid; s.max_particles = a.max_particles; s.iterations = a.iterations; s.side = a.side;
        ps = a.scene == "r" ? psb200::build_rigid_scene(a.grid, a.max_particles, a.iterations)
                            : ps_scenes::build<psb200::ParticleSystem>(s, psb200::colors, psb200::numColors);
        if (!ps) die("unknown GPU scene '" + a.scene + "' (1-9, c2, c3, r)");
        if (!ps->lastError().empty()) die(ps->lastError());
        ctx = ps->context();
    }
    if (a.flags) {
        PsParams p;
        check(ps_get_params(ctx, &p), "ps_get_params");
        p.flags |= a.flags;
        check(ps_set_params(ctx, &p), "ps_set_params");
    }
    if (a.xsph != 0.f || a.vorticity != 0.f) check(ps_set_viscosity(ctx, a.xsph, a.vorticity), "ps_set_viscosity");
    const float dt = a.dt > 0 ? (float)a.dt : 1.f / 60.f;
    if ((a.emit || a.shoot) && !ps) die("--emit / --shoot need a scene built through the host class (not --load)");
    const uint64_t n0 = ps_num_particles(ctx);
    uint64_t n = n0;
    std::vector<float> pos(4 * n), vel(4 * n), w(n);
    double dev_ms = 0, psteps = 0;
    float emit_timer = 0.f;   // ParticleApp::m_timer
    bool full = false;        // an emitter batch was dropped because maxParticles was reached
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 1; s <= a.steps; s++) {
        // ParticleApp::tick (particleapp.cpp:72-83): the emitter queues a block of fluid, then update() steps and appends what is queued
        if (a.emit && emit_timer <= 0.f) {
            ps->addFluid(make_int3(-1, 0, -1), make_int3(1, 1, 1), 1.f, 1.f, make_float3(0, 0, 1));
            emit_timer = 0.1f;
        }
        emit_timer -= dt;
        if (a.shoot > 0 && s % a.shoot == 0) {  // ParticleApp::mousePressed, left button
            const float3 eye = make_float3(0.f, 10.f, 30.f);
            const float len = std::sqrt(5.f * 5.f + 30.f * 30.f);
            ps->setParticleToAdd(eye, make_float3(0.f, -5.f / len * 30.f, -30.f / len * 30.f), 2.f);
        }
        if (ps) ps->update(dt); else check(ps_step(ctx, dt), "ps_step");
        // a full system drops the emitter's batch and carries on, like the reference (particlesystem.cpp:335); anything else is fatal
        if (ps && !ps->lastError().empty()) {
            if (ps->lastError().find("batch dropped") == std::string::npos) die(ps->lastError());
            full = true;
        }
        float ms = 0;
        check(ps_last_step_ms(ctx, &ms), "ps_last_step_ms");
        dev_ms += ms;
        psteps += (double)n;   // the step ran over the particles present before this step's appends
        n = ps_num_particles(ctx);
        if (a.dump_every > 0 && s % a.dump_every == 0) {
            pos.resize(4 * n);
            check(ps_download(ctx, PS_ARR_POS, pos.data(), 0, 4 * n), "ps_download");
            dump(a.out, s, pos.data(), 4, 4, n);
        }
    }
    check(ps_sync(ctx), "ps_sync");
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    pos.resize(4 * n); vel.resize(4 * n); w.resize(n);
    check(ps_download(ctx, PS_ARR_POS, pos.data(), 0, 4 * n), "ps_download");
    check(ps_download(ctx, PS_ARR_VEL, vel.data(), 0, 4 * n), "ps_download");
    check(ps_download(ctx, PS_ARR_INV_MASS, w.data(), 0, n), "ps_download");
    double ke = 0, sum = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (w[i] > 0) ke += .5 * (vel[4 * i] * (double)vel[4 * i] + vel[4 * i + 1] * (double)vel[4 * i + 1] + vel[4 * i + 2] * (double)vel[4 * i + 2]) / w[i];
        sum += pos[4 * i] + 2.0 * pos[4 * i + 1] + 3.0 * pos[4 * i + 2];
    }
    if (!a.save.empty()) check(ps_save(ctx, a.save.c_str()), "ps_save");
    double derr_mean = 0, derr_max = 0;
    check(ps_fluid_stats(ctx, &derr_mean, &derr_max, nullptr), "ps_fluid_stats");
    if (a.json)
        printf("{\"app\": \"gpu\", \"scene\": \"%s\", \"particles\": %llu, \"particles_at_start\": %llu, \"steps\": %d, \"dt\": %.9g, \"device_ms_per_step\": %.4f, \"wall_ms_per_step\": %.4f, "
               "\"particle_steps_per_s\": %.1f, \"kinetic_energy\": %.9g, \"position_checksum\": %.9g, \"launches_per_step\": %u, "
               "\"density_error_mean\": %.6g, \"density_error_max\": %.6g, \"emitter_hit_capacity\": %s}\n",
               a.scene.c_str(), (unsigned long long)n, (unsigned long long)n0, a.steps, dt, a.steps ? dev_ms / a.steps : 0., a.steps ? 1e3 * wall / a.steps : 0.,
               dev_ms > 0 ? psteps / (dev_ms * 1e-3) : 0., ke, sum, ps_launches_per_step(ctx), derr_mean, derr_max, full ? "true" : "false");
    else
        printf("gpu scene %s: %llu particles, %d steps, %.3f ms/step on the device (%.3f wall), KE %.6g\n", a.scene.c_str(), (unsigned long long)n, a.steps,
               a.steps ? dev_ms / a.steps : 0., a.steps ? 1e3 * wall / a.steps : 0., ke);
    if (ps) delete ps; else ps_destroy(ctx);
    return 0;
}

// ---- config C5: synthetic dam break, slab-decomposed over the ranks (particlesolver_b200/slab.py: dam_break_block, SlabDomain) ----
// counter-based uniforms in [0,1): splitmix64 finaliser of (global lattice index, stream, seed) — any rank count builds the same particles
double hash_uniform(uint64_t idx, uint64_t stream, uint64_t seed = 1234) {
    uint64_t z = (idx * 3ull + stream) * 0x9E3779B97F4A7C15ull + seed;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 40) / (double)(1 << 24);
}

int run_c5(const Args &a) {
    const double spacing = 0.625, jitter = 0.0025, origin = 0.3125;
    const float rho0 = 4.1f, drift = 0.25f;
    const int nx = a.planes, ny = a.ny, nz = a.nz, W = a.ranks, R = a.rank;
    if (W < 1 || R < 0 || R >= W || nx < W) die("--ranks >= 1, 0 <= --rank < ranks, --planes >= ranks");
    if (W > 1 && a.id_file.empty()) die("--ranks > 1 needs --id-file (rank 0 writes the NCCL id there)");
    const int device = a.device_given ? a.device : R;
    const int ix0 = (int)((long long)R * nx / W), ix1 = (int)((long long)(R + 1) * nx / W);
    const uint64_t plane = (uint64_t)ny * nz, n_mine = (uint64_t)(ix1 - ix0) * plane;
    const float halo_w = 2.f + drift;
    auto pow2_at_least = [](int v) { int p = 1; while (p < v) p <<= 1; return p; };
    const int halo_cells = (int)std::ceil(halo_w / 0.5) + 1, slab_cells = (int)std::ceil((ix1 - ix0) * spacing / 0.5);
    PsParams p;
    ps_default_params(&p);
    p.grid_size[0] = (uint32_t)pow2_at_least(slab_cells + 2 * halo_cells + 8);  // no aliasing inside a slab + its halo
    p.grid_size[1] = (uint32_t)pow2_at_least((int)std::ceil(ny * spacing / 0.5) * 2);
    p.grid_size[2] = (uint32_t)pow2_at_least((int)std::ceil(nz * spacing / 0.5) + 8);
    p.min_bounds[0] = 0; p.min_bounds[1] = 0; p.min_bounds[2] = 0;
    p.max_bounds[0] = (int)(2 * nx * spacing); p.max_bounds[1] = (int)(ny * spacing * 1.6); p.max_bounds[2] = (int)(nz * spacing);
    p.solver_iterations = (uint32_t)a.iterations;
    p.flags |= a.flags;
    const uint64_t halo_cap = (uint64_t)(plane * (halo_w + 2.0) / spacing), migr_cap = std::max<uint64_t>(plane * 2, 1u << 16);
    PsCtx *ctx = nullptr;
    check(ps_create(device, &p, n_mine + (W > 1 ? 2 * halo_cap + n_mine / 8 : 1024), &ctx), "ps_create");
    {   // this rank's lattice columns ix0 <= ix < ix1 (x fastest, then y, then z), a plane-pair at a time
        std::vector<float> pos, vel, w, ro;
        std::vector<int> ph;
        for (int zc = 0; zc < nz; zc++) {
            pos.clear();
            for (int y = 0; y < ny; y++)
                for (int x = ix0; x < ix1; x++) {
                    const uint64_t gid = ((uint64_t)zc * ny + y) * nx + x;
                    const int lat[3] = {x, y, zc};
                    for (int cdim = 0; cdim < 3; cdim++)
                        pos.push_back((float)(origin + lat[cdim] * spacing + (hash_uniform(gid, (uint64_t)cdim) * 2.0 - 1.0) * jitter));
                    pos.push_back(1.f);
                }
            const size_t k = pos.size() / 4;
            vel.assign(4 * k, 0.f); w.assign(k, 1.f); ro.assign(k, rho0); ph.assign(k, 0);
            for (size_t q = 0; q < k; q++) vel[4 * q] = a.vx;
            check(ps_append_particles(ctx, pos.data(), vel.data(), w.data(), ro.data(), ph.data(), k), "ps_append_particles");
        }
    }
    if (W > 1) {
        unsigned char id[PS_COMM_ID_BYTES];
        if (R == 0) {
            check(ps_comm_get_unique_id(id), "ps_comm_get_unique_id");
            const std::string tmp = a.id_file + ".tmp";
            FILE *f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) die("cannot write " + tmp);
            fclose(f);
            if (rename(tmp.c_str(), a.id_file.c_str()) != 0) die("cannot publish " + a.id_file);
        } else {
            FILE *f = nullptr;
            for (int tries = 0; tries < 6000 && !(f = fopen(a.id_file.c_str(), "rb")); tries++) usleep(10000);
            if (!f || fread(id, 1, sizeof id, f) != sizeof id) die("no NCCL id in " + a.id_file);
            fclose(f);
        }
        check(ps_comm_init(ctx, id, R, W), "ps_comm_init");
        const double cut_lo = nx * spacing * R / W, cut_hi = nx * spacing * (R + 1) / W;   // slab.uniform_cuts(0, nx * spacing, W)
        check(ps_comm_set_slab(ctx, R == 0 ? -INFINITY : (float)cut_lo, R == W - 1 ? INFINITY : (float)cut_hi, drift, 1, halo_cap, migr_cap), "ps_comm_set_slab");
    }
    const float dt = a.dt > 0 ? (float)a.dt : 1.f / 60.f;
    const int warm = std::min(3, a.steps / 2);
    double ms_total = 0;
    for (int s = 0; s < a.steps; s++) {
        if (s == warm) check(ps_timer_start(ctx), "ps_timer_start");
        check(W > 1 ? ps_comm_step(ctx, dt) : ps_step(ctx, dt), "step");
    }
    if (a.steps > warm) { float ms = 0; check(ps_timer_stop(ctx, &ms), "ps_timer_stop"); ms_total = ms; }
    check(ps_sync(ctx), "ps_sync");
    const uint64_t owned = ps_num_owned(ctx);
    std::vector<float> pos(4 * owned), vel(4 * owned);
    check(ps_download(ctx, PS_ARR_POS, pos.data(), 0, 4 * owned), "ps_download");
    check(ps_download(ctx, PS_ARR_VEL, vel.data(), 0, 4 * owned), "ps_download");
    double ke = 0, sum = 0;
    for (uint64_t i = 0; i < owned; i++) {
        ke += .5 * (vel[4 * i] * (double)vel[4 * i] + vel[4 * i + 1] * (double)vel[4 * i + 1] + vel[4 * i + 2] * (double)vel[4 * i + 2]);
        sum += pos[4 * i] + 2.0 * pos[4 * i + 1] + 3.0 * pos[4 * i + 2];
    }
    if (!a.dump_final.empty()) {
        FILE *f = fopen(a.dump_final.c_str(), "wb");
        if (!f) die("cannot write " + a.dump_final);
        for (uint64_t i = 0; i < owned; i++) { fwrite(&pos[4 * i], 4, 4, f); fwrite(&vel[4 * i], 4, 4, f); }
        fclose(f);
    }
    double derr_mean = 0, derr_max = 0;
    check(ps_fluid_stats(ctx, &derr_mean, &derr_max, nullptr), "ps_fluid_stats");
    double glob[4] = {(double)owned, ke, sum, derr_mean * (double)owned};
    uint64_t st[4] = {0, 0, 0, 0};
    if (W > 1) { check(ps_comm_allreduce_sum(ctx, glob, 4), "ps_comm_allreduce_sum"); check(ps_comm_stats(ctx, st), "ps_comm_stats"); }
    const int timed = a.steps - warm;
    printf("{\"app\": \"gpu\", \"scene\": \"c5\", \"rank\": %d, \"ranks\": %d, \"device\": %d, \"particles_owned\": %llu, \"particles_total\": %.0f, "
           "\"particles_expected\": %llu, \"steps\": %d, \"device_ms_per_step\": %.4f, \"particle_steps_per_s_total\": %.1f, \"kinetic_energy_total\": %.9g, "
           "\"position_checksum_total\": %.9g, \"density_error_mean_total\": %.6g, \"migrated_out\": %llu, \"ghosts\": %llu, \"bytes_sent\": %llu}\n",
           R, W, device, (unsigned long long)owned, glob[0], (unsigned long long)((uint64_t)nx * plane), a.steps, timed > 0 ? ms_total / timed : 0.,
           timed > 0 && ms_total > 0 ? glob[0] * timed / (ms_total * 1e-3) : 0., glob[1], glob[2], glob[0] > 0 ? glob[3] / glob[0] : 0., (unsigned long long)st[0],
           (unsigned long long)st[1], (unsigned long long)st[2]);
    ps_destroy(ctx);
    return 0;
}

int run_cpu_app(const Args &a) {
    Ps2dCtx *ctx = nullptr;
    if (!a.load.empty()) check(ps2d_load(a.load.c_str(), a.device, &ctx), "ps2d_load");
    else check(ps2d_build_scene(a.scene.c_str(), a.device, 0, &ctx), "ps2d_build_scene");
    if (a.stabilization >= 0) check(ps2d_set_stabilization_iterations(ctx, (uint32_t)a.stabilization), "ps2d_set_stabilization_iterations");
    const double dt = a.dt > 0 ? a.dt : .01;  // cpu/src/view.cpp:197
    std::vector<double> p;
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 1; s <= a.steps; s++) {
        check(ps2d_tick(ctx, dt), "ps2d_tick");
        if (a.dump_every > 0 && s % a.dump_every == 0) {
            const uint64_t n = ps2d_num_particles(ctx);
            p.resize(2 * n);
            check(ps2d_download(ctx, PS2D_ARR_P, p.data()), "ps2d_download");
            dump(a.out, s, p.data(), 2, 8, n);
        }
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const uint64_t n = ps2d_num_particles(ctx);
    double ke = 0;
    check(ps2d_kinetic_energy(ctx, &ke), "ps2d_kinetic_energy");
    if (!a.save.empty()) check(ps2d_save(ctx, a.save.c_str()), "ps2d_save");
    const char *name = ps2d_scene_name(a.scene.c_str());
    if (a.json)
        printf("{\"app\": \"cpu\", \"scene\": \"%s\", \"scene_name\": \"%s\", \"particles\": %llu, \"ticks\": %d, \"dt\": %.9g, \"wall_ms_per_tick\": %.4f, "
               "\"particle_steps_per_s\": %.1f, \"kinetic_energy\": %.17g, \"rand_calls\": %llu, \"launches_per_tick\": %u}\n",
               a.scene.c_str(), name ? name : "", (unsigned long long)n, a.steps, dt, a.steps ? 1e3 * wall / a.steps : 0., wall > 0 ? n * (double)a.steps / wall : 0., ke,
               (unsigned long long)ps2d_rand_calls(ctx), ps2d_launches_per_tick(ctx));
    else
        printf("cpu-app scene %s (%s): %llu particles, %d ticks, %.3f ms/tick, KE %.17g\n", a.scene.c_str(), name ? name : "checkpoint", (unsigned long long)n, a.steps,
               a.steps ? 1e3 * wall / a.steps : 0., ke);
    ps2d_destroy(ctx);
    return 0;
}

// a scripted session of the CPU app through the host class (include/simulation2d.h): "KEY:TICKS,KEY:TICKS,..."
int run_session(const Args &a, const std::string &script) {
    using namespace psb200;
    const SimulationType types[] = {FRICTION_TEST, SDF_TEST, GRANULAR_TEST, STACKS_TEST, WALL_TEST, PENDULUM_TEST, ROPE_TEST, FLUID_TEST, FLUID_SOLID_TEST,
                                    GAS_ROPE_TEST, WATER_BALLOON_TEST, CRADLE_TEST, SMOKE_OPEN_TEST, SMOKE_CLOSED_TEST, VOLCANO_TEST, WRECKING_BALL};
    try {
        Simulation sim(a.device);
        printf("[");
        size_t at = 0;
        bool first = true;
        while (at < script.size()) {
            size_t comma = script.find(',', at);
            if (comma == std::string::npos) comma = script.size();
            const std::string item = script.substr(at, comma - at);
            at = comma + 1;
            const size_t colon = item.find(':');
            const std::string key = item.substr(0, colon);
            const int ticks = colon == std::string::npos ? 0 : atoi(item.c_str() + colon + 1);
            bool found = false;
            for (SimulationType t : types)
                if (key == Simulation::key_of(t)) { sim.init(t); found = true; break; }
            if (!found) die("unknown scene key '" + key + "' in --script");
            for (int k = 0; k < ticks; k++) sim.tick(a.dt > 0 ? a.dt : .01);
            printf("%s{\"scene\": \"%s\", \"particles\": %d, \"ticks\": %d, \"kinetic_energy\": %.17g, \"rand_calls\": %llu}", first ? "" : ", ", key.c_str(),
                   sim.getNumParticles(), ticks, sim.getKineticEnergy(), (unsigned long long)ps2d_rand_calls(sim.context()));
            first = false;
        }
        printf("]\n");
    } catch (const std::exception &e) {
        die(e.what());
    }
    return 0;
}
}  // namespace

int main(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) die("missing value after " + k); return argv[++i]; };
        if (k == "--app") a.app = val();
        else if (k == "--scene") a.scene = val();
        else if (k == "--steps" || k == "--ticks") a.steps = atoi(val());
        else if (k == "--dt") a.dt = atof(val());
        else if (k == "--grid") a.grid = atoi(val());
        else if (k == "--side") a.side = atoi(val());
        else if (k == "--iterations") a.iterations = atoi(val());
        else if (k == "--max-particles") a.max_particles = (unsigned)strtoul(val(), nullptr, 10);
        else if (k == "--script") a.script = val();
        else if (k == "--load") a.load = val();
        else if (k == "--save") a.save = val();
        else if (k == "--dump-every") a.dump_every = atoi(val());
        else if (k == "--out") a.out = val();
        else if (k == "--xsph") a.xsph = (float)atof(val());
        else if (k == "--vorticity") a.vorticity = (float)atof(val());
        else if (k == "--self-collision") a.flags |= PS_FLAG_SELF_COLLISION;
        else if (k == "--gas") a.flags |= PS_FLAG_GAS;
        else if (k == "--staged-lambda") a.flags |= PS_FLAG_STAGED_LAMBDA;
        else if (k == "--emit") a.emit = true;
        else if (k == "--shoot") a.shoot = atoi(val());
        else if (k == "--stabilization") a.stabilization = atoi(val());
        else if (k == "--device") { a.device = atoi(val()); a.device_given = true; }
        else if (k == "--ranks") a.ranks = atoi(val());
        else if (k == "--rank") a.rank = atoi(val());
        else if (k == "--planes") a.planes = atoi(val());
        else if (k == "--ny") a.ny = atoi(val());
        else if (k == "--nz") a.nz = atoi(val());
        else if (k == "--id-file") a.id_file = val();
        else if (k == "--vx") a.vx = (float)atof(val());
        else if (k == "--dump-final") a.dump_final = val();
        else if (k == "--json") a.json = true;
        else if (k == "--help" || k == "-h") { printf("see the header of particlesolver_b200/csrc/psolver_cli.cpp\n"); return 0; }
        else die("unknown option " + k);
    }
    if (a.steps < 0 || a.grid <= 0 || (a.grid & (a.grid - 1))) die("--steps >= 0, --grid a power of two");
    if (a.app == "gpu" && a.scene == "c5") return run_c5(a);
    if (a.app == "gpu") return run_gpu(a);
    if (a.app == "cpu") return run_cpu_app(a);
    if (a.app == "session") return run_session(a, a.script);
    die("--app gpu | cpu | session");
}
