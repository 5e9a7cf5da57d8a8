// Headless driver around the reference's UNMODIFIED 2-D CPU solver (cpu/src/simulation.cpp + constraint/*.cpp,
// compiled from /root/reference by oracle/Makefile into oracle/_ref/ref_cpu).  TEST INFRASTRUCTURE ONLY.
//   ref_cpu --scene KEY --ticks N [--dump dir [--dump-every K] [--scene-at T]] [--json]
//   ref_cpu --script "6:100,1:50,w:20"      one Simulation object, scenes switched like key presses
// Builds a scene exactly as the Qt app does (Simulation() runs init(WRECKING_BALL) first, then the key handler
// calls init(type): cpu/src/simulation.cpp:11-16, cpu/src/view.cpp:121-179), runs tick(.01) N times
// (cpu/src/view.cpp:185-202) and reports kinetic energy / timing; --dump writes particle state after chosen ticks.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
// every system / third-party header first, so that the access hack below only touches the reference's classes
#include <algorithm>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <QList>
#include <QHash>
#include <QSet>
#include "includes.h"
#define private public
#define protected public
#include "simulation.h"
#include "totalfluidconstraint.h"
#include "gasconstraint.h"
#include "distanceconstraint.h"
#include "totalshapeconstraint.h"
#include "opensmokeemitter.h"
#include "fluidemitter.h"
#undef protected
#undef private

// The reference draws glibc rand() for scene jitter and, every solver iteration, for the wall jitter of fluid particles
// (cpu/src/constraint/boundaryconstraint.cpp:19).  rand() is random() under the hood (same state, same algorithm), so this
// definition — which takes precedence over libc's for every translation unit of the executable — keeps the stream bit
// for bit and counts the draws: a parity run must start at the same position of the stream.
static long g_rand_calls = 0;
extern "C" int rand(void) { g_rand_calls++; return (int)random(); }

static SimulationType scene_of(const std::string &k) {
    // key bindings of cpu/src/view.cpp:129-177
    if (k == "1") return GRANULAR_TEST;
    if (k == "2") return STACKS_TEST;
    if (k == "3") return WALL_TEST;
    if (k == "4") return PENDULUM_TEST;
    if (k == "5") return ROPE_TEST;
    if (k == "6") return FLUID_TEST;
    if (k == "7") return FLUID_SOLID_TEST;
    if (k == "8") return GAS_ROPE_TEST;
    if (k == "9") return FRICTION_TEST;
    if (k == "0") return WATER_BALLOON_TEST;
    if (k == "n") return CRADLE_TEST;
    if (k == "s") return SMOKE_OPEN_TEST;
    if (k == "d") return SMOKE_CLOSED_TEST;
    if (k == ".") return SDF_TEST;
    if (k == "v") return VOLCANO_TEST;
    if (k == "w") return WRECKING_BALL;
    fprintf(stderr, "unknown scene key '%s'\n", k.c_str());
    exit(2);
}

// Everything a tick depends on besides the per-particle state of dump_state: friction coefficients, rigid bodies
// (member range, r vectors, SDF, centre, angle, inverse mass, stiffness), the STANDARD constraint list in order
// (fluid / gas / distance) and the smoke emitters.  JSON, 17 significant digits.
static void dump_scene(Simulation &sim, const std::string &dir, int tick) {
    char nm[64];
    snprintf(nm, sizeof nm, tick ? "/scene_t%05d.json" : "/scene.json", tick);
    std::string name = dir + nm;
    FILE *f = fopen(name.c_str(), "w");
    if (!f) { perror("fopen"); exit(1); }
    int n = sim.m_particles.size();
    fprintf(f, "{\"n\": %d, \"rand_calls\": %ld,\n \"xbounds\": [%.17g, %.17g], \"ybounds\": [%.17g, %.17g], \"gravity\": [%.17g, %.17g],\n", n, g_rand_calls,
            sim.m_xBoundaries.x, sim.m_xBoundaries.y, sim.m_yBoundaries.x, sim.m_yBoundaries.y, sim.m_gravity.x, sim.m_gravity.y);
    fprintf(f, " \"particles\": [");  // [px, py, vx, vy, imass, phase, bod, sFriction, kFriction, fx, fy, t]
    for (int i = 0; i < n; i++) {
        Particle *p = sim.m_particles[i];
        fprintf(f, "%s[%.17g, %.17g, %.17g, %.17g, %.17g, %d, %d, %.17g, %.17g, %.17g, %.17g, %.17g]", i ? ",\n  " : "", p->p.x, p->p.y, p->v.x, p->v.y, p->imass,
                (int)p->ph, p->bod, p->sFriction, p->kFriction, p->f.x, p->f.y, p->t);
    }
    fprintf(f, "],\n \"bodies\": [");
    for (int b = 0; b < sim.m_bodies.size(); b++) {
        Body *B = sim.m_bodies[b];
        fprintf(f, "%s{\"imass\": %.17g, \"center\": [%.17g, %.17g], \"angle\": %.17g, \"stiffness\": %.17g, \"particles\": [", b ? ",\n  " : "", B->imass,
                B->center.x, B->center.y, B->angle, B->shape->stiffness);
        for (int k = 0; k < B->particles.size(); k++) fprintf(f, "%s%d", k ? ", " : "", B->particles[k]);
        fprintf(f, "], \"rs\": [");
        for (int k = 0; k < B->particles.size(); k++) { glm::dvec2 r = B->rs[B->particles[k]]; fprintf(f, "%s[%.17g, %.17g]", k ? ", " : "", r.x, r.y); }
        fprintf(f, "], \"sdf\": [");
        for (int k = 0; k < B->particles.size(); k++) { SDFData d = B->sdf[B->particles[k]]; fprintf(f, "%s[%.17g, %.17g, %.17g]", k ? ", " : "", d.gradient.x, d.gradient.y, d.distance); }
        fprintf(f, "]}");
    }
    fprintf(f, "],\n \"standard\": [");
    QList<Constraint *> &glob = sim.m_globalConstraints[STANDARD];
    for (int c = 0; c < glob.size(); c++) {
        fprintf(f, "%s", c ? ",\n  " : "");
        if (TotalFluidConstraint *t = dynamic_cast<TotalFluidConstraint *>(glob[c])) {
            fprintf(f, "{\"type\": \"fluid\", \"p0\": %.17g, \"ps\": [", t->p0);
            for (int k = 0; k < t->ps.size(); k++) fprintf(f, "%s%d", k ? ", " : "", t->ps[k]);
            fprintf(f, "]}");
        } else if (GasConstraint *g = dynamic_cast<GasConstraint *>(glob[c])) {
            fprintf(f, "{\"type\": \"gas\", \"p0\": %.17g, \"open\": %d, \"ps\": [", g->p0, g->m_open ? 1 : 0);
            for (int k = 0; k < g->ps.size(); k++) fprintf(f, "%s%d", k ? ", " : "", g->ps[k]);
            fprintf(f, "]}");
        } else if (DistanceConstraint *d = dynamic_cast<DistanceConstraint *>(glob[c])) {
            fprintf(f, "{\"type\": \"distance\", \"i1\": %d, \"i2\": %d, \"d\": %.17g}", d->i1, d->i2, d->d);
        } else {
            fprintf(f, "{\"type\": \"unknown\"}");
        }
    }
    fprintf(f, "],\n \"smoke_emitters\": [");
    for (int e = 0; e < sim.m_smokeEmitters.size(); e++) {
        OpenSmokeEmitter *E = sim.m_smokeEmitters[e];
        int gi = -1;
        for (int c = 0; c < glob.size(); c++) if ((Constraint *)E->m_gs == glob[c]) gi = c;
        fprintf(f, "%s{\"posn\": [%.17g, %.17g], \"rate\": %.17g, \"timer\": %.17g, \"standard_index\": %d}", e ? ", " : "", E->m_posn.x, E->m_posn.y,
                E->m_particlesPerSec, E->timer, gi);
    }
    fprintf(f, "], \"fluid_emitters\": [");
    for (int e = 0; e < sim.m_fluidEmitters.size(); e++) {
        FluidEmitter *E = sim.m_fluidEmitters[e];
        int gi = -1;
        for (int c = 0; c < glob.size(); c++) if ((Constraint *)E->m_fs == glob[c]) gi = c;
        fprintf(f, "%s{\"posn\": [%.17g, %.17g], \"rate\": %.17g, \"timer\": %.17g, \"total_timer\": %.17g, \"standard_index\": %d}", e ? ", " : "", E->m_posn.x,
                E->m_posn.y, E->m_particlesPerSec, E->timer, E->totalTimer, gi);
    }
    fprintf(f, "]}\n");
    fclose(f);
}

static void dump_state(Simulation &sim, const std::string &dir, int tick) {
    char name[256];
    snprintf(name, sizeof name, "%s/tick%05d.bin", dir.c_str(), tick);
    FILE *f = fopen(name, "wb");
    if (!f) { perror("fopen"); exit(1); }
    int n = sim.m_particles.size();
    fwrite(&n, sizeof n, 1, f);
    // which TotalFluidConstraint (in m_globalConstraints[STANDARD] order) a particle belongs to, -1 if none
    std::vector<int> fluid(n, -1);
    QList<Constraint *> &glob = sim.m_globalConstraints[STANDARD];
    int nf = 0;
    for (int c = 0; c < glob.size(); c++)
        if (TotalFluidConstraint *t = dynamic_cast<TotalFluidConstraint *>(glob[c])) {
            for (int k = 0; k < t->ps.size(); k++) fluid[t->ps[k]] = nf;
            nf++;
        }
    for (int i = 0; i < n; i++) {
        Particle *p = sim.m_particles[i];
        double rec[8] = {p->p.x, p->p.y, p->v.x, p->v.y, p->imass, (double)p->ph, (double)p->bod, (double)fluid[i]};
        fwrite(rec, sizeof rec, 1, f);
    }
    fclose(f);
    snprintf(name, sizeof name, "%s/tick%05d.txt", dir.c_str(), tick);
    f = fopen(name, "w");
    fprintf(f, "n %d\nrand_calls %ld\nke %.17g\nxbounds %.17g %.17g\nybounds %.17g %.17g\ngravity %.17g %.17g\nfluids %d", n, g_rand_calls,
            sim.getKineticEnergy(), sim.m_xBoundaries.x, sim.m_xBoundaries.y, sim.m_yBoundaries.x, sim.m_yBoundaries.y, sim.m_gravity.x,
            sim.m_gravity.y, nf);
    for (int c = 0; c < glob.size(); c++)
        if (TotalFluidConstraint *t = dynamic_cast<TotalFluidConstraint *>(glob[c])) fprintf(f, " %.17g", t->p0);
    fprintf(f, "\n");
    fclose(f);
}

int main(int argc, char **argv) {
    std::string scene = "6", dump, script;
    int ticks = 100, dump_every = 0, scene_at = -1;
    double fluid_scale = 0;  // --fluid-scale S: scene 6 rebuilt at that scale (initFluid hard-codes 4) through the reference's own createFluid
    bool json = false;
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i];
        if (k == "--scene" && i + 1 < argc) scene = argv[++i];
        else if (k == "--ticks" && i + 1 < argc) ticks = atoi(argv[++i]);
        else if (k == "--dump" && i + 1 < argc) dump = argv[++i];
        else if (k == "--dump-every" && i + 1 < argc) dump_every = atoi(argv[++i]);
        else if (k == "--scene-at" && i + 1 < argc) scene_at = atoi(argv[++i]);  // full restart state after that tick
        else if (k == "--fluid-scale" && i + 1 < argc) fluid_scale = atof(argv[++i]);
        else if (k == "--script" && i + 1 < argc) script = argv[++i];  // a session: "KEY:TICKS,KEY:TICKS,..." on ONE Simulation object
        else if (k == "--json") json = true;
    }
    Simulation sim;  // constructor builds WRECKING_BALL first, consuming rand() like the app does
    if (!script.empty()) {  // key presses and ticks like a user's session: the rand() stream runs on across scenes
        printf("[");
        size_t at = 0;
        bool first = true;
        while (at < script.size()) {
            size_t comma = script.find(',', at);
            if (comma == std::string::npos) comma = script.size();
            const std::string item = script.substr(at, comma - at);
            at = comma + 1;
            const size_t colon = item.find(':');
            const int ticks = colon == std::string::npos ? 0 : atoi(item.c_str() + colon + 1);
            sim.init(scene_of(item.substr(0, colon)));
            for (int k = 0; k < ticks; k++) sim.tick(.01);
            printf("%s{\"scene\": \"%s\", \"particles\": %d, \"ticks\": %d, \"kinetic_energy\": %.17g, \"rand_calls\": %ld}", first ? "" : ", ",
                   item.substr(0, colon).c_str(), sim.getNumParticles(), ticks, sim.getKineticEnergy(), g_rand_calls);
            first = false;
        }
        printf("]\n");
        return 0;
    }
    sim.init(scene_of(scene));
    if (fluid_scale > 0 && scene == "6") {
        // BASELINE.md section 4's scaled replicas: initFluid's recipe (simulation.cpp:894-913: spacing 0.7, jitter +-0.1, two fluids of
        // density 1 and 1.75) at another scale, built through the reference's own createFluid; init()'s epilogue repeated
        sim.clear();
        sim.m_counts = nullptr;
        const double scale = fluid_scale, delta = .7, num = 2.;
        sim.m_gravity = glm::dvec2(0, -9.8);
        sim.m_xBoundaries = glm::dvec2(-2 * scale, 2 * scale);
        sim.m_yBoundaries = glm::dvec2(-2 * scale, 10 * scale);
        QList<Particle *> particles;
        for (int d = 0; d < num; d++) {
            double start = -2 * scale + 4 * scale * (d / num);
            for (double x = start; x < start + (4 * scale / num); x += delta)
                for (double y = -2 * scale; y < scale; y += delta)
                    particles.append(new Particle(glm::dvec2(x, y) + .2 * glm::dvec2(frand() - .5, frand() - .5), 1));
            sim.createFluid(&particles, 1 + .75 * d);
            particles.clear();
        }
        sim.m_standardSolver.setupM(&sim.m_particles);
        sim.m_counts = new int[sim.m_particles.size()];
    }
    int n = sim.getNumParticles();
    if (!dump.empty()) { std::string c = "mkdir -p '" + dump + "'"; if (system(c.c_str())) return 1; dump_scene(sim, dump, 0); dump_state(sim, dump, 0); }
    std::vector<double> ke;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 1; t <= ticks; t++) {
        sim.tick(.01);  // cpu/src/view.cpp:197
        if (t == 1 || t == 100 || t == 1000 || t == ticks) ke.push_back(sim.getKineticEnergy());
        if (!dump.empty() && t == scene_at) dump_scene(sim, dump, t);
        if (!dump.empty() && (t == 1 || (dump_every > 0 && t % dump_every == 0) || t == ticks)) dump_state(sim, dump, t);
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (json) {
        printf("{\"impl\": \"reference_cpu_unmodified\", \"scene\": \"%s\", \"n\": %d, \"ticks\": %d, \"ms_per_tick\": %.4f, "
               "\"particle_steps_per_s\": %.1f, \"solver_iterations\": %d, \"cores\": 1, \"ke_last\": %.17g}\n",
               scene.c_str(), n, ticks, 1e3 * sec / ticks, n * (double)ticks / sec, SOLVER_ITERATIONS, ke.empty() ? 0.0 : ke.back());
    } else {
        printf("scene %s: %d particles, %d ticks in %.3f s\n", scene.c_str(), n, ticks, sec);
        for (double k : ke) printf("KE %.17g\n", k);
    }
    return 0;
}
