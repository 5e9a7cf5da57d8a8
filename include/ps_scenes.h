/*
 * include/ps_scenes.h — the reference's demo-scene table (gpu/src/particleapp.cpp:141-215, keys 1-9) plus the
 * scaled benchmark scenes of SURVEY.md §8 (C2, C3), written once as a template over the particle-system type
 * so that the very same script builds a scene in this repo's psb200::ParticleSystem and in the reference's
 * own ParticleSystem (oracle/ref_gpu_driver.cu) — scene parity by construction, including the order in which
 * glibc rand() is consumed for jitter and colours.
 */
#ifndef PS_SCENES_H
#define PS_SCENES_H
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector_types.h>
#include <vector_functions.h>

namespace ps_scenes {

struct SceneSpec {
    std::string scene = "7";  /* "1".."9" = reference keys, "c2" / "c3" / "c5r" = scaled configs */
    int grid = 64;            /* grid is grid^3 cells (reference: 64, particleapp.cpp:25) */
    unsigned max_particles = 15000; /* reference MAX_PARTICLES, particleapp.cpp:23 */
    int iterations = 5;       /* particleapp.cpp:38 */
    int side = 100;           /* c3: side^3 fluid particles; c2: side x side cloth */
};

/* COLORS must be the `colors` table that lives beside PS (the app indexes it with rand() % numColors) */
template <class PS, class COLORS>
PS *build(const SceneSpec &a, const COLORS &colors, int numColors) {
    const float R = 0.25f; /* PARTICLE_RADIUS, particleapp.cpp:24 */
    const uint3 g = make_uint3(a.grid, a.grid, a.grid);
    const int3 lo = make_int3(-50, 0, -50), hi = make_int3(50, 200, 50);
    const std::string &s = a.scene;
    PS *ps = nullptr;
    if (s == "1") { /* single rope (makeInitScene, particleapp.cpp:66-69) */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addRope(make_float3(0, 20, 0), make_float3(0, -.5, 0), .4f, 32, 1.f, true);
    } else if (s == "2") { /* single cloth */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addHorizCloth(make_int2(0, -3), make_int2(6, 3), make_float3(.5f, 7.f, .5f), make_float2(.3f, .3f), 3.f, false);
    } else if (s == "c2") { /* scene 2 scaled to side x side particles, bounds widened */
        ps = new PS(R, g, a.max_particles, make_int3(-200, 0, -200), make_int3(200, 200, 200), a.iterations);
        const int half = a.side / 4;
        ps->addHorizCloth(make_int2(0, -half), make_int2(2 * half, half), make_float3(.5f, 7.f, .5f), make_float2(.3f, .3f), 3.f, false);
    } else if (s == "3") { /* two fluids, different densities */
        ps = new PS(R, g, a.max_particles, make_int3(-7, 0, -5), make_int3(7, 20, 5), a.iterations);
        ps->addFluid(make_int3(-7, 0, -5), make_int3(7, 5, 5), 1.f, 2.f, colors[rand() % numColors]);
        ps->addFluid(make_int3(-7, 5, -5), make_int3(7, 10, 5), 1.f, 3.f, colors[rand() % numColors]);
    } else if (s == "4") { /* one solid stack */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addParticleGrid(make_int3(-3, 0, -3), make_int3(3, 20, 3), 1.f, false);
    } else if (s == "5") { /* three solid stacks */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addParticleGrid(make_int3(-10, 0, -3), make_int3(-7, 10, 3), 1.f, false);
        ps->addParticleGrid(make_int3(-3, 0, -3), make_int3(3, 10, 3), 1.f, false);
        ps->addParticleGrid(make_int3(7, 0, -3), make_int3(10, 10, 3), 1.f, false);
    } else if (s == "6") { /* particles on cloth */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addHorizCloth(make_int2(-10, -10), make_int2(10, 10), make_float3(.3f, 5.5f, .3f), make_float2(.1f, .1f), 10.f, true);
        ps->addParticleGrid(make_int3(-3, 6, -3), make_int3(3, 15, 3), 1.f, false);
    } else if (s == "7") { /* fluid blob */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addFluid(make_int3(-7, 6, -7), make_int3(7, 13, 7), 1.f, 1.5f, colors[rand() % numColors]);
    } else if (s == "c3") { /* scene 7 scaled to side^3 fluid particles (side=100 -> (-31,6,-31)..(32,69,32)) */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        const int ext = (int)std::ceil(a.side * 0.625f); /* builder count = (int)ceil(ur-ll)/0.625 */
        const int l = -(ext / 2);
        ps->addFluid(make_int3(l, 6, l), make_int3(l + ext, 6 + ext, l + ext), 1.f, 1.5f, colors[rand() % numColors]);
    } else if (s == "c5r") { /* the per-GPU share of the multi-GPU dam break (lattice spacing 2.5 r, rho0 4.1, 250 x 400 columns) built
                                through addFluid: `side` lattice planes in x (80 -> 7,968,000 particles), grid 128 x 512 x 512 like a rank of bench.py */
        const int xe = (int)std::ceil(a.side * 0.625f);
        unsigned gx = 1; while ((int)gx < 2 * xe + 16) gx <<= 1;
        ps = new PS(R, make_uint3(gx, 512, 512), a.max_particles, make_int3(0, 0, 0), make_int3(2 * xe, 256, 250), a.iterations);
        ps->addFluid(make_int3(0, 0, 0), make_int3(xe, 156, 250), 1.f, 4.1f, colors[rand() % numColors]);
    } else if (s == "8") { /* combo scene */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        ps->addHorizCloth(make_int2(14, -4), make_int2(24, 6), make_float3(.3f, 2.5f, .3f), make_float2(.25f, .25f), 10.f, true);
        ps->addHorizCloth(make_int2(10, -10), make_int2(25, -5), make_float3(.3f, 15.5f, .3f), make_float2(.25f, .25f), 3.f, false);
        ps->addRope(make_float3(-17, 20, -17), make_float3(0, -.5, 0.001f), .4f, 30, 1.f, true);
        ps->addRope(make_float3(-16, 20, -17), make_float3(0, 0, .5f), .4f, 50, 1.f, true);
        ps->addRope(make_float3(-17, 20, -16), make_float3(0, -.5, 0.001f), .4f, 40, 1.f, true);
        ps->addParticleGrid(make_int3(17, 6, 0), make_int3(21, 11, 4), 1.f, false);
        ps->addParticleGrid(make_int3(-12, 0, -20), make_int3(0, 12, -17), 1.f, false);
        ps->addParticleGrid(make_int3(-18, 0, -15), make_int3(-16, 9, -12), 1.f, false);
        ps->addStaticSphere(make_int3(5, 5, -10), make_int3(10, 10, -5), .5f);
    } else if (s == "9") { /* ropes on an immovable sphere */
        ps = new PS(R, g, a.max_particles, lo, hi, a.iterations);
        const float3 h = make_float3(0, 10, 0);
        for (int i = 0; i < 50; i++) {
            float angle = M_PI * i * 0.02f;
            float3 vec = make_float3(cos(angle), sin(angle), 0.f);
            ps->addRope(make_float3(vec.x * 5.f + h.x, vec.y * 5.f + h.y, vec.z * 5.f + h.z), make_float3(vec.x * .5f, vec.y * .5f, vec.z * .5f),
                        .35f, 30, 1.f, true);
        }
        ps->addStaticSphere(make_int3(-4, 7, -4), make_int3(4, 16, 4), .5f);
    }
    return ps;
}

}  // namespace ps_scenes
#endif
