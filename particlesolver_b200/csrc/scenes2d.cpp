// particlesolver_b200/csrc/scenes2d.cpp — the reference CPU application's scene builders (include/ps_scenes2d.h).
// Host-side setup code (g++ -ffp-contract=off): every coordinate is computed with the reference's own expression, in
// double precision, and jitter is drawn from the context's glibc rand() stream in the reference's order, so that a
// scene is the reference's bit for bit (tests/test_scenes2d.py compares with the reference's dumped scenes).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/ps_scenes2d.h"
#include "../../include/psolver.h"

void ps_set_error(const char *fmt, ...);
extern "C" int ps2d_rand(Ps2dCtx *ctx, int *out);

namespace {
constexpr double RAD = .25, DIAM = .5, EPSILON = .0001;  // particle.h:6-7, includes.h:34
inline double D2R(double d) { return d * M_PI / 180; }    // includes.h:36

struct Part {  // Particle(pos, mass, phase) — particle.h:31-52
    double x, y, vx = 0, vy = 0, imass, sf = 0, kf = 0;
    int phase, bod = -1;
    Part(double px, double py, double mass, int ph = PS2D_PHASE_SOLID) : x(px), y(py), phase(ph) { imass = mass <= 0 ? -mass : 1. / mass; }
};
struct Sdf { double gx, gy, d; };
inline Sdf sdf(double x, double y, double dist) {  // SDFData(glm::normalize(dvec2(x, y)), dist)
    const double inv = 1. / std::sqrt(x * x + y * y);
    return Sdf{x * inv, y * inv, dist};
}

struct Builder {
    Ps2dCtx *c = nullptr;
    int err = PS_OK;
    uint64_t count = 0;

    float frand() {  // includes.h:25 — float-typed
        int r = 0;
        if (err == PS_OK) err = ps2d_rand(c, &r);
        return (float)((double)r / (double)2147483647);
    }
    // glm::dvec2(x, y) + .2 * glm::dvec2(frand() - .5, frand() - .5): g++ evaluates the arguments of the inner constructor
    // right to left, so the y jitter is drawn first (checked against the reference's dumped scenes)
    Part jittered(double x, double y, double mass, int phase) {
        const double jy = frand() - .5;
        const double jx = frand() - .5;
        return Part(x + .2 * jx, y + .2 * jy, mass, phase);
    }
    uint64_t add(const std::vector<Part> &ps) {
        const size_t n = ps.size();
        std::vector<double> p(2 * n), v(2 * n), im(n), sf(n), kf(n);
        std::vector<int32_t> ph(n), bd(n);
        for (size_t k = 0; k < n; k++) {
            p[2 * k] = ps[k].x; p[2 * k + 1] = ps[k].y; v[2 * k] = ps[k].vx; v[2 * k + 1] = ps[k].vy;
            im[k] = ps[k].imass; sf[k] = ps[k].sf; kf[k] = ps[k].kf; ph[k] = ps[k].phase; bd[k] = ps[k].bod;
        }
        uint64_t first = count;
        if (err == PS_OK && n) err = ps2d_add_particles(c, p.data(), v.data(), im.data(), ph.data(), bd.data(), sf.data(), kf.data(), n, &first);
        count += n;
        return first;
    }
    void body(const std::vector<Part> &ps, const std::vector<Sdf> &data) {  // createRigidBody(verts, sdfData)
        const size_t n = ps.size();
        std::vector<double> p(2 * n), v(2 * n), im(n), sf(n), kf(n), s3(3 * n);
        for (size_t k = 0; k < n; k++) {
            p[2 * k] = ps[k].x; p[2 * k + 1] = ps[k].y; v[2 * k] = ps[k].vx; v[2 * k + 1] = ps[k].vy;
            im[k] = ps[k].imass; sf[k] = ps[k].sf; kf[k] = ps[k].kf;
            s3[3 * k] = data[k].gx; s3[3 * k + 1] = data[k].gy; s3[3 * k + 2] = data[k].d;
        }
        if (err == PS_OK) err = ps2d_create_rigid_body(c, p.data(), v.data(), im.data(), sf.data(), kf.data(), s3.data(), n, nullptr);
        count += n;
    }
    void group(std::vector<Part> &ps, double density, bool gas, bool open, uint32_t *index = nullptr) {  // createFluid / createGas
        (void)frand();  // int bod = 100 * frand();  (the tag is never read for fluids: only SOLID pairs compare bodies)
        const size_t n = ps.size();
        std::vector<double> p(2 * n), v(2 * n), im(n);
        for (size_t k = 0; k < n; k++) { p[2 * k] = ps[k].x; p[2 * k + 1] = ps[k].y; v[2 * k] = ps[k].vx; v[2 * k + 1] = ps[k].vy; im[k] = ps[k].imass; }
        if (err == PS_OK) err = gas ? ps2d_create_gas(c, p.data(), v.data(), im.data(), n, density, open ? 1 : 0, index) : ps2d_create_fluid(c, p.data(), v.data(), im.data(), n, density);
        count += n;
        ps.clear();
    }
    void dist(uint32_t i1, uint32_t i2, double d = -1.) { if (err == PS_OK) err = ps2d_add_distance_constraint(c, i1, i2, d); }
};

struct SceneInfo { const char *key, *name; double xb[2], yb[2]; uint64_t room; };
const SceneInfo kScenes[] = {
    {"1", "GRANULAR_TEST", {-100, 100}, {-5, 1000}, 1024},     {"2", "STACKS_TEST", {-20, 20}, {0, 1000000}, 1024},
    {"3", "WALL_TEST", {-50, 50}, {0, 1000000}, 2048},         {"4", "PENDULUM_TEST", {-10, 10}, {0, 1000000}, 64},
    {"5", "ROPE_TEST", {-5, 5}, {0, 1000000}, 256},            {"6", "FLUID_TEST", {-8, 8}, {-8, 40}, 1024},
    {"7", "FLUID_SOLID_TEST", {-6, 6}, {-6, 300}, 1024},       {"8", "GAS_ROPE_TEST", {-8, 8}, {-4, 200}, 8192},
    {"9", "FRICTION_TEST", {-20, 20}, {0, 1000000}, 64},       {"0", "WATER_BALLOON_TEST", {-10, 10}, {-10, 1000000}, 1024},
    {"n", "CRADLE_TEST", {-10, 10}, {-5, 1000000}, 64},        {"s", "SMOKE_OPEN_TEST", {-6, 6}, {-4, 200}, 8192},
    {"d", "SMOKE_CLOSED_TEST", {-4, 4}, {-4, 4}, 1024},        {".", "SDF_TEST", {-20, 20}, {0, 1000000}, 64},
    {"w", "WRECKING_BALL", {-15, 100}, {0, 1000000}, 1024},  {"v", "VOLCANO_TEST", {-20, 20}, {0, 100}, 1024},
    // scaled replicas of FLUID_TEST (the scene built through createFluid with initFluid's spacing / jitter recipe at
    // twice and four times the linear size: 1,728 and 6,912 particles); not key-bound in the reference
    {"6x2", "FLUID_TEST x2", {-16, 16}, {-16, 80}, 4096},      {"6x4", "FLUID_TEST x4", {-32, 32}, {-32, 160}, 16384},
};
const SceneInfo *find_scene(const char *key) {
    if (!key) return nullptr;
    for (const SceneInfo &s : kScenes) if (!std::strcmp(s.key, key)) return &s;
    return nullptr;
}

std::vector<Sdf> box_sdf(int columns) {  // the SDF table of a (columns x 2) box: corners sqrt(2) r, edges r
    const double root2 = std::sqrt(2);
    std::vector<Sdf> d;
    d.push_back(sdf(-1, -1, RAD * root2)); d.push_back(sdf(-1, 1, RAD * root2));
    for (int i = 0; i < columns - 2; i++) { d.push_back(sdf(0, -1, RAD)); d.push_back(sdf(0, 1, RAD)); }
    d.push_back(sdf(1, -1, RAD * root2)); d.push_back(sdf(1, 1, RAD * root2));
    return d;
}

void brick_wall(Builder &B, int height, int width, double mass, double sf, double kf) {  // initWall / initWreckingBall, simulation.cpp:806-842,1220-1252
    const double dimx = 6, dimy = 2;
    const std::vector<Sdf> data = box_sdf(6);
    for (int j = -width; j <= width; j++)
        for (int i = height - 1; i >= 0; i--) {
            std::vector<Part> v;
            for (int x = 0; x < dimx; x++) {
                const double num = (i % 2 == 0 ? 3 : -1);
                const double xVal = j * (EPSILON + dimx / 2.) + DIAM * (x % (int)dimx) - num * RAD;
                for (int y = 0; y < dimy; y++) {
                    const double yVal = (i * dimy + (y % (int)dimy) + EPSILON) * DIAM + RAD;
                    Part p(xVal, yVal, mass);
                    p.sf = sf; p.kf = kf;
                    v.push_back(p);
                }
            }
            B.body(v, data);
        }
}

void rope_row(Builder &B, double x0, double x1, double top, double dist, double mass, bool end_particle) {  // initRope / initRopeGas
    Part e1(x0, top, 0);
    e1.bod = -2;
    B.add({e1});
    for (double i = x0 + dist; i < x1 - dist; i += dist) {
        Part p(i, top, mass);
        p.bod = -2;
        B.add({p});
        B.dist((uint32_t)B.count - 2, (uint32_t)B.count - 1, dist);
    }
    if (end_particle) {
        Part e2(x1, top, 0);
        e2.bod = -2;
        B.add({e2});
    }
    B.dist((uint32_t)B.count - 2, (uint32_t)B.count - 1, dist);
}

void build(Builder &B, const std::string &k, const SceneInfo &S) {
    const double root2 = std::sqrt(2);
    if (k == "9") {  // initFriction, simulation.cpp:658-686
        std::vector<Sdf> data = {sdf(-1, -1, RAD * root2), sdf(-1, 1, RAD * root2), sdf(0, -1, RAD), sdf(0, 1, RAD), sdf(1, -1, RAD * root2), sdf(1, 1, RAD * root2)};
        std::vector<Part> v;
        const int dx = 3, dy = 2;
        for (int x = 0; x < dx; x++) {
            const double xVal = DIAM * ((x % dx) - dx / 2);
            for (int y = 0; y < dy; y++) {
                const double yVal = (dy + (y % dy) + 1) * DIAM;
                Part p(xVal, yVal, 1.);
                p.vx = 5; p.kf = .01; p.sf = .1;
                v.push_back(p);
            }
        }
        B.body(v, data);
    } else if (k == "1") {  // initGranular, :688-707
        std::vector<Part> v;
        for (int i = -15; i <= 15; i++)
            for (int j = 0; j < 30; j++) {
                Part p(i * (DIAM + EPSILON), std::pow(j, 1.2) * DIAM + RAD + S.yb[0], 1);
                p.sf = .35; p.kf = .3;
                v.push_back(p);
            }
        Part jerk(-25.55, 40, 100.f);
        jerk.vx = 8.5;
        v.push_back(jerk);
        B.add(v);
    } else if (k == ".") {  // initSdf, :709-739
        std::vector<Sdf> data = {sdf(-1, -1, RAD * root2), sdf(-1, 0, RAD), sdf(-1, 1, RAD * root2), sdf(1, -1, RAD * root2), sdf(1, 0, RAD), sdf(1, 1, RAD * root2)};
        const int dx = 2, dy = 3;
        for (int i = 2 - 1; i >= 0; i--) {
            std::vector<Part> v;
            for (int x = 0; x < dx; x++) {
                const double xVal = DIAM * ((x % dx) - dx / 2) + i * RAD;
                for (int y = 0; y < dy; y++) {
                    const double yVal = ((40 * i) * dy + (y % dy) + 1) * DIAM;
                    Part p(xVal, yVal, 4.);
                    if (i > 0) p.vy = -120;
                    v.push_back(p);
                }
            }
            B.body(v, data);
        }
    } else if (k == "2") {  // initBoxes, :741-776
        std::vector<Sdf> data = {sdf(-1, -1, RAD * root2), sdf(-1, 1, RAD * root2), sdf(0, -1, RAD), sdf(0, 1, RAD), sdf(1, -1, RAD * root2), sdf(1, 1, RAD * root2)};
        const int numBoxes = 25, numColumns = 2, dx = 3, dy = 2;
        for (int j = -numColumns; j <= numColumns; j++)
            for (int i = numBoxes - 1; i >= 0; i--) {
                std::vector<Part> v;
                for (int x = 0; x < dx; x++) {
                    const double xVal = j * 4 + DIAM * ((x % dx) - dx / 2);
                    for (int y = 0; y < dy; y++) {
                        const double yVal = ((2 * i + 1) * dy + (y % dy) + 1) * DIAM;
                        Part p(xVal, yVal, 4.);
                        p.sf = 1.; p.kf = 1.;
                        v.push_back(p);
                    }
                }
                B.body(v, data);
            }
    } else if (k == "3") {  // initWall, :778-842
        brick_wall(B, 11, 5, 1., 1, 0);
    } else if (k == "4") {  // initPendulum, :844-881
        const int chainLength = 3;
        B.add({Part(0 * DIAM + 0, (chainLength * 3 + 6) * DIAM + 2, 0)});
        std::vector<Sdf> data = {sdf(-1, -1, RAD), sdf(-1, 1, RAD), sdf(0, -1, RAD), sdf(0, 1, RAD), sdf(1, -1, RAD), sdf(1, 1, RAD)};
        const double xs[6] = {-1, -1, 0, 0, 1, 1};
        for (int i = chainLength; i >= 0; i--) {
            std::vector<Part> v;
            for (int j = 0; j < 6; j++) {
                const double y = ((i + 1) * 3 + (j % 2)) * DIAM + 2;
                Part p(xs[j] * DIAM, y, 1.);
                p.vx = 3;
                v.push_back(p);
            }
            B.body(v, data);
            if (i < chainLength) {
                const int basePrev = 1 + (chainLength - i - 1) * 6, baseCur = basePrev + 6;
                B.dist(baseCur + 1, basePrev);
                B.dist(baseCur + 5, basePrev + 4);
            }
        }
        B.dist(0, 4);
    } else if (k == "5") {  // initRope, :883-921
        const double scale = 5., delta = .7;
        rope_row(B, S.xb[0], S.xb[1], 6, RAD, 1., true);
        std::vector<Part> f;
        for (double x = -scale; x < scale; x += delta)
            for (double y = 10; y < 10 + scale; y += delta) f.push_back(B.jittered(x, y, 1, PS2D_PHASE_FLUID));
        B.group(f, 1.75, false, false);
    } else if (k == "6" || k == "6x2" || k == "6x4") {  // initFluid, :923-943 (scale 4; the replicas: 8 and 16)
        const double scale = k == "6" ? 4. : k == "6x2" ? 8. : 16., delta = .7, num = 2.;
        std::vector<Part> f;
        for (int d = 0; d < num; d++) {
            const double start = -2 * scale + 4 * scale * (d / num);
            for (double x = start; x < start + (4 * scale / num); x += delta)
                for (double y = -2 * scale; y < scale; y += delta) f.push_back(B.jittered(x, y, 1, PS2D_PHASE_FLUID));
            B.group(f, 1 + .75 * d, false, false);
        }
    } else if (k == "7") {  // initFluidSolid, :945-1011
        const double scale = 3., delta = .7, num = 1.;
        std::vector<Part> f;
        for (int d = 0; d < num; d++) {
            const double start = -2 * scale + 4 * scale * (d / num);
            for (double x = start; x < start + (4 * scale / num); x += delta)
                for (double y = -2 * scale; y < 2 * scale; y += delta) f.push_back(B.jittered(x, y + 3, 1, PS2D_PHASE_FLUID));
            B.group(f, 1. + 1.25 * (d + 1), false, false);
        }
        const std::vector<Sdf> data = box_sdf(5);
        const int dx = 5, dy = 2;
        for (int b = 0; b < 2; b++) {
            std::vector<Part> v;
            for (int x = 0; x < dx; x++) {
                const double xVal = DIAM * ((x % dx) - dx / 2);
                for (int y = 0; y < dy; y++) {
                    const double yVal = (dy + (y % dy) + 1) * DIAM;
                    v.push_back(b == 0 ? Part(xVal - 3, yVal + 10, 2) : Part(xVal + 3, yVal + 10, .2));
                }
            }
            B.body(v, data);
        }
    } else if (k == "0") {  // initWaterBalloon, :1047-1103
        const double samples = 60, da = 360. / samples;
        for (int ring = 0; ring < 2; ring++) {
            const uint32_t first = (uint32_t)B.count;
            for (int i = 0; i < samples; i++) {
                const double angle = D2R(i * da);
                Part p(std::sin(angle) * 3., (ring == 0 ? std::cos(angle) : std::cos(angle) + 3) * 3., 1);
                p.bod = ring == 0 ? -2 : -3;
                const uint32_t idx = (uint32_t)B.count;
                B.add({p});
                if (i > 0) B.dist(idx, idx - 1);
            }
            B.dist(first, (uint32_t)B.count - 1);
        }
        const double delta = 1.5 * RAD;
        for (int blob = 0; blob < 2; blob++) {
            std::vector<Part> f;
            for (double x = -2; x <= 2; x += delta)
                for (double y = -2; y <= 2; y += delta) f.push_back(blob == 0 ? B.jittered(x, y, 1, PS2D_PHASE_FLUID) : B.jittered(x, y + 9, 1, PS2D_PHASE_FLUID));
            B.group(f, 1.75, false, false);
        }
    } else if (k == "n") {  // initNewtonsCradle, :1105-1122
        const int n = 2;
        for (int i = -n; i <= n; i++) {
            const uint32_t idx = (uint32_t)B.count;
            B.add({Part(i * DIAM, 0, 0.f)});
            if (i != -n) B.add({Part(i * DIAM, -3, 1.f)});
            else B.add({Part(i * DIAM - 3, 0, 1.f)});
            B.dist(idx, idx + 1);
        }
    } else if (k == "s" || k == "d") {  // initSmokeOpen / initSmokeClosed, :1124-1164
        const double scale = 2., delta = .63, start = -2 * scale;
        std::vector<Part> g;
        for (double x = start; x < start + (4 * scale); x += delta)
            for (double y = -2 * scale; y < 2 * scale; y += delta) g.push_back(B.jittered(x, y, 1, PS2D_PHASE_GAS));
        uint32_t gs = 0;
        B.group(g, 1.5, true, k == "s", &gs);
        const double posn[2] = {0, -2 * scale + 1};
        if (B.err == PS_OK) B.err = ps2d_create_smoke_emitter(B.c, posn, 15, k == "s" ? gs : UINT32_MAX, 0.);
    } else if (k == "8") {  // initRopeGas, :1166-1204
        const double scale = 2., delta = .63;
        rope_row(B, 0, 4 * scale, 12, RAD, 2, false);
        std::vector<Part> g;
        const double start = -.5 * scale;
        for (double x = start; x < start + (1 * scale); x += delta)
            for (double y = -.5 * scale; y < .5 * scale; y += delta) g.push_back(B.jittered(x, y, 1, PS2D_PHASE_GAS));
        uint32_t gs = 0;
        B.group(g, 1.5, true, true, &gs);
        const double posn[2] = {0, 0};
        if (B.err == PS_OK) B.err = ps2d_create_smoke_emitter(B.c, posn, 15, gs, 0.);
    } else if (k == "v") {  // initVolcano, :1206-1216 + :1177-1203
        const double scale = 10.;
        double delta = .2;
        std::vector<Part> slope;
        for (double x = 1.; x <= scale; x += delta) {
            slope.push_back(Part(-x, scale - x, 0));
            slope.push_back(Part(x, scale - x, 0));
        }
        B.add(slope);
        delta = .8;
        std::vector<Part> f;
        for (double y = 0.; y < scale - 1.; y += delta)
            for (double x = 0.; x < scale - y - 1; x += delta) {
                f.push_back(B.jittered(x, y, 1.1, PS2D_PHASE_FLUID));
                f.push_back(B.jittered(-x, y, 1.1, PS2D_PHASE_FLUID));
            }
        B.group(f, 1, false, false);
        const double posn[2] = {0, 0};
        if (B.err == PS_OK) B.err = ps2d_create_fluid_emitter(B.c, posn, scale * 4, 0, 0., 0.);
    } else if (k == "w") {  // initWreckingBall, :1218-1286
        brick_wall(B, 8, 2, 30., 1, 1);
        const double scale = 6., delta = .4, num = 1.;
        std::vector<Part> f;
        const double start = S.xb[0] + 1;
        for (double x = start; x < start + (scale / num); x += delta)
            for (double y = 0; y < 1.2 * scale; y += delta) f.push_back(B.jittered(x, y, 1, PS2D_PHASE_FLUID));
        B.group(f, 2.5, false, false);
        const uint32_t idx = (uint32_t)B.count;
        B.add({Part(10, 50, 0)});
        std::vector<Sdf> data;
        std::vector<Part> v;
        const double bx = 57, by = 50;
        v.push_back(Part(bx, by, 1000));
        for (double a = 0; a <= 360; a += 30) {
            const double vx = std::cos(D2R(a)), vy = std::sin(D2R(a));
            v.push_back(Part(vx * RAD + bx, vy * RAD + by, 1000));
            data.push_back(Sdf{vx, vy, RAD * 1.5});
        }
        data.push_back(Sdf{0., 0., -1.});  // SDFData()
        B.body(v, data);
        B.dist(idx, idx + 1);
    }
}
}  // namespace

extern "C" const char *ps2d_scene_name(const char *key) {
    const SceneInfo *s = find_scene(key);
    return s ? s->name : nullptr;
}

extern "C" int ps2d_build_scene(const char *key, int device, uint64_t max_particles, Ps2dCtx **out) {
    // the app's stream: seed 1 (never seeded), the constructor's WRECKING_BALL already built
    return ps2d_build_scene_from(key, device, max_particles, 1, PS2D_APP_START_DRAWS, out);
}

extern "C" int ps2d_build_scene_from(const char *key, int device, uint64_t max_particles, uint32_t seed, uint64_t draws_consumed, Ps2dCtx **out) {
    if (!out) { ps_set_error("ps2d_build_scene: null output"); return PS_ERR_INVALID; }
    *out = nullptr;
    const SceneInfo *S = find_scene(key);
    if (!S) {
        ps_set_error("ps2d_build_scene: unknown scene key");
        return PS_ERR_INVALID;
    }
    Ps2dParams P;
    ps2d_default_params(&P);
    P.x_bounds[0] = S->xb[0]; P.x_bounds[1] = S->xb[1]; P.y_bounds[0] = S->yb[0]; P.y_bounds[1] = S->yb[1];
    Builder B;
    int r = ps2d_create(device, &P, max_particles ? max_particles : S->room + 2048, &B.c);
    if (r != PS_OK) return r;
    ps2d_seed_rand(B.c, seed, draws_consumed);
    build(B, key, *S);
    if (B.err != PS_OK) { ps2d_destroy(B.c); return B.err; }
    *out = B.c;
    return PS_OK;
}
