/*
 * include/simulation2d.h — header-only C++ host class with the public interface of the reference CPU application's
 * `Simulation` (cpu/src/simulation.h:24-42 SimulationType, :45-72 class) over the ps2d_* C ABI of libpsolver.so, so that
 * code written against the reference's class (its View: cpu/src/view.cpp:121-202) compiles against this one:
 *     psb200::Simulation sim;  sim.init(psb200::FLUID_TEST);  sim.tick(.01);  sim.getKineticEnergy();
 * Like the reference, the constructor builds WRECKING_BALL, and every init() continues the process-wide rand() stream
 * where earlier scenes and ticks left it (the reference never seeds), so a session replays the reference's bit for bit.
 * Not here: draw() / resize() (viewer), the individual init*() members (init(type) selects them).
 */
#ifndef SIMULATION2D_H
#define SIMULATION2D_H
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "ps_scenes2d.h"
#include "psolver.h"

namespace psb200 {

/* cpu/src/simulation.h:24-42, same enumerators in the same order */
enum SimulationType {
    FRICTION_TEST, SDF_TEST, GRANULAR_TEST, STACKS_TEST, WALL_TEST, PENDULUM_TEST, ROPE_TEST, FLUID_TEST, FLUID_SOLID_TEST,
    GAS_ROPE_TEST, WATER_BALLOON_TEST, CRADLE_TEST, NUM_SIMULATION_TYPES, SMOKE_OPEN_TEST, SMOKE_CLOSED_TEST, VOLCANO_TEST,
    WRECKING_BALL
};

class Simulation {
public:
    explicit Simulation(int device = 0) : m_device(device) { init(WRECKING_BALL); debug = true; }  /* simulation.cpp:11-16 */
    virtual ~Simulation() { if (m_ctx) ps2d_destroy(m_ctx); }
    Simulation(const Simulation &) = delete;
    Simulation &operator=(const Simulation &) = delete;

    /* Simulation::init(type): clears the scene and builds the chosen one (simulation.cpp:63-112) */
    void init(SimulationType type) {
        const uint64_t consumed = m_ctx ? ps2d_rand_calls(m_ctx) : 0;
        Ps2dCtx *next = nullptr;
        check(ps2d_build_scene_from(key_of(type), m_device, 0, 1, consumed, &next));
        if (m_ctx) ps2d_destroy(m_ctx);
        m_ctx = next;
    }
    void tick(double seconds) { check(ps2d_tick(m_ctx, seconds)); }
    void mousePressed(double x, double y) { check(ps2d_mouse_pressed(m_ctx, x, y)); }

    int getNumParticles() { return (int)ps2d_num_particles(m_ctx); }
    double getKineticEnergy() { double e = 0; check(ps2d_kinetic_energy(m_ctx, &e)); return e; }
    bool debug;

    /* beyond the reference's interface */
    std::vector<double> getPositions() { std::vector<double> p(2 * (size_t)getNumParticles()); check(ps2d_download(m_ctx, PS2D_ARR_P, p.data())); return p; }
    Ps2dCtx *context() const { return m_ctx; }

    static const char *key_of(SimulationType t) {  /* the app's key for the scene, cpu/src/view.cpp:129-177 */
        switch (t) {
            case GRANULAR_TEST: return "1"; case STACKS_TEST: return "2"; case WALL_TEST: return "3"; case PENDULUM_TEST: return "4";
            case ROPE_TEST: return "5"; case FLUID_TEST: return "6"; case FLUID_SOLID_TEST: return "7"; case GAS_ROPE_TEST: return "8";
            case FRICTION_TEST: return "9"; case WATER_BALLOON_TEST: return "0"; case CRADLE_TEST: return "n"; case SMOKE_OPEN_TEST: return "s";
            case SMOKE_CLOSED_TEST: return "d"; case SDF_TEST: return "."; case VOLCANO_TEST: return "v"; case WRECKING_BALL: return "w";
            default: return "2";  /* Simulation::init's default branch builds initBoxes (simulation.cpp:104-105) */
        }
    }

private:
    void check(int r) { if (r != PS_OK) throw std::runtime_error(ps_last_error()); }
    int m_device;
    Ps2dCtx *m_ctx = nullptr;
};

}  // namespace psb200
#endif /* SIMULATION2D_H */
