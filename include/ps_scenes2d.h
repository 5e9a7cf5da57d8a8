/*
 * include/ps_scenes2d.h — the scene builders of the reference's 2-D CPU application (Simulation::init*,
 * cpu/src/simulation.cpp:659-1286) over the ps2d_* C ABI, selected by the app's key bindings (cpu/src/view.cpp:129-177):
 *   "1" GRANULAR  "2" STACKS  "3" WALL  "4" PENDULUM  "5" ROPE  "6" FLUID  "7" FLUID_SOLID  "8" GAS_ROPE  "9" FRICTION
 *   "0" WATER_BALLOON  "n" CRADLE  "s" SMOKE_OPEN  "d" SMOKE_CLOSED  "." SDF  "v" VOLCANO  "w" WRECKING_BALL
 * A scene is reproduced bit for bit, jitter included: the builders draw from the context's glibc rand() stream in the
 * reference's order, starting where the app's stream stands when a key is pressed — `Simulation::Simulation()` has
 * built WRECKING_BALL once already (simulation.cpp:11-16), which consumes 541 draws.
 */
#ifndef PS_SCENES2D_H
#define PS_SCENES2D_H
#include "psolver2d.h"
#ifdef __cplusplus
extern "C" {
#endif
/* draws the app's constructor consumed before the first key press */
#define PS2D_APP_START_DRAWS 541
/* Simulation::init(type): creates a context with the scene's bounds / gravity and builds the scene into it.
 * max_particles 0 = scene size + room for emitted particles. */
int ps2d_build_scene(const char *key, int device, uint64_t max_particles, Ps2dCtx **out);
/* the same with the position of the rand() stream given explicitly: srand(seed), `draws_consumed` draws already taken
 * (Simulation::init continues the stream wherever earlier scenes and ticks left it) */
int ps2d_build_scene_from(const char *key, int device, uint64_t max_particles, uint32_t seed, uint64_t draws_consumed, Ps2dCtx **out);
const char *ps2d_scene_name(const char *key); /* "GRANULAR_TEST", ...; NULL for an unknown key */
#ifdef __cplusplus
}
#endif
#endif /* PS_SCENES2D_H */
