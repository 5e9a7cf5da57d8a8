/*
 * include/psolver2d.h — C ABI of the 2-D double-precision path of libpsolver.so: one Simulation::tick of the
 * reference's CPU application (ebirenbaum/ParticleSolver cpu/src/simulation.cpp:115-369) on the GPU, for the
 * constraint groups an all-fluid scene exercises (config C1 = CPU scene 6, two-fluid Rayleigh-Taylor):
 *   prediction                         Particle::guess                    cpu/src/particle.h:56-58
 *   wall constraints + jitter          BoundaryConstraint::project        cpu/src/constraint/boundaryconstraint.cpp:14-93
 *   PBF density constraint per fluid   TotalFluidConstraint::project      cpu/src/constraint/totalfluidconstraint.cpp:41-115
 *   velocity update + sleeping         Particle::confirmGuess             cpu/src/particle.h:60-65
 * in the reference's order: 3 solver iterations of [every wall constraint in list order, then fluid 0, fluid 1, ...],
 * each fluid Jacobi inside and Gauss-Seidel against the previous one.  The wall jitter comes from
 * the same glibc rand() stream the reference draws from (ps2d_seed_rand positions it), so a tick reproduces the
 * reference's to rounding (double precision; sums run in ascending particle index like the reference's O(N^2) loops).
 *
 * The reference keeps this state in `Simulation` (QList<Particle*> + TotalFluidConstraint objects); there is no
 * extern "C" boundary on its CPU side, so this header defines one.  Not on this path (PS_ERR_STATE): SOLID / GAS
 * particles, contact, distance and shape constraints — SURVEY §8(f) "next" rows.  No CPU fallback.
 */
#ifndef PSOLVER2D_H
#define PSOLVER2D_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* phase codes of the CPU application — cpu/src/particle.h:10-15 */
#define PS2D_PHASE_SOLID 0
#define PS2D_PHASE_FLUID 1
#define PS2D_PHASE_GAS 2

typedef struct Ps2dParams {
    double x_bounds[2];          /* m_xBoundaries  (scene 6: -8, 8   simulation.cpp:898) */
    double y_bounds[2];          /* m_yBoundaries  (scene 6: -8, 40  simulation.cpp:899) */
    double gravity[2];           /* m_gravity      (0, -9.8) */
    uint32_t solver_iterations;  /* SOLVER_ITERATIONS 3, simulation.h:11 */
} Ps2dParams;

typedef struct Ps2dCtx Ps2dCtx;

enum { PS2D_ARR_P = 0, PS2D_ARR_V = 1, PS2D_ARR_EP = 2, PS2D_ARR_LAMBDA = 3 }; /* double[2n], double[2n], double[2n], double[n] */

void ps2d_default_params(Ps2dParams *p);
int ps2d_create(int device, const Ps2dParams *params, uint64_t max_particles, Ps2dCtx **out);
int ps2d_destroy(Ps2dCtx *ctx);
/* Simulation::createFluid(particles, density) (simulation.cpp:431-452): appends n FLUID particles as one more fluid
 * (one TotalFluidConstraint) of rest density `density`; p2 / v2: double[2n], inv_mass: double[n], all non-zero. */
int ps2d_create_fluid(Ps2dCtx *ctx, const double *p2, const double *v2, const double *inv_mass, uint64_t n, double density);
/* position of the glibc rand() stream the wall jitter is drawn from: srand(seed), then `skip` draws already consumed
 * (the reference never seeds — seed 1 — and its scene builders consume draws before the first tick) */
int ps2d_seed_rand(Ps2dCtx *ctx, uint32_t seed, uint64_t skip);
uint64_t ps2d_rand_calls(Ps2dCtx *ctx);              /* draws consumed so far, including `skip` */
int ps2d_tick(Ps2dCtx *ctx, double seconds);         /* Simulation::tick(seconds); the reference's app uses .01 (view.cpp:197) */
uint64_t ps2d_num_particles(Ps2dCtx *ctx);
uint32_t ps2d_last_num_boundary_constraints(Ps2dCtx *ctx);
int ps2d_download(Ps2dCtx *ctx, int which, double *host);
int ps2d_kinetic_energy(Ps2dCtx *ctx, double *out);  /* Simulation::getKineticEnergy, simulation.cpp:1293-1303 */
uint32_t ps2d_launches_per_tick(Ps2dCtx *ctx);
#ifdef __cplusplus
}
#endif
#endif /* PSOLVER2D_H */
