/*
 * include/psolver2d.h — C ABI of the 2-D double-precision path of libpsolver.so: one Simulation::tick of the
 * reference's CPU application (ebirenbaum/ParticleSolver cpu/src/simulation.cpp:115-369) on the GPU, with every
 * constraint group the reference's ITERATIVE solver has:
 *   prediction, mass scaling           Particle::guess / scaleMass        cpu/src/particle.h:56-58,67-73
 *   contact discovery                  simulation.cpp:165-225 (O(N^2) pairs, then walls, per particle)
 *   wall constraints + jitter+friction BoundaryConstraint::project        cpu/src/constraint/boundaryconstraint.cpp:14-93
 *   particle contacts                  ContactConstraint::project         cpu/src/constraint/contactconstraint.cpp:13-42
 *   rigid (SDF) contacts + friction    RigidContactConstraint::project    cpu/src/constraint/rigidcontactconstraint.cpp:13-96
 *   distance constraints               DistanceConstraint::project        cpu/src/constraint/distanceconstraint.cpp:20-40
 *   PBF density constraint per fluid   TotalFluidConstraint::project      cpu/src/constraint/totalfluidconstraint.cpp:41-158
 *   gas constraint                     GasConstraint::project             cpu/src/constraint/gasconstraint.cpp:30-116
 *   shape matching per rigid body      TotalShapeConstraint::project, Body::updateCOM
 *                                      cpu/src/constraint/totalshapeconstraint.cpp:14-24,80-87, cpu/src/solver/particle.cpp:15-57
 *   velocity update + sleeping         Particle::confirmGuess             cpu/src/particle.h:60-65
 *   smoke emitter (particle injection) OpenSmokeEmitter::tick             cpu/src/opensmokeemitter.cpp:17-29
 *   fluid emitter (emission, freezing) FluidEmitter::tick                 cpu/src/fluidemitter.cpp:13-79
 * in the reference's order: 3 solver iterations of [CONTACT list, STANDARD list, SHAPE list].  The reference projects
 * each list sequentially (Gauss-Seidel); here a list is LEVEL-SCHEDULED on the device: a constraint's level is one more
 * than the highest level of the earlier constraints that share a particle with it, constraints of one level touch
 * disjoint particles and are projected in parallel, levels run in order.  That executes exactly the reference's
 * sequence of updates per particle, so a tick reproduces the reference's to rounding (double precision, no FMA
 * contraction; the wall jitter comes from the same glibc rand() stream, positioned with ps2d_seed_rand).
 *
 * The reference keeps this state in `Simulation` (QList<Particle*>, Body, Constraint objects); there is no extern "C"
 * boundary on its CPU side, so this header defines one whose entry points mirror the reference's constructors and
 * Simulation::create* helpers.  The stabilization pass (the reference's compile-time option USE_STABILIZATION, off in its build) is Ps2dParams.stabilization_iterations.
 * Not on this path: the UMFPACK matrix
 * solver (dead under #define ITERATIVE), the emitters' display-only tracer particles.  No CPU fallback.
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
/* most particle-particle contacts one particle can take part in during a tick (2-D discs of one size: 6 touch) */
#define PS2D_MAX_CONTACTS 14

typedef struct Ps2dParams {
    double x_bounds[2];          /* m_xBoundaries  (scene 6: -8, 8   simulation.cpp:898) */
    double y_bounds[2];          /* m_yBoundaries  (scene 6: -8, 40  simulation.cpp:899) */
    double gravity[2];           /* m_gravity      (0, -9.8) */
    uint32_t solver_iterations;  /* SOLVER_ITERATIONS 3, simulation.h:11 */
    uint32_t stabilization_iterations; /* 0 = the reference as built; 2 = built with its option USE_STABILIZATION (commented out in
                                    simulation.h:17; STABILIZATION_ITERATIONS 2, :18): that many passes over stabile copies of the tick's
                                    rigid contacts and wall constraints, which move p as well as ep, before the solver iterations
                                    (simulation.cpp:249-271).  Occupies former padding: sizeof(Ps2dParams) is unchanged. */
} Ps2dParams;

typedef struct Ps2dCtx Ps2dCtx;

/* array selectors of ps2d_download */
enum {
    PS2D_ARR_P = 0, PS2D_ARR_V = 1, PS2D_ARR_EP = 2,   /* double[2n]: position, velocity, predicted position */
    PS2D_ARR_LAMBDA = 3,                               /* double[n]  of the last fluid / gas constraint projected */
    PS2D_ARR_F = 4,                                    /* double[2n] Particle::f */
    PS2D_ARR_COUNTS = 5,                               /* uint32[n]  constraints per particle of the last tick */
    PS2D_ARR_TMASS = 6, PS2D_ARR_IMASS = 7,            /* double[n]  height-scaled / plain inverse mass */
    PS2D_ARR_SFRICTION = 8, PS2D_ARR_KFRICTION = 9,    /* double[n] */
    PS2D_ARR_SDF_DIST = 10,                            /* double[n]  SDFData::distance (-1: none) */
    PS2D_ARR_RS = 11, PS2D_ARR_SDF_GRAD = 12,          /* double[2n] Body::rs, SDFData::gradient */
    PS2D_ARR_PHASE = 13, PS2D_ARR_BOD = 14,            /* int32[n] */
    PS2D_ARR_GROUP = 15                                /* int32[n]   STANDARD index of the particle's fluid / gas constraint, -1 none */
};

void ps2d_default_params(Ps2dParams *p);
int ps2d_create(int device, const Ps2dParams *params, uint64_t max_particles, Ps2dCtx **out);
int ps2d_destroy(Ps2dCtx *ctx);

/* ---- particles and constraints, as the reference's constructors take them ---- */
/* n x Particle(pos, vel, mass, phase) appended to the particle list (particle.h:31-52).  inv_mass 0 = immovable.
 * bod: body tag used to disable collisions between solids (-1 none, < -1 free tags such as the rope's -2); NULL = -1.
 * s_friction / k_friction: NULL = 0.  Returns the index of the first appended particle in *first (may be NULL). */
int ps2d_add_particles(Ps2dCtx *ctx, const double *p2, const double *v2, const double *inv_mass, const int32_t *phase, const int32_t *bod,
                       const double *s_friction, const double *k_friction, uint64_t n, uint64_t *first);
/* new DistanceConstraint(d, i1, i2) appended to the STANDARD list; d < 0: the current distance of the two particles,
 * DistanceConstraint(i1, i2, particles) (distanceconstraint.cpp:3-13) */
int ps2d_add_distance_constraint(Ps2dCtx *ctx, uint32_t i1, uint32_t i2, double d);
/* new TotalFluidConstraint(density, indices) / new GasConstraint(density, indices, open) appended to the STANDARD
 * list; the particles must exist, be FLUID / GAS and have non-zero inverse mass.  *standard_index: position in the list. */
int ps2d_add_fluid_constraint(Ps2dCtx *ctx, const uint32_t *indices, uint64_t n, double density, uint32_t *standard_index);
int ps2d_add_gas_constraint(Ps2dCtx *ctx, const uint32_t *indices, uint64_t n, double density, int open, uint32_t *standard_index);
/* A rigid body over the existing SOLID particles [first, first + n) with explicit shape-matching state: r vectors
 * (double[2n]), SDF samples (gx, gy, distance; double[3n]), total inverse mass, centre, angle, stiffness — what `Body`
 * holds (particle.h:126-139).  Used to restore a checkpoint; ps2d_create_rigid_body derives these like the reference. */
int ps2d_restore_rigid_body(Ps2dCtx *ctx, uint32_t first, uint32_t n, const double *rs2, const double *sdf3, double inv_mass, const double *center2,
                            double angle, double stiffness, uint32_t *body_index);

/* ---- Simulation::create* (simulation.cpp:369-452) ---- */
/* createRigidBody(verts, sdfData): appends n SOLID particles as one body; centre of mass, r vectors and total mass are
 * computed as Body::updateCOM(false) / computeRs do; shape stiffness 1. */
int ps2d_create_rigid_body(Ps2dCtx *ctx, const double *p2, const double *v2, const double *inv_mass, const double *s_friction, const double *k_friction,
                           const double *sdf3, uint64_t n, uint32_t *body_index);
/* createFluid(particles, density): appends n FLUID particles as one more fluid (one TotalFluidConstraint) */
int ps2d_create_fluid(Ps2dCtx *ctx, const double *p2, const double *v2, const double *inv_mass, uint64_t n, double density);
/* createGas(particles, density, open): appends n GAS particles as one GasConstraint */
int ps2d_create_gas(Ps2dCtx *ctx, const double *p2, const double *v2, const double *inv_mass, uint64_t n, double density, int open, uint32_t *standard_index);
/* createSmokeEmitter(posn, particlesPerSec, gs): every 1/rate seconds of simulated time one GAS particle of mass 1 is
 * appended at posn and added to the gas constraint at `standard_index` (no injection if it is UINT32_MAX, like gs == NULL).
 * The emitter's display-only tracer particles are not simulated.  timer: time already accumulated (0 for a new one). */
int ps2d_create_smoke_emitter(Ps2dCtx *ctx, const double *posn2, double rate, uint32_t standard_index, double timer);

/* createFluidEmitter(posn, particlesPerSec, fs) (simulation.cpp:459-461; FluidEmitter::tick, fluidemitter.cpp:13-79 — the
 * VOLCANO scene): for the first 5 seconds one FLUID particle of mass 1 and velocity (frand(), 1) is emitted at posn every
 * 1/rate seconds into the fluid constraint at `standard_index`; fluid particles that are slow (|v| < .06) and low (y <= 5)
 * count down Particle::t and then freeze into immovable SOLID particles that leave the fluid.  timer / total_timer: time
 * already accumulated (0, 0 for a new emitter). */
int ps2d_create_fluid_emitter(Ps2dCtx *ctx, const double *posn2, double rate, uint32_t standard_index, double timer, double total_timer);

/* ---- state access (checkpoint / restore) ---- */
int ps2d_set_particle_timers(Ps2dCtx *ctx, const double *t);   /* Particle::t of every particle (4 when created) */
int ps2d_get_particle_timers(Ps2dCtx *ctx, double *t);
int ps2d_set_forces(Ps2dCtx *ctx, const double *f2);     /* Particle::f of every particle (read by the next tick's prediction) */
int ps2d_body_state(Ps2dCtx *ctx, uint32_t body, double *center2, double *angle);
uint32_t ps2d_num_bodies(Ps2dCtx *ctx);

/* Ps2dParams.stabilization_iterations of an existing context (0 = the reference as built, 2 = built with USE_STABILIZATION) */
int ps2d_set_stabilization_iterations(Ps2dCtx *ctx, uint32_t iterations);

/* position of the glibc rand() stream the wall jitter is drawn from: srand(seed), then `skip` draws already consumed
 * (the reference never seeds — seed 1 — and its scene builders consume draws before the first tick) */
int ps2d_seed_rand(Ps2dCtx *ctx, uint32_t seed, uint64_t skip);
uint64_t ps2d_rand_calls(Ps2dCtx *ctx);              /* draws consumed so far, including `skip` */
int ps2d_rand(Ps2dCtx *ctx, int *out);               /* one rand() of that stream (the scene builders' frand() jitter) */
int ps2d_mouse_pressed(Ps2dCtx *ctx, double x, double y);  /* Simulation::mousePressed: v += 7 normalize(point - p), every particle */
int ps2d_tick(Ps2dCtx *ctx, double seconds);         /* Simulation::tick(seconds); the reference's app uses .01 (view.cpp:197) */
uint64_t ps2d_num_particles(Ps2dCtx *ctx);
uint32_t ps2d_last_num_boundary_constraints(Ps2dCtx *ctx);  /* wall constraints of the last tick that draw jitter (fluid / gas particles) */
uint32_t ps2d_last_num_contact_constraints(Ps2dCtx *ctx);   /* size of the last tick's CONTACT list: pairs + walls */
uint32_t ps2d_last_num_levels(Ps2dCtx *ctx);                /* depth of the last tick's CONTACT level schedule */
int ps2d_download(Ps2dCtx *ctx, int which, void *host);
/* checkpoints: everything a run needs to continue bit-identically (particles, bodies, the STANDARD list, emitters, the
 * rand() stream).  The reference has no persistence. */
int ps2d_save(Ps2dCtx *ctx, const char *path);
int ps2d_load(const char *path, int device, Ps2dCtx **out);
int ps2d_kinetic_energy(Ps2dCtx *ctx, double *out);  /* Simulation::getKineticEnergy, simulation.cpp:1293-1303 */
uint32_t ps2d_launches_per_tick(Ps2dCtx *ctx);
#ifdef __cplusplus
}
#endif
#endif /* PSOLVER2D_H */
