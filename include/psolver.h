/*
 * include/psolver.h — C ABI of libpsolver.so, the B200-native unified particle solver step.
 *
 * Drop-in boundary.  In the reference (ebirenbaum/ParticleSolver) the solver step sits behind the
 * extern "C" wrapper layer declared in gpu/src/cuda/wrappers.cuh:12-97, gpu/src/cuda/util.cuh:6-25 and
 * gpu/src/cuda/shared_variables.cuh:13-36 and is called only from ParticleSystem
 * (gpu/src/particlesystem.cpp).  This library exports
 *   (1) a context-based interface (ps_*): opaque context, 64-bit sizes, error codes, one CUDA stream per
 *       context, whole-step entry point ps_step() == ParticleSystem::update (particlesystem.cpp:144-246)
 *       plus one entry point per stage for parity checks; and
 *   (2) the reference's own wrapper names and argument orders (include/ps_reference_abi.h), implemented
 *       on the same kernels, so the reference's ParticleSystem links against it unchanged.
 * Plain pointers and sizes only; no torch / thrust / C++ types cross this boundary.
 * There is no CPU fallback: every entry point returns PS_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef PSOLVER_H
#define PSOLVER_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* phase codes — gpu/src/cuda/shared_variables.cuh:4-9 */
#define PS_PHASE_NO_COLLIDE (-1)
#define PS_PHASE_FLUID 0
#define PS_PHASE_GAS 1
#define PS_PHASE_CLOTH 2
#define PS_PHASE_SOLID 3
#define PS_PHASE_RIGID 4 /* + body id */

enum {
    PS_OK = 0,
    PS_ERR_INVALID = 1,  /* bad argument (null, non power-of-two grid, ...) */
    PS_ERR_CUDA = 2,     /* CUDA runtime / cuRAND failure; see ps_last_error() */
    PS_ERR_CAPACITY = 3, /* append beyond max_particles (the reference silently drops: particlesystem.cpp:311,335) */
    PS_ERR_STATE = 4     /* call order violated (e.g. neighbour solve before a grid build) */
};

/* Superset of SimParams (gpu/src/cuda/kernel.cuh:9-22) + what ParticleSystem keeps beside it
 * (bounds, iteration count: gpu/src/particlesystem.h:113-116).  Defaults: ps_default_params(). */
typedef struct PsParams {
    float gravity[3];            /* (0,-9.8,0)  particlesystem.cpp:65 */
    float global_damping;        /* 1.0, unused by the reference's live kernels */
    float particle_radius;       /* 0.25        particleapp.cpp:24 */
    uint32_t grid_size[3];       /* powers of two; 64^3 in the reference (particleapp.cpp:25) */
    float world_origin[3];       /* 0           particlesystem.cpp:61 */
    float cell_size[3];          /* 2*radius    particlesystem.cpp:62-63 */
    int32_t min_bounds[3];       /* scene box, integer like the reference's int3 */
    int32_t max_bounds[3];
    uint32_t solver_iterations;  /* 5           particleapp.cpp:38 */
    float omega;                 /* SOR factor on every Jacobi-averaged delta: contacts / numNeighbors (K5), PBF delta-p /
                                    (rho0 + numNeighbors) (K7), distance constraints / occurrences (K9); 1.0 == reference */
    uint32_t flags;              /* PS_FLAG_* */
    uint32_t neighbor_list_rows; /* neighbour lists kept between the two PBF passes: rows (128 B: one entry for each of a warp's 32
                                    particles) of every warp's list region.  512 (default) >= the 500-neighbour cap: a list always fits;
                                    2 KB of address space per particle, of which only the rows actually used are ever touched.  With
                                    fewer rows a warp whose longest list does not fit falls back to a second grid walk (results
                                    identical).  0 = keep no lists.  Multiple of 8. */
} PsParams;

#define PS_FLAG_NONE 0u
#define PS_FLAG_ZERO_NONFLUID_LAMBDA 1u /* deviation: lambda of non-fluid slots reads 0 instead of a stale value */
#define PS_FLAG_GAS 2u /* not in the reference's GPU solver (its kernels ignore phase GAS): GAS particles take part in the
                          fluid density constraint with their own rest density and rise — they are predicted with gravity x -0.2, the
                          CPU app's ALPHA (cpu/src/simulation.h:21, simulation.cpp:144).  Parity unpinned in 3-D. */

#define PS_FLAG_SELF_COLLISION 4u /* not in the reference: there all particles of one phase > SOLID skip each other in the contact
                          pass (integration_kernel.cuh:336-337), so its cloth never self-collides.  With this flag two
                          such particles collide (contact + friction, like particles of different bodies) when both carry
                          distance constraints and no distance constraint joins them: a cloth or rope touches itself, its
                          constrained neighbours and the members of shape-matched bodies still skip each other.  Unpinned. */
#define PS_FLAG_STAGED_LAMBDA 8u /* K6 (PBF lambda) by the TMA-staged kernel (csrc/ps_fluid_staged.cu: the neighbour rows of a CTA are
                          copied into shared memory by cp.async.bulk) instead of the grid walk through L1.  Same results, bit for
                          bit; measured slower than the walk on B200, kept as the measured alternative.  Needs neighbour
                          lists (neighbor_list_rows != 0). */

typedef struct PsCtx PsCtx;

/* array selectors for ps_download / ps_upload / ps_device_ptr */
enum {
    PS_ARR_POS = 0,          /* float4[n]  current positions (the reference's VBO) */
    PS_ARR_VEL = 1,          /* float4[n]  V            integration.cu:23 */
    PS_ARR_PREV = 2,         /* float4[n]  Xstar        shared_variables.cu:8 */
    PS_ARR_INV_MASS = 3,     /* float[n]   W            shared_variables.cu:9 */
    PS_ARR_PHASE = 4,        /* int[n]     phase        shared_variables.cu:10 */
    PS_ARR_REST_DENSITY = 5, /* float[n]   ros          integration.cu:27 */
    PS_ARR_HASH = 6,         /* uint[n]    sorted cell keys      m_dGridParticleHash */
    PS_ARR_INDEX = 7,        /* uint[n]    sorted order          m_dGridParticleIndex */
    PS_ARR_CELL_START = 8,   /* uint[cells] 0xffffffff = empty   m_dCellStart */
    PS_ARR_CELL_END = 9,     /* uint[cells] valid where start != 0xffffffff  m_dCellEnd */
    PS_ARR_SORTED_POS = 10,  /* float4[n]  m_dSortedPos (download only; ps_device_ptr gives the resident array, whose .w holds the
                                sorted slot's bit pattern instead of pos.w) */
    PS_ARR_SORTED_INV_MASS = 11,
    PS_ARR_SORTED_PHASE = 12,
    PS_ARR_LAMBDA = 13,      /* float[n] by sorted slot  integration.cu:24 */
    PS_ARR_NUM_NEIGHBORS = 14, /* uint[n] by sorted slot integration.cu:30 */
    PS_ARR_RANDS = 15,       /* float[iterations*6] wall-jitter uniforms of the last step */
    PS_ARR_OCCURRENCES = 16, /* uint[n]    solver.cu:41 */
    PS_ARR_CELL_BEGIN = 17,  /* uint[cells+1] dense lower-bound table (internal; exposed for tests) */
    PS_ARR_NEIGHBOR_ROWS = 18 /* uint[ceil(n/32)] per-warp list status of the last lambda pass (diagnostics): 0 = no active fluid
                                 particle, 1 = list valid, 0xffffffff = the warp has no list and its delta-p pass walks the grid again
                                 (csrc/ps_fluid_lists.cuh) */
};

void ps_default_params(PsParams *p);
const char *ps_last_error(void);
const char *ps_version(void);

int ps_create(int device, const PsParams *params, uint64_t max_particles, PsCtx **out);
int ps_destroy(PsCtx *ctx);
int ps_set_params(PsCtx *ctx, const PsParams *params);   /* setParameters, integration.cu:96-100 */
int ps_get_params(PsCtx *ctx, PsParams *out);
uint64_t ps_num_particles(PsCtx *ctx);
uint64_t ps_num_cells(PsCtx *ctx);

/* host pointers, copied.  pos4/vel4: float[4n]; inv_mass, rest_density: float[n]; phase: int[n].
 * == the VBO write + appendIntegrationParticle + appendPhaseAndMass + appendSolverParticle of
 * ParticleSystem::addParticleMultiple (particlesystem.cpp:333-348). */
int ps_append_particles(PsCtx *ctx, const float *pos4, const float *vel4, const float *inv_mass,
                        const float *rest_density, const int32_t *phase, uint64_t n);
/* addDistanceConstraint (solver.cu:125-156): idx_pairs uint[2m], rest float[m] */
int ps_add_distance_constraints(PsCtx *ctx, const uint32_t *idx_pairs, const float *rest, uint64_t m);
/* addPointConstraint (solver.cu:108-123): idx uint[p], xyz float[3p] */
int ps_add_point_constraints(PsCtx *ctx, const uint32_t *idx, const float *xyz, uint64_t p);

/* constraint lists as held by the context, in insertion order (host copies; any pointer may be NULL) */
uint64_t ps_num_distance_constraints(PsCtx *ctx);
uint64_t ps_num_point_constraints(PsCtx *ctx);
int ps_copy_distance_constraints(PsCtx *ctx, uint32_t *idx_pairs, float *rest);
int ps_copy_point_constraints(PsCtx *ctx, uint32_t *idx, float *xyz);

/* One whole step == ParticleSystem::update(dt) (particlesystem.cpp:144-246), asynchronous on the context's
 * stream, replayed from a CUDA graph after the first call.  dt is clamped to 0.05 like the reference (:149). */
int ps_step(PsCtx *ctx, float dt);
int ps_sync(PsCtx *ctx);
/* milliseconds of device time of the last ps_step (CUDA events on the context's stream); syncs. */
int ps_last_step_ms(PsCtx *ctx, float *ms);

/* One whole step for a host that owns positions and velocities (the reference's host reads them back through
 * copyArrayFromDevice / its GL buffer, particlesystem.cpp:122-142,248-262): the step's inputs come from host memory, its result
 * goes to host memory, float4 per particle each; any of the four pointers may be NULL (that array stays on the device / is not
 * delivered).  Asynchronous: the transfers run on two copy streams through double-buffered staging frames, so the upload of
 * the next call and the download of the previous one overlap the solver (csrc/ps_stream_io.cu).  Host memory should be pinned,
 * input buffers must stay untouched until the call after next has been made (or ps_io_wait(ctx, 0) returned), output buffers
 * are valid after ps_io_wait.  ps_io_wait(ctx, k) blocks until the outputs of the call made k calls ago (0 = the last) landed. */
int ps_step_streamed(PsCtx *ctx, float dt, const float *pos_in, const float *vel_in, float *pos_out, float *vel_out);
int ps_io_wait(PsCtx *ctx, uint32_t calls_back);
/* the two halves of ps_step_streamed for callers that issue the step themselves (a slab context's step is a sequence of stage calls
 * and exchanges): inputs before the step, outputs after it, over the OWNED particles; one begin and one end per step */
int ps_io_begin(PsCtx *ctx, const float *pos_in, const float *vel_in);
/* starts the transfer of the NEXT step's inputs now (call it before issuing the current step): needed where issuing a step blocks
 * the host (ps_comm_step waits for its neighbours' counts), so that the upload still overlaps the step; the next ps_io_begin
 * must name the same buffers */
int ps_io_prefetch(PsCtx *ctx, const float *pos_in, const float *vel_in);
int ps_io_end(PsCtx *ctx, float *pos_out, float *vel_out);

/* Device-time a region of work on the context's stream with CUDA events (stop synchronises). */
int ps_timer_start(PsCtx *ctx);
int ps_timer_stop(PsCtx *ctx, float *ms);

/* Instrumented step for roofline reporting: the same launches as ps_step, issued eagerly with a CUDA event after
 * every stage.  stage_ms[PS_NUM_STAGES] receives the device time of each stage summed over the step's solver
 * iterations, stage_launches[PS_NUM_STAGES] (may be NULL) the number of kernel launches behind it. */
#define PS_NUM_STAGES 12
int ps_step_profiled(PsCtx *ctx, float dt, float *stage_ms, uint32_t *stage_launches);
const char *ps_stage_name(int stage); /* "predict","hash","sort","reorder","cell_table","contacts","lambda","delta_p","world","distance","point","velocity" */

/* Per-stage entry points, in the order update() calls them; each is asynchronous on the context's stream. */
int ps_begin_step(PsCtx *ctx);               /* draws this step's wall-jitter uniforms (cuRAND XORWOW, seed 1234) */
int ps_predict(PsCtx *ctx, float dt);        /* K1  integrateSystem               integration.cu:122-135 */
int ps_build_grid(PsCtx *ctx);               /* K2-K4 calcHash, sortParticles, reorderDataAndFindCellStart */
int ps_solve_contacts(PsCtx *ctx);           /* K5  collide                       integration.cu:338-386 */
int ps_solve_fluid(PsCtx *ctx);              /* K6+K7 solveFluids                 integration.cu:453-508 */
int ps_solve_fluid_lambda(PsCtx *ctx);       /* K6 alone  findLambdasD            integration_kernel.cuh:521-593 */
int ps_solve_fluid_delta(PsCtx *ctx);        /* K7 alone  solveFluidsD            integration_kernel.cuh:596-642 */
int ps_collide_world(PsCtx *ctx, uint32_t iteration); /* K8 collideWorld          integration.cu:319-336 */
int ps_solve_distance(PsCtx *ctx);           /* K9  solveDistanceConstraints      solver.cu:196-231 */
int ps_solve_point(PsCtx *ctx);              /* K10 solvePointConstraints         solver.cu:180-194 */
int ps_update_velocity(PsCtx *ctx, float dt);/* K11 calcVelocity                  integration.cu:416-426 */

/* Synchronous copies (they sync the context's stream first). count is in elements of the array's type
 * (float4 counts as 4 floats: pass 4*n). */
int ps_download(PsCtx *ctx, int which, void *host, uint64_t offset_elems, uint64_t count_elems);
int ps_upload(PsCtx *ctx, int which, const void *host, uint64_t offset_elems, uint64_t count_elems);
/* asynchronous variants on the context's stream; host memory should be pinned */
int ps_download_async(PsCtx *ctx, int which, void *host, uint64_t offset_elems, uint64_t count_elems);
int ps_upload_async(PsCtx *ctx, int which, const void *host, uint64_t offset_elems, uint64_t count_elems);
/* raw device pointer of an array (for zero-copy interop with a viewer or torch); NULL if unknown */
void *ps_device_ptr(PsCtx *ctx, int which);
/* the context's cudaStream_t, as an opaque pointer */
void *ps_stream(PsCtx *ctx);

/* number of kernel launches issued by the last ps_step (counted at capture time) */
uint32_t ps_launches_per_step(PsCtx *ctx);

/* ---- spatial slab decomposition (multi-GPU; one context per GPU; the exchange itself is the caller's: NCCL
 * send/recv of the record buffers between neighbouring ranks, see particlesolver_b200/slab.py) ----
 * A context owns the particles whose x lies in its slab [x_lo, x_hi) and holds them first in every array, followed by
 * `ghost` copies of the neighbours' particles near the two faces.  Ghosts take part in the grid build and are read as
 * neighbours (their lambda is computed locally, which is exact for ghosts within H of a face when the halo is 2H wide),
 * but they are never moved and never written back.  Use +-INFINITY for the outer faces of the first / last slab.
 * All buffers are DEVICE pointers of `capacity_records` records; counts are returned on the host (the calls
 * synchronise the context's stream).  Record order is ascending particle index, so runs are reproducible.
 * Slab contexts hold no distance / point constraints (PS_ERR_STATE otherwise) and are stepped stage by stage. */
#define PS_HALO_RECORD_BYTES 32u    /* float4 pos | float inv_mass | float rest_density | int32 phase | pad */
#define PS_MIGRANT_RECORD_BYTES 64u /* float4 pos | float4 prev | float4 vel | float inv_mass | float rest_density | int32 phase | pad */
/* records of the OWNED particles with x < x_lo + width (-> left_buf) and with x >= x_hi - width (-> right_buf) */
int ps_slab_pack_halo(PsCtx *ctx, float x_lo, float x_hi, float width, void *left_buf, void *right_buf, uint64_t capacity_records,
                      uint32_t counts[2]);
/* replaces the ghosts by the records received from the left and right neighbours (left first) */
int ps_slab_set_ghosts(PsCtx *ctx, const void *from_left, uint64_t n_left, const void *from_right, uint64_t n_right);
/* removes the owned particles with x < x_lo (-> left_buf) or x >= x_hi (-> right_buf), keeping the order of the rest;
 * drops the ghosts.  Call it after ps_predict: prev / vel travel with the particle. */
int ps_slab_pack_migrants(PsCtx *ctx, float x_lo, float x_hi, void *left_buf, void *right_buf, uint64_t capacity_records, uint32_t counts[2]);
/* appends received migrants as owned particles (left neighbour's first) */
int ps_slab_append_migrants(PsCtx *ctx, const void *from_left, uint64_t n_left, const void *from_right, uint64_t n_right);
/* Lambda exchange (the K6 -> K7 dependency across a face): between ps_solve_fluid_lambda and ps_solve_fluid_delta a rank
 * sends the lambdas of the particles of its last ps_slab_pack_halo — one float per record, in record order, counts[] = the
 * halo pack's — and hands the values received for its own ghosts (left neighbour's first) to ps_slab_set_ghost_lambda.
 * With the exchange the halo need only be H (+ drift) wide and K6 skips the ghosts (ps_slab_set_lambda_range(ctx, 1, -1)). */
int ps_slab_pack_lambda(PsCtx *ctx, void *left_buf, void *right_buf, uint64_t capacity_values, uint32_t counts[2]);
int ps_slab_set_ghost_lambda(PsCtx *ctx, const void *from_left, uint64_t n_left, const void *from_right, uint64_t n_right);
/* the two lambda messages as sinks of the lambda pass: ps_solve_fluid_lambda then fills them as it computes (no separate pack pass) and
 * ps_slab_pack_lambda on the same buffers only reports the counts.  All-FLUID contexts only; NULL, NULL, 0 switches it off. */
int ps_slab_set_lambda_sinks(PsCtx *ctx, void *left_buf, void *right_buf, uint64_t capacity_values);
/* lambda is computed for owned particles and for ghosts with x in [x_min, x_max] only (default: everywhere) */
int ps_slab_set_lambda_range(PsCtx *ctx, float x_min, float x_max);
/* load balancing: counts of the owned particles' x in `bins` (<= 65536) equal bins of [x_min, x_max) (values outside fall
 * into the end bins), copied to host_counts[bins].  The ranks' histograms summed are what the slabs are re-cut from. */
int ps_slab_x_histogram(PsCtx *ctx, float x_min, float x_max, uint32_t bins, uint64_t *host_counts);

/* Diagnostics of the current state, the quantities north_star's 1000-step parity bar is stated in: mean and largest density
 * error |rho_i / rho0_i - 1| over the fluid particles (rho_i: the lambda pass's estimate, integration_kernel.cuh:565-589, on a
 * freshly built grid) and the kinetic energy sum 1/2 m v^2 (the CPU app shows it on screen, cpu/src/view.cpp:66-68).  Any
 * output pointer may be NULL.  Rebuilds the grid and the neighbour lists; positions and velocities are not touched. */
int ps_fluid_stats(PsCtx *ctx, double *mean_density_error, double *max_density_error, double *kinetic_energy);

/* ---- checkpoints (the reference has no persistence: scenes exist only as code, particleapp.cpp:141-215) ----
 * Everything a run needs to continue bit-identically: parameters, particle arrays, constraint lists in insertion order,
 * rigid bodies, viscosity coefficients, position of the wall-jitter stream. */
int ps_save(PsCtx *ctx, const char *path);
int ps_load(const char *path, int device, PsCtx **out);

/* ---- parts of the unified solver that the reference's GPU code does not contain (its rigid_body_functor is an empty
 * stub, solver_kernel.cuh:289-312; XSPH / vorticity exist nowhere in it — SURVEY §0).  Off unless asked for; every
 * parity run of the reference's scenes is unaffected.  Parity unpinned (no reference implementation). ---- */
/* 3-D shape matching (Macklin et al. 2014 §5.1): a rigid body over existing particles whose rest shape is their current
 * configuration; projected once per solver iteration after the distance constraints (one warp per body: shuffle-reduced
 * moment matrix, polar decomposition by the quaternion iteration of Mueller et al. 2016).  stiffness in (0, 1]. */
int ps_add_rigid_body(PsCtx *ctx, const uint32_t *indices, uint64_t n, float stiffness, uint32_t *body);
uint64_t ps_num_rigid_bodies(PsCtx *ctx);
int ps_solve_shapes(PsCtx *ctx);                                         /* the stage alone */
int ps_rigid_body_rotation(PsCtx *ctx, uint32_t body, float *quat_xyzw); /* rotation found by the last projection */
/* SDF contacts between rigid bodies — the reference CPU app's RigidContactConstraint (cpu/src/constraint/rigidcontactconstraint.cpp:
 * 13-96, 2-D) lifted to 3-D inside the contact pass; the reference's GPU solver has nothing like it.  sdf4: per member of the body,
 * in ps_add_rigid_body's order, (gx, gy, gz, depth): the outward surface normal nearest to the particle, in the frame the body was
 * added in (it is turned by the body's current rotation before every contact pass), and the particle's depth below the surface
 * (SDFData, cpu/src/solver/particle.h:82-93; the CPU app's boxes: depth = radius on faces, radius * sqrt(2) at corners).  depth < 0:
 * no data for that member.  A contact of two particles that BOTH carry SDF data takes normal and depth from the shallower one
 * (ties: the lower particle index); for particles of the outermost layers (depth < diameter + EPS) the depth is the particles'
 * overlap and the normal is the direction to the partner, mirrored at the SDF normal when the partner lies behind the surface
 * (Macklin et al. 2014 eq. 13-14; the reference's 2-D code measures that direction the other way round, which in 3-D makes resting
 * edge contacts push sideways — see sdf_contact in csrc/ps_neighbor_kernels.cu); friction acts about that normal.
 * All other contacts are unchanged.  Parity unpinned in 3-D. */
int ps_set_rigid_body_sdf(PsCtx *ctx, uint32_t body, const float *sdf4);
/* XSPH viscosity (v_i += c sum_j (v_j - v_i) W_ij) and vorticity confinement (Macklin & Mueller 2013, eqs. 15-17) as a
 * velocity post-pass of ps_step on the PBF neighbour lists; both coefficients 0 (the default) = off. */
int ps_set_viscosity(PsCtx *ctx, float xsph_c, float vorticity_eps);
int ps_find_neighbors(PsCtx *ctx);            /* K6 alone on the current grid: lambda, neighbour counts, neighbour lists */
int ps_apply_viscosity(PsCtx *ctx, float dt); /* the post-pass alone (after ps_build_grid + ps_find_neighbors) */
int ps_set_ghost_count(PsCtx *ctx, uint64_t ghosts);

/* ---- the slab-decomposed step over NCCL, behind the C ABI (csrc/ps_comm.cu) ----
 * One context per GPU, one process per context.  Rank 0 obtains a 128-byte id (ncclGetUniqueId) and hands it to the other ranks
 * by whatever channel the host has (a file, MPI, torch.distributed ...); every rank calls ps_comm_init, describes its slab and
 * then steps with ps_comm_step instead of ps_step: predict, migration, and per solver iteration the halo refresh, the solver
 * stages and the ghost-lambda exchange, with the neighbour exchange as NCCL send / recv on the context's stream.  NCCL is loaded
 * at run time (dlopen of libnccl.so.2): a single-GPU host never needs it.  Scope: fluid scenes (BASELINE config C5).  Index-based
 * constraints and rigid bodies are refused on slab contexts, and contact FRICTION across a face would need the neighbour's previous
 * position, which the 32-byte halo record does not carry (positions, inverse mass, rest density and phase travel: enough for the
 * density constraint and for frictionless contacts). */
#define PS_COMM_ID_BYTES 128
int ps_comm_get_unique_id(void *id128);
int ps_comm_init(PsCtx *ctx, const void *id128, int rank, int nranks);
int ps_comm_destroy(PsCtx *ctx);
/* this rank owns x in [x_lo, x_hi) (first / last rank: -INFINITY / INFINITY); drift bounds the motion inside one step's solver
 * iterations (0.25 for the reference's constants); exchange_lambda != 0: ghost lambdas come from their owners (halo H + drift),
 * else they are computed locally (halo 2H + 2 drift); capacities in records of the halo / migrant buffers.  Collective: every rank
 * calls it (it also takes the global phase census: a run in which no rank was ever handed a contact-phase particle skips the contact
 * pass); call it again on every rank after appending particles to any of them. */
int ps_comm_set_slab(PsCtx *ctx, float x_lo, float x_hi, float drift, int exchange_lambda, uint64_t halo_capacity, uint64_t migrant_capacity);
/* load balancing: all the cut planes (nranks + 1 floats; the first / last are taken as -INFINITY / INFINITY) and a schedule — every
 * `every` steps (0 = never) the planes move to the equal-count quantiles of the particles' x (per-rank histograms over `bins` <= 65536
 * bins of [x_min, x_max), all-reduced); the migration of that step hands the particles over.  Same arguments on every rank. */
int ps_comm_set_recut(PsCtx *ctx, const float *cuts, uint32_t every, float x_min, float x_max, uint32_t bins);
int ps_comm_get_cuts(PsCtx *ctx, float *cuts, uint32_t *recuts);
int ps_comm_step(PsCtx *ctx, float dt);
/* out[4]: particles handed to neighbours so far, ghosts held in the last iteration, payload bytes sent, steps */
int ps_comm_stats(PsCtx *ctx, uint64_t out[4]);
/* element-wise sum of up to 8 doubles over the ranks (global counts / energies of a decomposed run); blocking */
int ps_comm_allreduce_sum(PsCtx *ctx, double *values, uint32_t count);
uint64_t ps_num_owned(PsCtx *ctx);
#ifdef __cplusplus
}
#endif
#endif /* PSOLVER_H */
