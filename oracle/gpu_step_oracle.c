/*
 * oracle/gpu_step_oracle.c — CPU restatement (plain C, float32) of ONE solver step of the reference's
 * 3-D GPU path, ebirenbaum/ParticleSolver gpu/src/particlesystem.cpp:144-246 and the kernels it launches.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA product path and the "port"
 * CPU baseline in bench.py.  Nothing under particlesolver_b200/ may import, link or execute it.
 *
 * Parity pin: checked against golden vectors dumped from the reference's own unmodified GPU sources run
 * on a B200 (oracle/_ref/ref_gpu, built by oracle/Makefile; fixtures + generator under tests/golden/).
 * Integer outputs (hash, sorted index, cellStart, cellEnd) must match those dumps bit for bit; float
 * outputs match within the tolerance stated in tests/test_oracle_golden.py (the reference is compiled
 * with -use_fast_math, this file with IEEE arithmetic and -ffp-contract=off).
 *
 * Every function cites the reference lines it follows.  The traversal orders (cell scan z,y,x; sorted
 * slot ascending; 500-neighbour cap; constraint order) are kept literally, because they decide which
 * neighbours survive the cap and the floating-point summation order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t u32;

/* phase codes: gpu/src/cuda/shared_variables.cuh:4-9 */
#define PH_FLUID 0
#define PH_CLOTH 2
#define PH_SOLID 3

/* constants: gpu/src/cuda/integration_kernel.cuh:20-40 */
#define EPS 0.001f
#define MAX_FLUID_NEIGHBORS 500
#define H_ 2.f
#define H2 4.f
#define H6 64.f
#define POLY6_COEFF 0.00305992474f
#define SPIKEY_COEFF 0.22381163872f
#define FLUID_RELAXATION .01f
#define K_P .1f
#define E_P 4.f
#define DQ_P .2f
#define S_FRICTION .005f
#define K_FRICTION .0002f

/* mirrors SimParams (gpu/src/cuda/kernel.cuh:9-22) + the scene bounds ParticleSystem keeps
 * (gpu/src/particlesystem.h:113-114) */
typedef struct {
    float gravity[3];
    float radius;
    u32 grid[3];
    float origin[3];
    float cell[3];
    int min_b[3];
    int max_b[3];
} OrParams;

/* SOR factor on the Jacobi-averaged deltas ("Jacobi averaging with SOR", BASELINE north_star).  The reference's live GPU path has
 * none (its only omega is the dead dense solver's, gpu/src/cuda/solver_kernel.cuh:244,278, and the CPU app's relaxation constant,
 * cpu/src/solver/solver.h:8): omega = 1 reproduces it exactly.  For omega != 1 every averaged delta — contacts (dp / numNeighbors,
 * integration_kernel.cuh:436-437), PBF delta-p (/ (rho0 + numNeighbors), :640) and distance constraints (/ occurrences,
 * solver_kernel.cuh:64-76) — is multiplied by omega, exactly where the CUDA path applies PsParams.omega. */
static float g_omega = 1.0f;
void or_set_omega(float omega) { g_omega = omega; }

/* ---- K1: integrateSystem + copyToXstar (integration.cu:122-135, integration_kernel.cuh:159-184,
 *          shared_variables.cu:52-57).  vel is NOT written back; inverse mass is not consulted. ---- */
void or_predict(float *pos, const float *vel, float *prev, u32 n, float dt, const float *g) {
    memcpy(prev, pos, (size_t)n * 16);
    #pragma omp parallel for schedule(static)
    for (u32 i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            float v = vel[4 * i + c] + g[c] * dt;
            pos[4 * i + c] = pos[4 * i + c] + v * dt;
        }
    }
}

/* ---- calcGridPos / calcGridHash (integration_kernel.cuh:187-203) ---- */
static inline void grid_pos(const OrParams *p, const float *x, int *gp) {
    for (int c = 0; c < 3; c++) gp[c] = (int)floorf((x[c] - p->origin[c]) / p->cell[c]);
}
static inline u32 grid_hash(const OrParams *p, int gx, int gy, int gz) {
    u32 x = (u32)gx & (p->grid[0] - 1), y = (u32)gy & (p->grid[1] - 1), z = (u32)gz & (p->grid[2] - 1);
    /* __umul24 == 32-bit multiply while operands stay < 2^24 */
    return (z * p->grid[1]) * p->grid[0] + y * p->grid[0] + x;
}

/* ---- K2: calcHashD (integration_kernel.cuh:206-225) ---- */
void or_calc_hash(const float *pos, u32 n, const OrParams *p, u32 *hash, u32 *index) {
    #pragma omp parallel for schedule(static)
    for (u32 i = 0; i < n; i++) {
        int gp[3];
        grid_pos(p, pos + 4 * (size_t)i, gp);
        hash[i] = grid_hash(p, gp[0], gp[1], gp[2]);
        index[i] = i;
    }
}

/* ---- K3: sortParticles = thrust::sort_by_key on u32 keys (integration.cu:270-275): radix sort, hence
 *          stable: ties keep ascending original index.  LSD byte radix here. ---- */
void or_sort(u32 *hash, u32 *index, u32 n) {
    u32 *h2 = (u32 *)malloc((size_t)n * 4), *i2 = (u32 *)malloc((size_t)n * 4);
    u32 *hs = hash, *is = index, *hd = h2, *id = i2;
    for (int pass = 0; pass < 4; pass++) {
        size_t cnt[257];
        memset(cnt, 0, sizeof cnt);
        int sh = 8 * pass;
        for (u32 i = 0; i < n; i++) cnt[((hs[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (u32 i = 0; i < n; i++) {
            size_t dst = cnt[(hs[i] >> sh) & 255]++;
            hd[dst] = hs[i];
            id[dst] = is[i];
        }
        u32 *t = hs; hs = hd; hd = t;
        t = is; is = id; id = t;
    }
    /* 4 passes: data is back in the caller's arrays */
    free(h2);
    free(i2);
}

/* ---- K4: reorderDataAndFindCellStart (integration.cu:184-268, integration_kernel.cuh:229-299).
 *          cellStart is memset to 0xffffffff; cellEnd is NOT cleared (stale for empty cells). ---- */
void or_reorder(const u32 *hash, const u32 *index, const float *pos, const float *w, const int *phase, u32 n, u32 num_cells,
                u32 *cell_start, u32 *cell_end, float *spos, float *sw, int *sphase) {
    memset(cell_start, 0xff, (size_t)num_cells * 4);
    for (u32 i = 0; i < n; i++) {
        u32 h = hash[i];
        if (i == 0 || h != hash[i - 1]) {
            cell_start[h] = i;
            if (i > 0) cell_end[hash[i - 1]] = i;
        }
        if (i == n - 1) cell_end[h] = i + 1;
    }
    #pragma omp parallel for schedule(static)
    for (u32 i = 0; i < n; i++) {
        u32 s = index[i];
        memcpy(spos + 4 * (size_t)i, pos + 4 * (size_t)s, 16);
        sw[i] = w[s];
        sphase[i] = phase[s];
    }
}

static inline float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* dot(r, r) as the reference's own build computes it where it DECIDES the neighbour sets (collideCell :330-336, collideCellRadius
 * :505-508): nvcc -use_fast_math contracts helper_math's x*x + y*y + z*z into FMUL y*y, FFMA x*x + ., FFMA z*z + . (cuobjdump -sass
 * of oracle/_ref/ref_gpu, findLambdasD / collideD).  A pair within one ulp of the radius is in or out by that rounding, so the
 * neighbour COUNTS — an exact contract — need it; everything downstream of the decision keeps plain IEEE arithmetic. */
static inline float dot3_self_nvcc(const float *r) { return fmaf(r[2], r[2], fmaf(r[0], r[0], r[1] * r[1])); }

/* ---- K5: collide / collideD / collideCell (integration.cu:338-386, integration_kernel.cuh:303-462).
 *          Runs for sorted-phase >= CLOTH only; others leave pos[] and num_neighbors[] untouched. ---- */
/* adj_off / adj: NULL = the reference (particles of one phase > SOLID never collide with each other).  Otherwise the opt-in
 * self-collision of PS_FLAG_SELF_COLLISION (NOT in the reference): CSR adjacency of the distance constraints by
 * original index; a same-phase pair is skipped only when either particle has no distance constraint (a shape-matched body)
 * or the two are joined by one. */
static int same_body_skip(const u32 *adj_off, const u32 *adj, u32 oi, u32 oj) {
    if (!adj_off) return 1;
    if (adj_off[oi] == adj_off[oi + 1] || adj_off[oj] == adj_off[oj + 1]) return 1;
    for (u32 k = adj_off[oi]; k < adj_off[oi + 1]; k++)
        if (adj[k] == oj) return 1;
    return 0;
}

/* SDF contact of two rigid-body particles: RigidContactConstraint of the reference CPU app (cpu/src/constraint/
 * rigidcontactconstraint.cpp:13-66, 2-D) lifted to 3-D, as the gather form of K5 sees it (particle 1 = the particle being updated).
 * si / sj = (outward unit gradient in world frame, depth); r = x_i - x_j.  NOT in the reference's GPU solver; unpinned. */
static int sdf_contact(const float *si, const float *sj, int i_first, const float *r, float dist, float diam, float *d, float *e) {
    int mine = si[3] < sj[3] || (si[3] == sj[3] && i_first);
    if (mine) { *d = si[3]; e[0] = si[0]; e[1] = si[1]; e[2] = si[2]; }
    else { *d = sj[3]; e[0] = -sj[0]; e[1] = -sj[1]; e[2] = -sj[2]; }
    if (*d < diam + EPS) {               /* initBoundary (:13-27) */
        *d = diam - dist;
        if (*d < EPS) return 0;
        /* direction from i to j (the reference's 2-D code takes p1 - p2, the opposite one: see sdf_contact in ps_neighbor_kernels.cu) */
        float x[3] = {0.f, i_first ? 1.f : -1.f, 0.f};
        if (dist > EPS) { x[0] = -r[0] / dist; x[1] = -r[1] / dist; x[2] = -r[2] / dist; }
        float dp = x[0] * e[0] + x[1] * e[1] + x[2] * e[2];
        if (dp < 0.f) { for (int c = 0; c < 3; c++) e[c] = x[c] - 2.f * dp * e[c]; }
        else { for (int c = 0; c < 3; c++) e[c] = x[c]; }
    }
    return 1;
}

/* sdf_world: NULL, or per ORIGINAL particle index (gx, gy, gz, depth), depth < 0 or NaN = none (ps_set_rigid_body_sdf) */
void or_collide_ext(float *pos, const float *prev, const float *spos, const float *sw, const int *sphase, const u32 *index,
                    const u32 *cell_start, const u32 *cell_end, u32 n, const OrParams *p, u32 *num_neighbors, const u32 *adj_off,
                    const u32 *adj, const float *sdf_world) {
    const float collideDist = p->radius * 2.001f;
    const float collideDist2 = collideDist * collideDist;
    #pragma omp parallel
    {
        u32 *nb = (u32 *)malloc(MAX_FLUID_NEIGHBORS * sizeof(u32));
        #pragma omp for schedule(dynamic, 256)
        for (u32 i = 0; i < n; i++) {
            int phase = sphase[i];
            if (phase < PH_CLOTH) continue;
            const float *x = spos + 4 * (size_t)i;
            int gp[3];
            grid_pos(p, x, gp);
            u32 nn = 0;
            for (int z = -1; z <= 1; z++)
                for (int y = -1; y <= 1; y++)
                    for (int xx = -1; xx <= 1; xx++) {
                        u32 h = grid_hash(p, gp[0] + xx, gp[1] + y, gp[2] + z);
                        u32 s = cell_start[h];
                        if (s == 0xffffffffu) continue;
                        u32 e = cell_end[h];
                        for (u32 j = s; j < e; j++) {
                            if (j == i) continue;
                            int phase2 = sphase[j];
                            if (phase > PH_SOLID && phase == phase2 && same_body_skip(adj_off, adj, index[i], index[j])) continue;
                            float d[3] = {x[0] - spos[4 * (size_t)j], x[1] - spos[4 * (size_t)j + 1], x[2] - spos[4 * (size_t)j + 2]};
                            float mag2 = dot3_self_nvcc(d);
                            if (mag2 < collideDist2 && nn < MAX_FLUID_NEIGHBORS) nb[nn++] = j;
                        }
                    }
            num_neighbors[i] = nn;
            float w = sw[i];
            float sW = (w != 0.f ? (1.f / ((1.f / w) * expf(-x[1]))) : w);
            u32 orig = index[i];
            const float *pp = prev + 4 * (size_t)orig;
            float delta[3] = {0.f, 0.f, 0.f};
            float fn = (g_omega == 1.0f) ? (float)nn : (float)nn / g_omega;  /* omega scales the averaged contact delta */
            for (u32 k = 0; k < nn; k++) {
                u32 j = nb[k];
                const float *x2 = spos + 4 * (size_t)j;
                float w2 = sw[j];
                int phase2 = sphase[j];
                float d[3] = {x[0] - x2[0], x[1] - x2[1], x[2] - x2[2]};
                float dist = sqrtf(dot3(d, d));
                float mag = dist - collideDist;
                float colW = w, colW2 = w2;
                if (phase >= PH_SOLID && phase2 >= PH_SOLID) {
                    colW = sW;
                    colW2 = (w2 != 0.f ? (1.f / ((1.f / w2) * expf(-x[1]))) : w2);
                }
                float scale = mag / (colW + colW2);
                float sd = scale / dist;
                float dp[3] = {d[0] * sd, d[1] * sd, d[2] * sd};
                float fnv[3] = {d[0], d[1], d[2]}, fd = dist; /* friction: normal before normalisation, length scale of the cone */
                if (sdf_world && phase >= PH_SOLID && phase2 >= PH_SOLID) {
                    const float *si = sdf_world + 4 * (size_t)orig, *sj = sdf_world + 4 * (size_t)index[j];
                    if (si[3] >= 0.f && sj[3] >= 0.f) {
                        float depth, e[3];
                        if (!sdf_contact(si, sj, orig < index[j], d, dist, 2.f * p->radius, &depth, e)) continue;
                        float s_ = depth / (colW + colW2);
                        for (int c = 0; c < 3; c++) { dp[c] = e[c] * s_; fnv[c] = e[c]; }
                    }
                }
                float dp1[3], dp2[3];
                for (int c = 0; c < 3; c++) {
                    dp1[c] = (-colW * dp[c]) / fn;
                    dp2[c] = (colW2 * dp[c]) / fn;
                    delta[c] += dp1[c];
                }
                if (phase < PH_SOLID || phase2 < PH_SOLID) continue;
                const float *pp2 = prev + 4 * (size_t)index[j];
                float inv = 1.0f / sqrtf(dot3(fnv, fnv)); /* normalize(): v * rsqrtf(dot(v,v)) */
                float nf[3] = {fnv[0] * inv, fnv[1] * inv, fnv[2] * inv};
                float rel[3];
                /* [sic] second term starts from prevPos of i, not pos2 (integration_kernel.cuh:447) */
                for (int c = 0; c < 3; c++) rel[c] = (x[c] + dp1[c] - pp[c]) - (pp[c] + dp2[c] - pp2[c]);
                float dn = dot3(rel, nf);
                float dpt[3] = {rel[0] - dn * nf[0], rel[1] - dn * nf[1], rel[2] - dn * nf[2]};
                float ldpt = sqrtf(dot3(dpt, dpt));
                if (ldpt < EPS) continue;
                if (ldpt < S_FRICTION * fd) {
                    for (int c = 0; c < 3; c++) delta[c] -= (dpt[c] * colW) / (colW + colW2);
                } else {
                    float m = fminf(K_FRICTION * fd / ldpt, 1.f);
                    for (int c = 0; c < 3; c++) delta[c] -= dpt[c] * m;
                }
            }
            float *o = pos + 4 * (size_t)orig;
            o[0] = x[0] + delta[0];
            o[1] = x[1] + delta[1];
            o[2] = x[2] + delta[2];
            o[3] = 1.0f;
        }
        free(nb);
    }
}

void or_collide(float *pos, const float *prev, const float *spos, const float *sw, const int *sphase, const u32 *index,
                const u32 *cell_start, const u32 *cell_end, u32 n, const OrParams *p, u32 *num_neighbors) {
    or_collide_ext(pos, prev, spos, sw, sphase, index, cell_start, cell_end, n, p, num_neighbors, NULL, NULL, NULL);
}

/* neighbour gather of findLambdasD / collideCellRadius (integration_kernel.cuh:482-559) */
static u32 fluid_neighbors(const OrParams *p, const float *spos, const u32 *cell_start, const u32 *cell_end, u32 i, u32 *nb) {
    const float *x = spos + 4 * (size_t)i;
    int gp[3];
    grid_pos(p, x, gp);
    int rad = (int)ceilf(H_ / p->cell[0]);
    u32 nn = 0;
    for (int z = -rad; z <= rad; z++)
        for (int y = -rad; y <= rad; y++)
            for (int xx = -rad; xx <= rad; xx++) {
                u32 h = grid_hash(p, gp[0] + xx, gp[1] + y, gp[2] + z);
                u32 s = cell_start[h];
                if (s == 0xffffffffu) continue;
                u32 e = cell_end[h];
                for (u32 j = s; j < e; j++) {
                    if (j == i) continue;
                    const float *x2 = spos + 4 * (size_t)j;
                    float r[3] = {x[0] - x2[0], x[1] - x2[1], x[2] - x2[2]};
                    float d2 = dot3_self_nvcc(r);
                    if (d2 < H2 && nn < MAX_FLUID_NEIGHBORS) nb[nn++] = j;
                }
            }
    return nn;
}

/* ---- K6 + K7: solveFluids = findLambdasD then solveFluidsD (integration.cu:453-508,
 *      integration_kernel.cuh:521-642).  lambda[] and num_neighbors[] are persistent, indexed by SORTED slot
 *      and only written at fluid slots (non-fluid neighbours contribute whatever the slot last held). ---- */
/* stages: 1 = lambda (findLambdasD), 2 = delta p (solveFluidsD), 3 = both.  The two halves can be called separately — a
 * slab-decomposed run exchanges ghost lambdas in between (tests/slab_oracle_engine.py); the second half then gathers the
 * neighbour lists again, which yields the lists the first half saw (positions in spos do not change in between). */
void or_solve_fluids_stages(const float *spos, const float *sw, const int *sphase, const u32 *index, const u32 *cell_start,
                            const u32 *cell_end, float *pos, u32 n, const OrParams *p, const float *ros, float *lambda,
                            u32 *num_neighbors, int stages) {
    u32 **lists = (u32 **)calloc(n, sizeof(u32 *));
    #pragma omp parallel
    {
        u32 *nb = (u32 *)malloc(MAX_FLUID_NEIGHBORS * sizeof(u32));
        #pragma omp for schedule(dynamic, 64)
        for (u32 i = 0; i < n; i++) {
            if (sphase[i] != PH_FLUID) continue;
            u32 nn = fluid_neighbors(p, spos, cell_start, cell_end, i, nb);
            num_neighbors[i] = nn;
            lists[i] = (u32 *)malloc((nn ? nn : 1) * sizeof(u32));
            memcpy(lists[i], nb, nn * sizeof(u32));
            if (!(stages & 1)) continue;
            const float *x = spos + 4 * (size_t)i;
            float w = sw[i];
            float ro0 = ros[index[i]];
            float ro = 0.f, denom = 0.f, grad[3] = {0.f, 0.f, 0.f};
            for (u32 k = 0; k < nn; k++) {
                const float *x2 = spos + 4 * (size_t)nb[k];
                float r[3] = {x[0] - x2[0], x[1] - x2[1], x[2] - x2[2]};
                float rlen2 = dot3(r, r);
                float rlen = sqrtf(rlen2);
                float hMinus2 = H2 - rlen2, hMinus = H_ - rlen;
                ro += (POLY6_COEFF * hMinus2 * hMinus2 * hMinus2) / w;
                float sg[3];
                if (rlen < 0.0001f) {
                    sg[0] = sg[1] = sg[2] = 0.f;
                } else {
                    for (int c = 0; c < 3; c++) sg[c] = (r[c] / rlen) * -SPIKEY_COEFF * hMinus * hMinus;
                }
                for (int c = 0; c < 3; c++) {
                    sg[c] /= ro0;
                    grad[c] += -sg[c];
                }
                denom += dot3(sg, sg);
            }
            ro += (POLY6_COEFF * H6) / w;
            denom += dot3(grad, grad);
            lambda[i] = -((ro / ro0) - 1) / (denom + FLUID_RELAXATION);
        }
        /* implicit barrier: all lambdas written before any Δp reads them (two kernel launches) */
        #pragma omp for schedule(dynamic, 64)
        for (u32 i = 0; i < n; i++) {
            if (sphase[i] != PH_FLUID || !(stages & 2)) continue;
            u32 nn = num_neighbors[i];
            const u32 *l = lists[i];
            const float *x = spos + 4 * (size_t)i;
            float delta[4] = {0.f, 0.f, 0.f, 0.f};
            for (u32 k = 0; k < nn; k++) {
                const float *x2 = spos + 4 * (size_t)l[k];
                float r[4] = {x[0] - x2[0], x[1] - x2[1], x[2] - x2[2], x[3] - x2[3]};
                float rlen2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
                float rlen = sqrtf(rlen2);
                float hMinus2 = H2 - rlen2, hMinus = H_ - rlen;
                float sg[4];
                if (rlen < 0.0001f) {
                    float e[4] = {0.f, EPS, 0.f, 0.f};
                    for (int c = 0; c < 4; c++) sg[c] = e[c] * -SPIKEY_COEFF * hMinus * hMinus;
                } else {
                    for (int c = 0; c < 4; c++) sg[c] = (r[c] / rlen) * -SPIKEY_COEFF * hMinus * hMinus;
                }
                float term2 = H2 - (DQ_P * DQ_P * H2);
                float numer = (POLY6_COEFF * hMinus2 * hMinus2 * hMinus2);
                float den = (POLY6_COEFF * term2 * term2 * term2);
                float lambdaCorr = -K_P * powf(numer / den, E_P);
                float s = lambda[i] + lambda[l[k]] + lambdaCorr;
                for (int c = 0; c < 4; c++) delta[c] += s * sg[c];
            }
            u32 orig = index[i];
            float div = ros[orig] + (float)nn;
            if (g_omega != 1.0f) div /= g_omega;
            for (int c = 0; c < 4; c++) pos[4 * (size_t)orig + c] += delta[c] / div;
        }
        free(nb);
    }
    for (u32 i = 0; i < n; i++) free(lists[i]);
    free(lists);
}
void or_solve_fluids(const float *spos, const float *sw, const int *sphase, const u32 *index, const u32 *cell_start,
                     const u32 *cell_end, float *pos, u32 n, const OrParams *p, const float *ros, float *lambda,
                     u32 *num_neighbors) {
    or_solve_fluids_stages(spos, sw, sphase, index, cell_start, cell_end, pos, n, p, ros, lambda, num_neighbors, 3);
}

/* ---- K8: collideWorld / collide_world_functor (integration.cu:319-336, integration_kernel.cuh:57-157).
 *          rands[6] are the uniforms cuRAND produced for this iteration. phase[] is in ORIGINAL order. ---- */
void or_collide_world(float *pos, const float *prev, const int *phase, u32 n, const float *rands, const OrParams *p) {
    const float rad = p->radius;
    #pragma omp parallel for schedule(static)
    for (u32 i = 0; i < n; i++) {
        float *P = pos + 4 * (size_t)i;
        const float *X = prev + 4 * (size_t)i;
        int ph = phase[i];
        float e[3] = {P[0], P[1], P[2]};
        float nrm[3] = {0.f, 0.f, 0.f};
        float d = rad;
        float eps = d * 0.f;
        if (ph < PH_SOLID) eps = d * 0.01f;
        if (e[1] < p->min_b[1] + rad) { e[1] = p->min_b[1] + rad + rands[5] * eps; nrm[1] += 1.f; }
        eps = d * 0.01f;
        if (e[0] > p->max_b[0] - rad) { e[0] = p->max_b[0] - (rad + rands[0] * eps); nrm[0] += -1.f; }
        if (e[0] < p->min_b[0] + rad) { e[0] = p->min_b[0] + (rad + rands[1] * eps); nrm[0] += 1.f; }
        if (e[1] > p->max_b[1] - rad) { e[1] = p->max_b[1] - (rad + rands[2] * eps); nrm[1] += -1.f; }
        if (e[2] > p->max_b[2] - rad) { e[2] = p->max_b[2] - (rad + rands[3] * eps); nrm[2] += -1.f; }
        if (e[2] < p->min_b[2] + rad) { e[2] = p->min_b[2] + (rad + rands[4] * eps); nrm[2] += 1.f; }
        if (sqrtf(dot3(nrm, nrm)) < EPS || ph < PH_CLOTH) {
            P[0] = e[0]; P[1] = e[1]; P[2] = e[2];
            continue;
        }
        float dp[3] = {e[0] - X[0], e[1] - X[1], e[2] - X[2]};
        float dn = dot3(dp, nrm);
        float dpt[3] = {dp[0] - dn * nrm[0], dp[1] - dn * nrm[1], dp[2] - dn * nrm[2]};
        float ldpt = sqrtf(dot3(dpt, dpt));
        if (ldpt < EPS) {
            P[0] = e[0]; P[1] = e[1]; P[2] = e[2];
            continue;
        }
        if (ldpt < sqrtf(S_FRICTION) * d) {
            for (int c = 0; c < 3; c++) e[c] -= dpt[c];
        } else {
            float m = fminf(sqrtf(K_FRICTION) * d / ldpt, 1.f);
            for (int c = 0; c < 3; c++) e[c] -= dpt[c] * m;
        }
        P[0] = e[0]; P[1] = e[1]; P[2] = e[2];
    }
}

/* ---- K9: solveDistanceConstraints (solver.cu:196-231; solver_kernel.cuh:27-72,155-163).
 *   Per constraint (a,b,d0): delta = 0.5*(d0-|pa-pb|)*(pa-pb)/|pa-pb| to a, -delta to b (0 if |.|<=1e-4).
 *   The reference sort_by_key's [all a's..., all b's...] (stable) and reduce_by_key's, so particle p sums its
 *   "a" deltas in constraint order, then its "b" deltas in constraint order; the sum is divided by
 *   occurences[p] and added to pos.  The reference applies the k-th distinct key's sum to particle k
 *   (solver.cu:224-230), which equals the by-index scatter iff constrained particles are the index prefix
 *   [0,K) — true for every built-in scene; returns 0 if that held here, 1 otherwise (and scatters by index). */
int or_solve_distance(float *pos, const u32 *idx, const float *rest, u32 m, const u32 *occ, u32 n) {
    if (m == 0) return 0;
    float *sum = (float *)calloc((size_t)n * 4, sizeof(float));
    float *dl = (float *)calloc((size_t)m * 3, sizeof(float));
    unsigned char *touched = (unsigned char *)calloc(n, 1);
    for (u32 c = 0; c < m; c++) {
        u32 a = idx[2 * c], b = idx[2 * c + 1];
        const float *pa = pos + 4 * (size_t)a, *pb = pos + 4 * (size_t)b;
        float r[3] = {pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]};
        float dist = sqrtf(dot3(r, r));
        if (dist > 0.0001f) {
            float mag = (rest[c] - dist) * .5f;
            for (int k = 0; k < 3; k++) dl[3 * (size_t)c + k] = (r[k] / dist) * mag;
        }
        touched[a] = touched[b] = 1;
    }
    /* sorted-by-key order: every "a" entry (constraint order) precedes every "b" entry of the same particle */
    for (u32 c = 0; c < m; c++)
        for (int k = 0; k < 3; k++) sum[4 * (size_t)idx[2 * c] + k] += dl[3 * (size_t)c + k];
    for (u32 c = 0; c < m; c++)
        for (int k = 0; k < 3; k++) sum[4 * (size_t)idx[2 * c + 1] + k] += -dl[3 * (size_t)c + k];
    int nonprefix = 0;
    u32 K = 0;
    for (u32 i = 0; i < n; i++) K += touched[i];
    for (u32 i = 0; i < K; i++) if (!touched[i]) nonprefix = 1;
    for (u32 i = 0; i < n; i++) {
        if (!touched[i]) continue;
        float o = (float)occ[i];
        for (int k = 0; k < 3; k++) pos[4 * (size_t)i + k] += g_omega * (sum[4 * (size_t)i + k] / o);
    }
    free(sum); free(dl); free(touched);
    return nonprefix;
}

/* ---- K10: solvePointConstraints (solver.cu:180-194; solver_kernel.cuh:10-25) ---- */
void or_solve_point(float *pos, const u32 *pidx, const float *pxyz, u32 np) {
    for (u32 c = 0; c < np; c++) memcpy(pos + 4 * (size_t)pidx[c], pxyz + 3 * (size_t)c, 12);
}

/* ---- K11: calcVelocity / subtract_functor (integration.cu:416-426; integration_kernel.cuh:465-475):
 *           transform(pos, Xstar) -> (Xstar - pos) / -dt, all four components. ---- */
void or_calc_velocity(const float *pos, const float *prev, float *vel, u32 n, float dt) {
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < (size_t)n * 4; i++) vel[i] = (prev[i] - pos[i]) / -dt;
}

/* ---- per-particle occurrence counts (solver.cu:72-106,122,149): #distance endpoints + #point pins ---- */
void or_occurrences(u32 *occ, u32 n, const u32 *didx, u32 m, const u32 *pidx, u32 np) {
    memset(occ, 0, (size_t)n * 4);
    for (u32 c = 0; c < 2 * m; c++) occ[didx[c]]++;
    for (u32 c = 0; c < np; c++) occ[pidx[c]]++;
}

/* persistent state the reference keeps in file-scope device vectors (integration.cu:23-34) */
typedef struct {
    u32 n, num_cells;
    u32 *hash, *index, *cell_start, *cell_end, *num_neighbors;
    float *spos, *sw, *lambda;
    int *sphase;
} OrState;

OrState *or_state_new(u32 n, u32 num_cells) {
    OrState *s = (OrState *)calloc(1, sizeof(OrState));
    s->n = n; s->num_cells = num_cells;
    s->hash = (u32 *)calloc(n, 4); s->index = (u32 *)calloc(n, 4);
    s->cell_start = (u32 *)calloc(num_cells, 4); s->cell_end = (u32 *)calloc(num_cells, 4);
    s->num_neighbors = (u32 *)calloc(n, 4);
    s->spos = (float *)calloc((size_t)n * 4, 4); s->sw = (float *)calloc(n, 4); s->lambda = (float *)calloc(n, 4);
    s->sphase = (int *)calloc(n, 4);
    return s;
}
void or_state_free(OrState *s) {
    free(s->hash); free(s->index); free(s->cell_start); free(s->cell_end); free(s->num_neighbors);
    free(s->spos); free(s->sw); free(s->lambda); free(s->sphase); free(s);
}
u32 *or_state_hash(OrState *s) { return s->hash; }
u32 *or_state_index(OrState *s) { return s->index; }
u32 *or_state_cell_start(OrState *s) { return s->cell_start; }
u32 *or_state_cell_end(OrState *s) { return s->cell_end; }
u32 *or_state_num_neighbors(OrState *s) { return s->num_neighbors; }
float *or_state_lambda(OrState *s) { return s->lambda; }
float *or_state_sorted_pos(OrState *s) { return s->spos; }

/* one grid build: K2, K3, K4 */
void or_build_grid(OrState *s, const float *pos, const float *w, const int *phase, const OrParams *p) {
    or_calc_hash(pos, s->n, p, s->hash, s->index);
    or_sort(s->hash, s->index, s->n);
    or_reorder(s->hash, s->index, pos, w, phase, s->n, s->num_cells, s->cell_start, s->cell_end, s->spos, s->sw, s->sphase);
}

/* ---- a0: ParticleSystem::update (particlesystem.cpp:144-246).  rands = iters x 6 uniforms. ---- */
void or_step(OrState *s, float *pos, float *vel, float *prev, const float *w, const int *phase, const float *ros,
             const u32 *didx, const float *drest, u32 m, const u32 *pidx, const float *pxyz, u32 np, const u32 *occ,
             const OrParams *p, float dt, int iters, const float *rands) {
    dt = fminf(dt, .05f);
    u32 n = s->n;
    or_predict(pos, vel, prev, n, dt, p->gravity);
    for (int it = 0; it < iters; it++) {
        or_build_grid(s, pos, w, phase, p);
        or_collide(pos, prev, s->spos, s->sw, s->sphase, s->index, s->cell_start, s->cell_end, n, p, s->num_neighbors);
        or_solve_fluids(s->spos, s->sw, s->sphase, s->index, s->cell_start, s->cell_end, pos, n, p, ros, s->lambda,
                        s->num_neighbors);
        or_collide_world(pos, prev, phase, n, rands + 6 * it, p);
        or_solve_distance(pos, didx, drest, m, occ, n);
        or_solve_point(pos, pidx, pxyz, np);
    }
    or_calc_velocity(pos, prev, vel, n, dt);
}

/* ---- statistics used by the long-run parity tests (not a reference function: the reference only shows
 *      kinetic energy on screen in the CPU app, cpu/src/simulation.cpp:1293-1303).  Density uses the K6
 *      estimator (integration_kernel.cuh:565-589) on a freshly built grid. out = {mean|rho/rho0-1|, max, KE}. */
void or_fluid_stats(OrState *s, const float *pos, const float *vel, const float *w, const int *phase, const float *ros,
                    const OrParams *p, double *out) {
    or_build_grid(s, pos, w, phase, p);
    u32 n = s->n;
    double sum = 0, mx = 0, ke = 0;
    u32 cnt = 0;
    #pragma omp parallel
    {
        u32 *nb = (u32 *)malloc(MAX_FLUID_NEIGHBORS * sizeof(u32));
        #pragma omp for schedule(dynamic, 64) reduction(+ : sum, cnt) reduction(max : mx)
        for (u32 i = 0; i < n; i++) {
            if (s->sphase[i] != PH_FLUID) continue;
            u32 nn = fluid_neighbors(p, s->spos, s->cell_start, s->cell_end, i, nb);
            const float *x = s->spos + 4 * (size_t)i;
            float wi = s->sw[i], ro = 0.f;
            for (u32 k = 0; k < nn; k++) {
                const float *x2 = s->spos + 4 * (size_t)nb[k];
                float r[3] = {x[0] - x2[0], x[1] - x2[1], x[2] - x2[2]};
                float h2 = H2 - dot3(r, r);
                ro += (POLY6_COEFF * h2 * h2 * h2) / wi;
            }
            ro += (POLY6_COEFF * H6) / wi;
            double err = fabs((double)ro / ros[s->index[i]] - 1.0);
            sum += err;
            cnt++;
            if (err > mx) mx = err;
        }
        free(nb);
    }
    for (u32 i = 0; i < n; i++) {
        double m = w[i] != 0.f ? 1.0 / w[i] : 0.0;
        ke += 0.5 * m * ((double)vel[4 * (size_t)i] * vel[4 * (size_t)i] + (double)vel[4 * (size_t)i + 1] * vel[4 * (size_t)i + 1] +
                         (double)vel[4 * (size_t)i + 2] * vel[4 * (size_t)i + 2]);
    }
    out[0] = cnt ? sum / cnt : 0.0;
    out[1] = mx;
    out[2] = ke;
}
