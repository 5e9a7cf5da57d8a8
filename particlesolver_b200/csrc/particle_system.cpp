// particlesolver_b200/csrc/particle_system.cpp — headless host class mirroring the reference's ParticleSystem
// (gpu/src/particlesystem.{h,cpp}) on top of the libpsolver C ABI.  Scene builders reproduce the reference's
// particle counts, positions, constraint lists and glibc rand() consumption order exactly,
// because the parity tests compare against scenes built by the reference's own builders.
#include "../../include/particle_system.h"
#include "../../include/ps_scenes.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace psb200 {

const float3 colors[numColors] = {{.722f, .141f, .447f}, {.886f, .455f, .173f}, {.110f, .569f, .478f}, {.588f, .824f, .161f},
                                  {.722f, .267f, .506f}, {.831f, .345f, .031f}, {.020f, .533f, .431f}, {.506f, .773f, .027f}};

// reference particlesystem.cpp:78-81
static inline float frand() { return rand() / (float)RAND_MAX; }

ParticleSystem::ParticleSystem(float particleRadius, uint3 gridSize, uint maxParticles, int3 minBounds, int3 maxBounds, int iterations)
    : m_ctx(nullptr), m_particleRadius(particleRadius), m_maxParticles(maxParticles), m_numParticles(0), m_gridSize(gridSize),
      m_rigidIndex(0), m_minBounds(minBounds), m_maxBounds(maxBounds), m_solverIterations((uint)iterations) {
    // same parameter block the reference's constructor fills (particlesystem.cpp:50-66)
    PsParams p;
    ps_default_params(&p);
    p.particle_radius = particleRadius;
    p.grid_size[0] = gridSize.x; p.grid_size[1] = gridSize.y; p.grid_size[2] = gridSize.z;
    p.world_origin[0] = p.world_origin[1] = p.world_origin[2] = 0.f;
    p.cell_size[0] = p.cell_size[1] = p.cell_size[2] = particleRadius * 2.0f;
    p.gravity[0] = 0.f; p.gravity[1] = -9.8f; p.gravity[2] = 0.f;
    p.global_damping = 1.0f;
    p.min_bounds[0] = minBounds.x; p.min_bounds[1] = minBounds.y; p.min_bounds[2] = minBounds.z;
    p.max_bounds[0] = maxBounds.x; p.max_bounds[1] = maxBounds.y; p.max_bounds[2] = maxBounds.z;
    p.solver_iterations = (uint)iterations;
    int dev = 0;
    if (const char *e = getenv("PS_DEVICE")) dev = atoi(e);
    note(ps_create(dev, &p, maxParticles, &m_ctx), "ps_create");
}

ParticleSystem::~ParticleSystem() { ps_destroy(m_ctx); }

void ParticleSystem::note(int rc, const char *where) {
    if (rc != PS_OK && m_error.empty()) m_error = std::string(where) + ": " + ps_last_error();
}

// reference particlesystem.cpp:144-246
void ParticleSystem::update(float deltaTime) {
    if (!m_ctx) return;
    deltaTime = std::min(deltaTime, .05f);
    if (m_numParticles != 0) note(ps_step(m_ctx, deltaTime), "ps_step");
    addNewStuff();
}

void ParticleSystem::addNewStuff() { addParticles(); addFluids(); }

// queued "shot" particles become SOLID particles of rest density 1.5 (reference :256-271)
void ParticleSystem::addParticles() {
    while (m_particlesToAdd.size() >= 2) {
        float4 pos = m_particlesToAdd.front(); m_particlesToAdd.pop_front();
        float4 vel = m_particlesToAdd.front(); m_particlesToAdd.pop_front();
        addParticle(pos, make_float4(vel.x, vel.y, vel.z, 0), vel.w, 1.5f, PS_PHASE_SOLID);
    }
}

// queued emitter particles become FLUID particles falling at -1 (reference :273-293)
void ParticleSystem::addFluids() {
    if (m_fluidsToAdd.empty()) return;
    const uint start = m_numParticles;
    float4 color = make_float4(0, 0, 0, 0);
    while (m_fluidsToAdd.size() >= 2) {
        float4 pos = m_fluidsToAdd.front(); m_fluidsToAdd.pop_front();
        color = m_fluidsToAdd.front(); m_fluidsToAdd.pop_front();
        addParticle(make_float4(pos.x, pos.y, pos.z, 1), make_float4(0, -1, 0, 0), pos.w, color.w, PS_PHASE_FLUID);
    }
    m_colorIndex.push_back(make_int2(start, m_numParticles));
    m_colors.push_back(make_float4(color.x, color.y, color.z, 1.f));
}

void ParticleSystem::setParticleToAdd(float3 pos, float3 vel, float mass) {
    const float jitter = m_particleRadius * 0.01f;
    pos.x += (frand() * 2.0f - 1.0f) * jitter;
    pos.y += (frand() * 2.0f - 1.0f) * jitter;
    m_particlesToAdd.push_back(make_float4(pos.x, pos.y, pos.z, 1.f));
    m_particlesToAdd.push_back(make_float4(vel.x, vel.y, vel.z, mass));
    m_colorIndex.push_back(make_int2(m_numParticles, m_numParticles + 1));
    const float3 c = colors[rand() % numColors];
    m_colors.push_back(make_float4(c.x, c.y, c.z, 1.f));
}

void ParticleSystem::setFluidToAdd(float3 pos, float3 color, float mass, float density) {
    m_fluidsToAdd.push_back(make_float4(pos.x, pos.y, pos.z, mass));
    m_fluidsToAdd.push_back(make_float4(color.x, color.y, color.z, density));
}

// single append; a full system drops the particle like the reference (:311-312)
void ParticleSystem::addParticle(float4 pos, float4 vel, float mass, float ro, int phase) {
    if (!m_ctx || m_numParticles == m_maxParticles) return;
    float w = 1.f / mass;
    int rc = ps_append_particles(m_ctx, &pos.x, &vel.x, &w, &ro, &phase, 1);
    note(rc, "ps_append_particles");
    if (rc == PS_OK) m_numParticles++;
}

// batch append; the reference rejects a batch that would reach maxParticles (">=", :335-336)
void ParticleSystem::addParticleMultiple(float *pos, float *vel, float *mass, float *ro, int *phase, int numParticles) {
    if (!m_ctx || m_numParticles + numParticles >= m_maxParticles) {
        if (m_ctx && m_error.empty()) m_error = "addParticleMultiple: batch dropped, maxParticles reached (reference behaviour)";
        return;
    }
    int rc = ps_append_particles(m_ctx, pos, vel, mass, ro, phase, (uint64_t)numParticles);
    note(rc, "ps_append_particles");
    if (rc == PS_OK) m_numParticles += numParticles;
}

// lattice extents exactly as the reference evaluates them: (int)ceil(ur-ll) / spacing, truncated (:363,:416,:621)
static inline int lattice_count(int lo, int hi, float spacing) { return (int)((int)std::ceil((double)(hi - lo)) / spacing); }

// shared by addFluid / addParticleGrid: z-major, x-fastest lattice with per-axis jitter drawn x,y,z (:374-392)
static void fill_lattice(std::vector<float> &pos, int3 ll, int3 count, float distance, float jitter) {
    size_t index = 0;
    for (int z = 0; z < count.z; z++)
        for (int y = 0; y < count.y; y++)
            for (int x = 0; x < count.x; x++) {
                pos[index * 4] = ll.x + x * distance + (frand() * 2.0f - 1.0f) * jitter;
                pos[index * 4 + 1] = ll.y + y * distance + (frand() * 2.0f - 1.0f) * jitter;
                pos[index * 4 + 2] = ll.z + z * distance + (frand() * 2.0f - 1.0f) * jitter;
                pos[index * 4 + 3] = 1.f;
                index++;
            }
}

void ParticleSystem::addFluid(int3 ll, int3 ur, float mass, float density, float3 color) {
    const int start = m_numParticles;
    const float jitter = m_particleRadius * 0.01f;
    const float distance = m_particleRadius * 2.5f;
    const int3 count = make_int3(lattice_count(ll.x, ur.x, distance), lattice_count(ll.y, ur.y, distance), lattice_count(ll.z, ur.z, distance));
    const size_t n = (size_t)std::max(count.x, 0) * std::max(count.y, 0) * std::max(count.z, 0);
    std::vector<float> pos(n * 4), vel(n * 4, 0.f), w(n, 1.f / mass), ro(n, density);
    std::vector<int> phase(n, PS_PHASE_FLUID);
    fill_lattice(pos, ll, count, distance, jitter);
    addParticleMultiple(pos.data(), vel.data(), w.data(), ro.data(), phase.data(), (int)n);
    m_colorIndex.push_back(make_int2(start, m_numParticles));
    m_colors.push_back(make_float4(color.x, color.y, color.z, 1.f));
}

void ParticleSystem::addParticleGrid(int3 ll, int3 ur, float mass, bool addJitter) {
    const int start = m_numParticles;
    const float jitter = addJitter ? m_particleRadius * 0.01f : 0.f;
    const float distance = m_particleRadius * 2.002f;
    const int3 count = make_int3(lattice_count(ll.x, ur.x, distance), lattice_count(ll.y, ur.y, distance), lattice_count(ll.z, ur.z, distance));
    const size_t n = (size_t)std::max(count.x, 0) * std::max(count.y, 0) * std::max(count.z, 0);
    std::vector<float> pos(n * 4), vel(n * 4, 0.f), w(n, 1.f / mass), ro(n, 1.f);
    std::vector<int> phase(n, PS_PHASE_SOLID);
    fill_lattice(pos, ll, count, distance, jitter);  // draws rand() even when jitter == 0, like the reference
    addParticleMultiple(pos.data(), vel.data(), w.data(), ro.data(), phase.data(), (int)n);
    m_colorIndex.push_back(make_int2(start, m_numParticles));
    const float3 c = colors[rand() % numColors];
    m_colors.push_back(make_float4(c.x, c.y, c.z, 1.f));
}

// Headless addition, not in the reference (its GPU solver has no rigid bodies: rigid_body_functor is an empty stub,
// solver_kernel.cuh:289-312): a lattice box like addParticleGrid whose particles form ONE shape-matched body (ps_add_rigid_body)
// with its own phase RIGID + k, and, when `sdf`, the box's signed-distance data the way the reference CPU app's builders write
// it down (cpu/src/simulation.cpp:666-672): outward normal of the nearest face (normalised sum on edges and corners) and depth
// (layer + 1/2) * diameter * sqrt(number of nearest faces), so that boxes meet each other through SDF contacts.
int ParticleSystem::addRigidBox(int3 ll, int3 ur, float mass, bool sdf, float stiffness) {
    const int start = m_numParticles;
    const float distance = m_particleRadius * 2.002f;
    const int3 count = make_int3(lattice_count(ll.x, ur.x, distance), lattice_count(ll.y, ur.y, distance), lattice_count(ll.z, ur.z, distance));
    const size_t n = (size_t)std::max(count.x, 0) * std::max(count.y, 0) * std::max(count.z, 0);
    if (n < 2) { m_error = "addRigidBox: a rigid body needs at least 2 particles"; return -1; }
    std::vector<float> pos(n * 4), vel(n * 4, 0.f), w(n, 1.f / mass), ro(n, 1.f);
    std::vector<int> phase(n, PS_PHASE_RIGID + m_rigidIndex);
    fill_lattice(pos, ll, count, distance, 0.f);
    addParticleMultiple(pos.data(), vel.data(), w.data(), ro.data(), phase.data(), (int)n);
    if ((size_t)(m_numParticles - start) != n) { m_error = "addRigidBox: no room for the box"; return -1; }
    m_rigidIndex++;
    m_colorIndex.push_back(make_int2(start, m_numParticles));
    const float3 c = colors[rand() % numColors];
    m_colors.push_back(make_float4(c.x, c.y, c.z, 1.f));
    std::vector<uint> idx(n);
    for (size_t k = 0; k < n; k++) idx[k] = (uint)(start + k);
    uint32_t body = 0;
    if (ps_add_rigid_body(m_ctx, idx.data(), n, stiffness, &body) != PS_OK) { m_error = ps_last_error(); return -1; }
    if (sdf) {
        std::vector<float> data(n * 4);
        size_t k = 0;
        for (int z = 0; z < count.z; z++)
            for (int y = 0; y < count.y; y++)
                for (int x = 0; x < count.x; x++, k++) {  // fill_lattice's order
                    const int depth[6] = {x, count.x - 1 - x, y, count.y - 1 - y, z, count.z - 1 - z};
                    const float nrm[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
                    int lo = depth[0];
                    for (int f = 1; f < 6; f++) lo = std::min(lo, depth[f]);
                    float g[3] = {0, 0, 0};
                    int faces = 0;
                    for (int f = 0; f < 6; f++)
                        if (depth[f] == lo) { g[0] += nrm[f][0]; g[1] += nrm[f][1]; g[2] += nrm[f][2]; faces++; }
                    float len = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
                    if (len < 1e-6f) { g[0] = 0.f; g[1] = 1.f; g[2] = 0.f; len = 1.f; faces = 1; }  // opposite faces equally near
                    data[4 * k] = g[0] / len; data[4 * k + 1] = g[1] / len; data[4 * k + 2] = g[2] / len;
                    data[4 * k + 3] = (lo + 0.5f) * 2.f * m_particleRadius * std::sqrt((float)faces);
                }
        if (ps_set_rigid_body_sdf(m_ctx, body, data.data()) != PS_OK) { m_error = ps_last_error(); return -1; }
    }
    return (int)body;
}

// horizontal cloth in the xz plane at height spacing.y: one distance constraint to the -x neighbour and one to
// the -z neighbour per particle, pins on the x == 0 column (all four edges when holdEdges) (reference :461-568)
void ParticleSystem::addHorizCloth(int2 ll, int2 ur, float3 spacing, float2 dist, float mass, bool holdEdges) {
    const int start = m_numParticles;
    const int2 count = make_int2((int)((int)std::ceil((double)(ur.x - ll.x)) / spacing.x), (int)((int)std::ceil((double)(ur.y - ll.y)) / spacing.z));
    const size_t n = (size_t)std::max(count.x, 0) * std::max(count.y, 0);
    std::vector<float> pos(n * 4), vel(n * 4, 0.f), w(n, 1.f / mass), ro(n, 1.f);
    std::vector<int> phase(n, PS_PHASE_RIGID + m_rigidIndex);
    std::vector<uint> pinIdx, distIdx;
    std::vector<float> pinXyz, distRest;
    auto pin = [&](uint particle, const float *p) { pinIdx.push_back(particle); pinXyz.insert(pinXyz.end(), p, p + 3); };
    size_t index = 0;
    for (int z = 0; z < count.y; z++)
        for (int x = 0; x < count.x; x++) {
            float *p = &pos[index * 4];
            p[0] = ll.x + x * spacing.x;
            p[1] = spacing.y;
            p[2] = ll.y + z * spacing.z;
            p[3] = 1.f;
            const uint particle = start + z * count.x + x;
            if (x > 0) { distIdx.push_back(particle - 1); distIdx.push_back(particle); distRest.push_back(dist.x); }
            else pin(particle, p);
            if (z > 0) { distIdx.push_back(particle - count.x); distIdx.push_back(particle); distRest.push_back(dist.y); }
            else if (holdEdges) pin(particle, p);
            if (x == count.x - 1 && holdEdges) pin(particle, p);
            if (z == count.y - 1 && holdEdges) pin(particle, p);
            index++;
        }
    const uint before = m_numParticles;
    addParticleMultiple(pos.data(), vel.data(), w.data(), ro.data(), phase.data(), (int)n);
    if (m_numParticles != before || n == 0) {  // constraints only make sense if the batch was accepted
        note(ps_add_point_constraints(m_ctx, pinIdx.data(), pinXyz.data(), pinIdx.size()), "ps_add_point_constraints");
        note(ps_add_distance_constraints(m_ctx, distIdx.data(), distRest.data(), distRest.size()), "ps_add_distance_constraints");
    }
    m_colorIndex.push_back(make_int2(start, m_numParticles));
    const float3 c = colors[rand() % numColors];
    m_colors.push_back(make_float4(c.x, c.y, c.z, 1.f));
    m_rigidIndex++;
}

// chain of numLinks distance constraints, optionally pinned at its first particle (reference :570-616)
void ParticleSystem::addRope(float3 start, float3 spacing, float dist, int numLinks, float mass, bool constrainStart) {
    const uint startI = m_numParticles;
    const size_t n = (size_t)numLinks + 1;
    std::vector<float> pos(n * 4), vel(n * 4, 0.f), w(n, 1.f / mass), ro(n, 1.f);
    std::vector<int> phase(n, PS_PHASE_RIGID + m_rigidIndex);
    std::vector<uint> distIdx;
    std::vector<float> distRest;
    pos[0] = start.x; pos[1] = start.y; pos[2] = start.z; pos[3] = 1.f;
    for (int i = 1; i <= numLinks; i++) {
        pos[i * 4] = start.x + i * spacing.x;  // start + i * spacing, component-wise float (helper_math operators)
        pos[i * 4 + 1] = start.y + i * spacing.y;
        pos[i * 4 + 2] = start.z + i * spacing.z;
        pos[i * 4 + 3] = 1.f;
        distIdx.push_back(startI + i - 1); distIdx.push_back(startI + i); distRest.push_back(dist);
    }
    const uint before = m_numParticles;
    addParticleMultiple(pos.data(), vel.data(), w.data(), ro.data(), phase.data(), (int)n);
    if (m_numParticles != before) {
        note(ps_add_distance_constraints(m_ctx, distIdx.data(), distRest.data(), distRest.size()), "ps_add_distance_constraints");
        if (constrainStart) note(ps_add_point_constraints(m_ctx, &startI, &start.x, 1), "ps_add_point_constraints");
    }
    m_colorIndex.push_back(make_int2(startI, m_numParticles));
    const float3 c = colors[rand() % numColors];
    m_colors.push_back(make_float4(c.x, c.y, c.z, 1.f));
    m_rigidIndex++;
}

// lattice points inside a ball, every one pinned in place with inverse mass 0.01 (reference :618-687)
void ParticleSystem::addStaticSphere(int3 ll, int3 ur, float spacing) {
    const uint startI = m_numParticles;
    const int3 count = make_int3(lattice_count(ll.x, ur.x, spacing), lattice_count(ll.y, ur.y, spacing), lattice_count(ll.z, ur.z, spacing));
    const float radius = (ur.x - ll.x) * .5f;
    const float cx = ll.x + radius, cy = ll.y + radius, cz = ll.z + radius;
    std::vector<float> pos, pinXyz;
    std::vector<uint> pinIdx;
    uint index = 0;
    for (int z = 0; z < count.z; z++)
        for (int y = 0; y < count.y; y++)
            for (int x = 0; x < count.x; x++) {
                const float px = ll.x + x * spacing, py = ll.y + y * spacing, pz = ll.z + z * spacing;
                const float dx = px - cx, dy = py - cy, dz = pz - cz;
                if (sqrtf(dx * dx + dy * dy + dz * dz) < radius) {
                    pos.insert(pos.end(), {px, py, pz, 1.f});
                    pinIdx.push_back(startI + index++);
                    pinXyz.insert(pinXyz.end(), {px, py, pz});
                }
            }
    const size_t n = pinIdx.size();
    std::vector<float> vel(n * 4, 0.f), w(n, .01f), ro(n, 1.f);
    std::vector<int> phase(n, PS_PHASE_RIGID + m_rigidIndex);
    const uint before = m_numParticles;
    addParticleMultiple(pos.data(), vel.data(), w.data(), ro.data(), phase.data(), (int)n);
    if (m_numParticles != before) note(ps_add_point_constraints(m_ctx, pinIdx.data(), pinXyz.data(), n), "ps_add_point_constraints");
    m_colorIndex.push_back(make_int2(startI, m_numParticles));
    const float3 c = colors[rand() % numColors];
    m_colors.push_back(make_float4(c.x, c.y, c.z, 1.f));
    m_rigidIndex++;
}

void ParticleSystem::makePointConstraint(uint index, float3 point) { note(ps_add_point_constraints(m_ctx, &index, &point.x, 1), "ps_add_point_constraints"); }
void ParticleSystem::makeDistanceConstraint(uint2 index, float distance) { note(ps_add_distance_constraints(m_ctx, &index.x, &distance, 1), "ps_add_distance_constraints"); }

void ParticleSystem::getPositions(float *h) const { if (m_ctx) ps_download(m_ctx, PS_ARR_POS, h, 0, 4ull * m_numParticles); }
void ParticleSystem::getVelocities(float *h) const { if (m_ctx) ps_download(m_ctx, PS_ARR_VEL, h, 0, 4ull * m_numParticles); }
const float *ParticleSystem::devicePositions() const { return m_ctx ? (const float *)ps_device_ptr(m_ctx, PS_ARR_POS) : nullptr; }
void ParticleSystem::sync() const { if (m_ctx) ps_sync(m_ctx); }

}  // namespace psb200

// ------------------------------------------------------------------------------------------------------------------
// C entry points over the host class, so that Python (ctypes) tests and bench.py drive the same C++ host code a
// C++ application would (the reference has no such layer: its only client is its Qt app).
// ------------------------------------------------------------------------------------------------------------------
using psb200::ParticleSystem;
extern "C" {
void *pshost_create(float radius, unsigned gx, unsigned gy, unsigned gz, unsigned maxParticles, const int *minB, const int *maxB, int iterations) {
    return new ParticleSystem(radius, make_uint3(gx, gy, gz), maxParticles, make_int3(minB[0], minB[1], minB[2]), make_int3(maxB[0], maxB[1], maxB[2]), iterations);
}
// scene: "1".."9", "c2", "c3" (include/ps_scenes.h).  Reseeds glibc rand() with `seed` first when seed >= 0 (the
// reference never seeds: seed 1 reproduces a fresh process).
void *pshost_build_scene(const char *scene, int grid, unsigned maxParticles, int iterations, int side, int seed) {
    if (seed >= 0) srand((unsigned)seed);
    if (std::string(scene) == "r") return psb200::build_rigid_scene(grid, maxParticles, iterations);
    ps_scenes::SceneSpec s;
    s.scene = scene; s.grid = grid; s.max_particles = maxParticles; s.iterations = iterations; s.side = side;
    return ps_scenes::build<ParticleSystem>(s, psb200::colors, psb200::numColors);
}
void pshost_destroy(void *h) { delete (ParticleSystem *)h; }
PsCtx *pshost_ctx(void *h) { return ((ParticleSystem *)h)->context(); }
const char *pshost_error(void *h) { return ((ParticleSystem *)h)->lastError().c_str(); }
unsigned pshost_num_particles(void *h) { return ((ParticleSystem *)h)->getNumParticles(); }
void pshost_update(void *h, float dt) { ((ParticleSystem *)h)->update(dt); }
void pshost_add_fluid(void *h, const int *ll, const int *ur, float mass, float density) {
    ((ParticleSystem *)h)->addFluid(make_int3(ll[0], ll[1], ll[2]), make_int3(ur[0], ur[1], ur[2]), mass, density, make_float3(0, 0, 1));
}
void pshost_add_particle_grid(void *h, const int *ll, const int *ur, float mass, int addJitter) {
    ((ParticleSystem *)h)->addParticleGrid(make_int3(ll[0], ll[1], ll[2]), make_int3(ur[0], ur[1], ur[2]), mass, addJitter != 0);
}
void pshost_add_horiz_cloth(void *h, const int *ll, const int *ur, const float *spacing, const float *dist, float mass, int holdEdges) {
    ((ParticleSystem *)h)->addHorizCloth(make_int2(ll[0], ll[1]), make_int2(ur[0], ur[1]), make_float3(spacing[0], spacing[1], spacing[2]),
                                         make_float2(dist[0], dist[1]), mass, holdEdges != 0);
}
void pshost_add_rope(void *h, const float *start, const float *spacing, float dist, int numLinks, float mass, int constrainStart) {
    ((ParticleSystem *)h)->addRope(make_float3(start[0], start[1], start[2]), make_float3(spacing[0], spacing[1], spacing[2]), dist, numLinks, mass,
                                   constrainStart != 0);
}
int pshost_add_rigid_box(void *h, const int *ll, const int *ur, float mass, int sdf, float stiffness) {
    return ((ParticleSystem *)h)->addRigidBox(make_int3(ll[0], ll[1], ll[2]), make_int3(ur[0], ur[1], ur[2]), mass, sdf != 0, stiffness);
}
void pshost_add_static_sphere(void *h, const int *ll, const int *ur, float spacing) {
    ((ParticleSystem *)h)->addStaticSphere(make_int3(ll[0], ll[1], ll[2]), make_int3(ur[0], ur[1], ur[2]), spacing);
}
void pshost_set_particle_to_add(void *h, const float *pos, const float *vel, float mass) {
    ((ParticleSystem *)h)->setParticleToAdd(make_float3(pos[0], pos[1], pos[2]), make_float3(vel[0], vel[1], vel[2]), mass);
}
void pshost_set_fluid_to_add(void *h, const float *pos, const float *color, float mass, float density) {
    ((ParticleSystem *)h)->setFluidToAdd(make_float3(pos[0], pos[1], pos[2]), make_float3(color[0], color[1], color[2]), mass, density);
}
void pshost_make_point_constraint(void *h, unsigned index, const float *point) {
    ((ParticleSystem *)h)->makePointConstraint(index, make_float3(point[0], point[1], point[2]));
}
void pshost_make_distance_constraint(void *h, unsigned a, unsigned b, float distance) {
    ((ParticleSystem *)h)->makeDistanceConstraint(make_uint2(a, b), distance);
}
void pshost_get_positions(void *h, float *out4n) { ((ParticleSystem *)h)->getPositions(out4n); }
void pshost_get_velocities(void *h, float *out4n) { ((ParticleSystem *)h)->getVelocities(out4n); }
}
