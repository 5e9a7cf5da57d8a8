/*
 * include/particle_system.h — headless C++ host class with the public interface of the reference's
 * ParticleSystem (gpu/src/particlesystem.h:22-118): same constructor, same scene builders, same per-frame
 * update(), same CUDA vector argument types, so code written against the reference's class compiles against
 * this one.  It owns no GL objects (Qt/GL is for viewing only) and no process-global state: it drives
 * libpsolver.so through the C ABI in include/psolver.h and nothing else.
 *
 * Differences from the reference, all deliberate:
 *   - scene arrays live on the heap (the reference uses stack VLAs, particlesystem.cpp:366-370) and sizes are
 *     64-bit inside the library; maxParticles may be in the tens of millions;
 *   - getCurrentReadBuffer() returns 0 (no VBO); positions are read with getPositions() or, zero-copy, through
 *     devicePositions();
 *   - lastError() reports what the reference either ignores (capacity overflow, particlesystem.cpp:311,335) or
 *     turns into exit(EXIT_FAILURE) (CUDA errors, helper_cuda.h:981-1008).
 */
#ifndef PS_PARTICLE_SYSTEM_H
#define PS_PARTICLE_SYSTEM_H

#include <vector_types.h>
#include <vector_functions.h>
#include <deque>
#include <string>
#include <vector>
#include "psolver.h"

namespace psb200 {

typedef unsigned int GLuint;
typedef unsigned int uint;

/* reference particlesystem.h:12-20 */
const int numColors = 8;
extern const float3 colors[numColors];

class ParticleSystem {
public:
    ParticleSystem(float particleRadius, uint3 gridSize, uint maxParticles, int3 minBounds, int3 maxBounds, int iterations);
    ~ParticleSystem();
    ParticleSystem(const ParticleSystem &) = delete;
    ParticleSystem &operator=(const ParticleSystem &) = delete;

    void update(float deltaTime);
    void resetGrid() {} /* declared but never defined in the reference (particlesystem.h:29) */

    void addFluid(int3 ll, int3 ur, float mass, float density, float3 color);
    void addParticleGrid(int3 ll, int3 ur, float mass, bool addJitter);
    void addHorizCloth(int2 ll, int2 ur, float3 spacing, float2 dist, float mass, bool holdEdges);
    void addRope(float3 start, float3 spacing, float dist, int numLinks, float mass, bool constrainStart);
    void addStaticSphere(int3 ll, int3 ur, float spacing);

    void setParticleToAdd(float3 pos, float3 vel, float mass);
    void setFluidToAdd(float3 pos, float3 color, float mass, float density);

    void makePointConstraint(uint index, float3 point);
    void makeDistanceConstraint(uint2 index, float distance);

    std::vector<int2> getColorIndex() { return m_colorIndex; }
    std::vector<float4> getColors() { return m_colors; }

    GLuint getCurrentReadBuffer() const { return 0; }
    uint getNumParticles() const { return m_numParticles; }
    float getParticleRadius() const { return m_particleRadius; }
    int3 getMinBounds() { return m_minBounds; }
    int3 getMaxBounds() { return m_maxBounds; }

    /* ---- headless additions ---- */
    /* not in the reference (its GPU solver has no rigid bodies): a lattice box that is one shape-matched body with its own phase,
     * with the box's SDF data when `sdf` (contacts between such boxes then follow their surfaces); returns the body index or -1 */
    int addRigidBox(int3 ll, int3 ur, float mass, bool sdf = true, float stiffness = 1.f);
    PsCtx *context() const { return m_ctx; }
    bool ok() const { return m_ctx != nullptr && m_error.empty(); }
    const std::string &lastError() const { return m_error; }
    void getPositions(float *host4n) const;      /* blocking device->host copy of float4[n] */
    void getVelocities(float *host4n) const;
    const float *devicePositions() const;        /* device pointer, float4[n] */
    void sync() const;

private:
    void addParticle(float4 pos, float4 vel, float mass, float ro, int phase);
    void addParticleMultiple(float *pos, float *vel, float *mass, float *ro, int *phase, int numParticles);
    void addParticles();
    void addFluids();
    void addNewStuff();
    void note(int rc, const char *where);

    PsCtx *m_ctx;
    std::string m_error;
    float m_particleRadius;
    uint m_maxParticles;
    uint m_numParticles;
    uint3 m_gridSize;
    int m_rigidIndex;
    std::deque<float4> m_particlesToAdd;
    std::deque<float4> m_fluidsToAdd;
    std::vector<int2> m_colorIndex;
    std::vector<float4> m_colors;
    int3 m_minBounds;
    int3 m_maxBounds;
    uint m_solverIterations;
};

/* Extension scene "r" (psolver_cli --app gpu --scene r; not one of the reference's): rigid boxes with SDF contacts — a tower of
 * three 5x5x5 boxes dropped onto each other, a fourth box beside it, a loose pile of solid particles beside. */
inline ParticleSystem *build_rigid_scene(int grid, uint maxParticles, int iterations) {
    ParticleSystem *ps = new ParticleSystem(0.25f, make_uint3(grid, grid, grid), maxParticles, make_int3(-50, 0, -50), make_int3(50, 200, 50), iterations);
    ps->addRigidBox(make_int3(0, 1, 0), make_int3(3, 4, 3), 1.f);   /* 5 x 5 x 5 particles each: two layers deep */
    ps->addRigidBox(make_int3(0, 5, 0), make_int3(3, 8, 3), 1.f);
    ps->addRigidBox(make_int3(0, 9, 0), make_int3(3, 12, 3), 1.f);
    ps->addRigidBox(make_int3(-6, 3, 0), make_int3(-3, 6, 3), 1.f);  /* a fourth box beside the tower */
    ps->addParticleGrid(make_int3(8, 0, 0), make_int3(10, 3, 2), 1.f, false);
    return ps;
}

}  // namespace psb200
#endif
