/*
 * include/ps_reference_abi.h — the reference's own extern "C" boundary, re-exported by libpsolver.so.
 *
 * Every function below keeps the NAME, ARGUMENT ORDER and MEANING it has in ebirenbaum/ParticleSolver, so
 * the reference's host class (gpu/src/particlesystem.cpp) links against libpsolver.so in place of its
 * gpu/src/cuda/{integration,solver,shared_variables}.cu translation units.
 * Declarations replaced, with the reference file:line of each:
 *   gpu/src/cuda/wrappers.cuh:12-97            integration + solver wrappers
 *   gpu/src/cuda/shared_variables.cuh:13-36    phase / inverse mass / Xstar state
 *   gpu/src/cuda/util.cuh:6-25                 device selection, raw device arrays, blocking copies, launch geometry
 * Not re-exported: the four GL-interop entry points of util.cuh (registerGLBufferObject, unregisterGLBufferObject,
 * mapGLBufferObject, unmapGLBufferObject, util.cuh:16-20): they belong to the viewer; the reference's util.cu keeps
 * them for the Qt application, a headless host maps an ordinary device allocation (oracle/ref_gpu_glue.cu).
 *
 * Like the reference: one particle system per process (state lives in a library-global context), default
 * stream, blocking copies of host inputs, and any CUDA failure prints a message and exit(EXIT_FAILURE)s
 * (reference helper_cuda.h:981-1008).  Device pointers are caller-owned exactly as in the reference
 * (ParticleSystem allocates them: particlesystem.cpp:95-107).
 */
#ifndef PS_REFERENCE_ABI_H
#define PS_REFERENCE_ABI_H
#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int uint;

/* byte-compatible with SimParams, gpu/src/cuda/kernel.cuh:9-22 (float3/uint3 are three packed 4-byte fields) */
typedef struct PsRefSimParams {
    float gravity[3];
    float globalDamping;
    float particleRadius;
    unsigned int gridSize[3];
    unsigned int numCells;
    float worldOrigin[3];
    float cellSize[3];
    unsigned int numBodies;
    unsigned int maxParticlesPerCell;
} PsRefSimParams;
/* ABI-compatible with CUDA's int3 when passed by value */
typedef struct PsRefInt3 { int x, y, z; } PsRefInt3;

#ifndef PS_REFERENCE_ABI_NO_PROTOTYPES /* define when the reference's own wrappers.cuh is also in scope */
/* ---- integration.cu ---- (wrappers.cuh:20-66,85-96) */
void initIntegration(void);
void freeIntegrationVectors(void);
void appendIntegrationParticle(float *v, float *ro, uint numParticles);
void setParameters(PsRefSimParams *hostParams);
void integrateSystem(float *pos, float deltaTime, uint numParticles);
void calcHash(uint *gridParticleHash, uint *gridParticleIndex, float *pos, int numParticles);
void sortParticles(uint *dGridParticleHash, uint *dGridParticleIndex, uint numParticles);
void reorderDataAndFindCellStart(uint *cellStart, uint *cellEnd, float *sortedPos, float *sortedW, int *sortedPhase,
                                 uint *gridParticleHash, uint *gridParticleIndex, float *oldPos, uint numParticles,
                                 uint numCells);
void collideWorld(float *pos, float *sortedPos, uint numParticles, PsRefInt3 minBounds, PsRefInt3 maxBounds);
void collide(float *particles, float *sortedPos, float *sortedW, int *sortedPhase, uint *gridParticleIndex,
             uint *cellStart, uint *cellEnd, uint numParticles, uint numCells);
void sortByType(float *dPos, uint numParticles);
void calcVelocity(float *dpos, float deltaTime, uint numParticles);
void solveFluids(float *sortedPos, float *sortedW, int *sortedPhase, uint *gridParticleIndex, uint *cellStart,
                 uint *cellEnd, float *particles, uint numParticles, uint numCells);
/* ---- solver.cu ---- (wrappers.cuh:69-83) */
void appendSolverParticle(uint numParticles);
void addPointConstraint(uint *index, float *point, uint numConstraints);
void addDistanceConstraint(uint *index, float *distance, uint numConstraints);
void freeSolverVectors(void);
void solvePointConstraints(float *particles);
void solveDistanceConstraints(float *particles);
/* ---- util.cu without OpenGL ---- (util.cuh:9-14,22-25) */
void cudaInit(void);
void allocateArray(void **devPtr, int size);
void freeArray(void *devPtr);
void copyArrayToDevice(void *device, const void *host, int offset, int size);
void copyArrayFromDevice(void *host, const void *device, int size);
uint iDivUp(uint a, uint b);
#ifdef __cplusplus
void computeGridSize(uint n, uint blockSize, uint &numBlocks, uint &numThreads);
#endif
/* ---- shared_variables.cu ---- (shared_variables.cuh:13-36) */
void freeSharedVectors(void);
void appendPhaseAndMass(int *fase, float *w, uint numParticles);
void copyToXstar(float *pos, uint numParticles);
int *getPhaseRawPtr(void);
float *getXstarRawPtr(void);
float *getWRawPtr(void);
void printXstar(void);
#endif

/* ---- additions (not in the reference): raw pointers to the state the reference keeps in file-scope
 * thrust vectors (integration.cu:23-34, solver.cu:32-41), for tests and viewers ---- */
float *psRefVelocityPtr(void);      /* V            float4[n] */
float *psRefLambdaPtr(void);        /* lambda       float[n], by sorted slot */
float *psRefRestDensityPtr(void);   /* ros          float[n] */
uint *psRefNumNeighborsPtr(void);   /* numNeighbors uint[n], by sorted slot */
float *psRefRandsPtr(void);         /* rands        float[6], last collideWorld draw */
uint *psRefOccurrencesPtr(void);    /* occurences   uint[n] (synchronises pending constraint uploads) */
uint psRefNumDistanceConstraints(void);
uint psRefNumPointConstraints(void);
/* host copies of the constraint lists in insertion order: idx uint[2m]/rest float[m]; idx uint[p]/xyz float[3p] */
void psRefCopyDistanceConstraints(uint *idx_pairs, float *rest);
void psRefCopyPointConstraints(uint *idx, float *xyz);

#ifdef __cplusplus
}
#endif
#endif /* PS_REFERENCE_ABI_H */
