// Headless replacement for the reference's gpu/src/cuda/util.cu (81 lines, GL-interop based) plus
// the ~6 OpenGL buffer calls used by gpu/src/particlesystem.cpp.  Same exported symbols as
// util.cuh:6-25; the "VBO" is an ordinary device allocation.  TEST INFRASTRUCTURE ONLY: this is
// what lets the reference's own, unmodified solver sources run on a GPU box without Qt/GL, so
// that its outputs can pin the oracle and the CUDA product path.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <algorithm>
#include "GL/glew.h"

typedef unsigned int uint;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

struct FakeVbo { void *dptr; size_t bytes; };
static std::map<GLuint, FakeVbo> g_vbos;
static GLuint g_next = 1, g_bound = 0;

extern "C" {
void glGenBuffers(GLsizei n, GLuint *ids) { for (int i = 0; i < n; i++) { ids[i] = g_next++; g_vbos[ids[i]] = FakeVbo{nullptr, 0}; } }
void glDeleteBuffers(GLsizei n, const GLuint *ids) {
    for (int i = 0; i < n; i++) { auto it = g_vbos.find(ids[i]); if (it != g_vbos.end()) { if (it->second.dptr) cudaFree(it->second.dptr); g_vbos.erase(it); } }
}
void glBindBuffer(GLenum, GLuint id) { g_bound = id; }
void glBufferData(GLenum, GLsizeiptr size, const void *data, GLenum) {
    FakeVbo &v = g_vbos[g_bound];
    if (v.dptr) cudaFree(v.dptr);
    CK(cudaMalloc(&v.dptr, size)); v.bytes = size;
    CK(cudaMemset(v.dptr, 0, size));
    if (data) CK(cudaMemcpy(v.dptr, data, size, cudaMemcpyHostToDevice));
}
void glBufferSubData(GLenum, GLintptr off, GLsizeiptr size, const void *data) {
    FakeVbo &v = g_vbos[g_bound];
    CK(cudaMemcpy((char *)v.dptr + off, data, size, cudaMemcpyHostToDevice));
}

#ifndef PS_DROPIN  /* the drop-in build takes these seven from libpsolver.so (include/ps_reference_abi.h), like a real integration would */
void cudaInit() { CK(cudaSetDevice(0)); }
void allocateArray(void **devPtr, size_t size) { CK(cudaMalloc(devPtr, size)); }
void freeArray(void *devPtr) { CK(cudaFree(devPtr)); }
void copyArrayToDevice(void *device, const void *host, int offset, int size) {
    CK(cudaMemcpy((char *)device + offset, host, size, cudaMemcpyHostToDevice));
}
void copyArrayFromDevice(void *host, const void *device, int size) { CK(cudaMemcpy(host, device, size, cudaMemcpyDeviceToHost)); }
#endif
// the "graphics resource" handle is just the vbo id smuggled through the pointer
void registerGLBufferObject(unsigned int vbo, struct cudaGraphicsResource **res) { *res = (struct cudaGraphicsResource *)(size_t)vbo; }
void unregisterGLBufferObject(struct cudaGraphicsResource *) {}
void *mapGLBufferObject(struct cudaGraphicsResource **res) { return g_vbos[(GLuint)(size_t)*res].dptr; }
void unmapGLBufferObject(struct cudaGraphicsResource *) {}
#ifndef PS_DROPIN
uint iDivUp(uint a, uint b) { return (a % b != 0) ? (a / b + 1) : (a / b); }
void computeGridSize(uint n, uint blockSize, uint &numBlocks, uint &numThreads) {
    numThreads = std::min(blockSize, n);
    numBlocks = iDivUp(n, numThreads);
}
#endif
}
