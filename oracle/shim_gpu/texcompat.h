// Force-included when compiling the reference's UNMODIFIED gpu/src/cuda/*.cu with CUDA 12.x.
// CUDA 12 removed texture references; the reference (CUDA 5 era) declares
//   texture<float4, 1, cudaReadModeElementType> oldPosTex;      (integration_kernel.cuh:45-51)
// and reads them with tex1Dfetch / binds them with cudaBindTexture.  This header maps that API
// onto a plain __device__ pointer holder so the kernels run the same loads from linear memory.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build); never part of the product library.
#pragma once
#include <cuda_runtime.h>
template <class T, int D, int M> struct texref_shim { const T *p; };
#define texture static __device__ texref_shim
template <class T, int D, int M>
__device__ __forceinline__ T tex1Dfetch(const texref_shim<T, D, M> &t, unsigned i) { return t.p[i]; }
template <class T, int D, int M>
cudaError_t cudaBindTexture(size_t *, const texref_shim<T, D, M> &t, const void *p, size_t) {
    texref_shim<T, D, M> h{(const T *)p};
    return cudaMemcpyToSymbol(t, &h, sizeof h);
}
template <class T, int D, int M>
cudaError_t cudaUnbindTexture(const texref_shim<T, D, M> &) { return cudaSuccess; }
