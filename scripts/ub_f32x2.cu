// Microbenchmark: issue rate of scalar FFMA against the packed FFMA2 (fma.rn.f32x2, new on sm_100), alone and mixed with
// integer instructions, to decide whether packing the PBF interaction bodies pays in an issue-bound kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ub_f32x2 scripts/ub_f32x2.cu && gpurun_out/ub_f32x2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
    float2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t &>(d)) : "l"(reinterpret_cast<uint64_t const &>(a)), "l"(reinterpret_cast<uint64_t const &>(b)), "l"(reinterpret_cast<uint64_t const &>(c)));
    return d;
}
constexpr int kIters = 4096, kAcc = 8;
template <int MODE>  // 0: FFMA, 1: FFMA2, 2: FFMA + IADD/LOP mix 1:1, 3: FFMA2 + int mix 1:1, 4: FFMA2 + int mix 1:2
__global__ void k(float *out, float a, float b, unsigned m) {
    float s[kAcc];
    float2 v[kAcc];
    unsigned u[kAcc];
    for (int i = 0; i < kAcc; i++) { s[i] = threadIdx.x + i; v[i] = make_float2(s[i], -s[i]); u[i] = threadIdx.x * 7 + i; }
    const float2 a2 = make_float2(a, a + 1.f), b2 = make_float2(b, b - 1.f);
#pragma unroll 1
    for (int t = 0; t < kIters; t++) {
#pragma unroll
        for (int i = 0; i < kAcc; i++) {
            if (MODE == 0 || MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a), "f"(b));
            if (MODE == 1 || MODE == 3 || MODE == 4) v[i] = f2fma(v[i], a2, b2);
            if (MODE >= 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(m));
            if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(m));
        }
    }
    float r = 0;
    for (int i = 0; i < kAcc; i++) r += s[i] + v[i].x + v[i].y + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, int inst_per_iter) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 1.0001f, 0.5f, 0x55u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, 1.0001f, 0.5f, 0x55u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)sms * 8 * 8 * kIters * kAcc * inst_per_iter;
    printf("%-28s %8.3f ms  %7.1f G warp-inst/s  (%.2f per SM-clock at 1.965 GHz)\n", name, ms, warp_inst / ms * 1e-6, warp_inst / ms * 1e-6 / (sms * 1.965));
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 1);
    run<1>("FFMA2", 1);
    run<2>("FFMA + LOP 1:1", 2);
    run<3>("FFMA2 + LOP 1:1", 2);
    run<4>("FFMA2 + LOP + IADD 1:2", 3);
    return 0;
}
