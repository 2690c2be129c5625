// Throughput probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.  Prints GFMA-lanes/s for both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
constexpr int ILP = 8, ITERS = 4096;
__global__ void k1(float* out, float a, float b) {
    float v[ILP];
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = fma1(v[i], a, b);
    float s = 0; for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k2(float* out, float a, float b) {
    u64 v[ILP], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int i = 0; i < ILP; i++) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(x)); }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = fma2(v[i], aa, bb);
    float s = 0;
    for (int i = 0; i < ILP; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; which++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (which == 0) k1<<<148 * 8, 512>>>(out, 1.0001f, 0.5f); else k2<<<148 * 8, 512>>>(out, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double lanes = 148.0 * 8 * 512 * ILP * ITERS * (which ? 2 : 1);
            if (rep == 2) printf("%s: %.3f ms, %.1f G fma-lanes/s (%.1f TFLOP/s)\n", which ? "FFMA2" : "FFMA ", ms, lanes / ms / 1e6, 2 * lanes / ms / 1e9);
        }
    }
    return 0;
}
