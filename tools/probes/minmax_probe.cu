// minmax_probe.cu — issue rates of the packed 16-bit min/max family and of HFMA2-based substitutes on sm_100a
// (which pipe bounds the top-2 epilogue of match_tc.cu, and how much of it can move to the FMA pipe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define DEF2(name, OP) __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b){ uint32_t d; asm volatile(OP " %0, %1, %2;" : "=r"(d) : "r"(a),"r"(b)); return d;}
DEF2(hmax2, "max.f16x2")
DEF2(hmin2, "min.f16x2")
DEF2(hadd2, "add.rn.f16x2")
DEF2(hsub2, "sub.rn.f16x2")
DEF2(smax2, "max.s16x2")
DEF2(smin2, "min.s16x2")
__device__ __forceinline__ uint32_t hfma_relu(uint32_t a, uint32_t b, uint32_t c){ uint32_t d; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a),"r"(b),"r"(c)); return d;}
__device__ __forceinline__ uint32_t hfma(uint32_t a, uint32_t b, uint32_t c){ uint32_t d; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a),"r"(b),"r"(c)); return d;}

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = seed * (threadIdx.x + 1) * (i + 3); y[i] = seed + i * 77 + threadIdx.x; }
    const uint32_t one = 0x3C003C00u, neg1 = 0xBC00BC00u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i] = hmax2(x[i], y[i]); y[i] = hmin2(y[i], x[(i + 1) & 7]); }                 // 2 HMNMX2
            if (MODE == 1) { x[i] = hmax2(hmax2(x[i], y[i]), y[(i + 3) & 7]); y[i] = hmin2(hmin2(y[i], x[(i + 1) & 7]), x[(i+5)&7]); }   // 2 VHMNMX (fused by ptxas)
            if (MODE == 2) { x[i] = hfma(x[i], one, y[i]); y[i] = hfma(y[i], neg1, x[(i + 1) & 7]); }        // 2 HFMA2
            if (MODE == 3) { x[i] = hfma_relu(x[i], one, y[i]); y[i] = hadd2(y[i], x[(i + 1) & 7]); }        // HFMA2.RELU + HADD2
            if (MODE == 4) { x[i] = smax2(x[i], y[i]); y[i] = smin2(y[i], x[(i + 1) & 7]); }                 // 2 VIMNMX.S16x2
            if (MODE == 5) { x[i] = hmax2(x[i], y[i]); y[i] = hfma(y[i], neg1, x[(i + 1) & 7]); }            // 1 HMNMX2 + 1 HFMA2
            if (MODE == 6) { x[i] = hmax2(x[i], y[i]); y[i] = hmin2(y[i], x[(i + 1) & 7]); x[(i+2)&7] = hfma(x[(i+2)&7], one, y[i]); y[(i+2)&7] = hfma(y[(i+2)&7], neg1, x[i]); }   // 2 + 2
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int ops_per_inner, uint32_t* out, int sms) {
    const int iters = 4096, blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(out, 16, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 /*warps*/ * iters * 8 * ops_per_inner;
    printf("%-34s %8.3f ms  %7.2f G warp-instr/s  = %.3f warp-instr/clk/SM @1.9GHz\n", name, ms, warp_instr / ms * 1e-6,
           warp_instr / (ms * 1e-3) / sms / 1.9e9);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; cudaMalloc(&out, sms * 8 * 256 * 4);
    run<0>("HMNMX2 x2", 2, out, sms);
    run<1>("VHMNMX (3-input) x2", 2, out, sms);
    run<2>("HFMA2 x2", 2, out, sms);
    run<3>("HFMA2.RELU + HADD2", 2, out, sms);
    run<4>("VIMNMX.S16x2 x2", 2, out, sms);
    run<5>("HMNMX2 + HFMA2", 2, out, sms);
    run<6>("2 HMNMX2 + 2 HFMA2", 4, out, sms);
    return 0;
}
