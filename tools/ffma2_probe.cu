// FFMA vs FFMA2 (fma.rn.f32x2) throughput on sm_100a.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(unsigned long long &c, float a, unsigned long long w) {
    unsigned long long A;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(A), "l"(w));
}
template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, float *sink) {
    const float a = 1.0f + 1e-7f * threadIdx.x, b = 1e-9f * blockIdx.x;
    float s = 0.f;
    if (MODE == 0) {
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = (float)j;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], a, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) s += acc[j];
    } else {
        unsigned long long acc[16], w;
        asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(b), "f"(b + 1e-9f));
#pragma unroll
        for (int j = 0; j < 16; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[j]) : "f"((float)j), "f"((float)j + 0.5f));
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 16; ++j) ffma2(acc[j], a, w);      // acc += a * w  (2 FMAs)
#pragma unroll
        for (int j = 0; j < 16; ++j) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[j])); s += lo + hi; }
    }
    if (s == 123.456f) sink[0] = s;
}
int main() {
    float *sink; cudaMalloc(&sink, 4);
    const int iters = 20000, ctas = 148 * 8;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<ctas, 256>>>(iters, sink); else probe<1><<<ctas, 256>>>(iters, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 32 * 4 * (double)iters * 256 * ctas;
            printf("%s: %.3f ms  %.2f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, flops / (ms * 1e-3) / 1e12);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
