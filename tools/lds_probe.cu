// Micro-benchmark: cost of LDS.{32,64,128} for different lane->address patterns (wavefront model).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_probe lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int W>
__device__ __forceinline__ float lds(unsigned addr) {
    float a, b, c, d;
    if (W == 4) {
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
        return a;
    } else if (W == 2) {
        asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
        return a;
    } else {
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(a) : "r"(addr));
        return a;
    }
}

__device__ int pattern(int p, int lane) {
    const int l = lane;
    switch (p) {
    case 0: return (l % 2) + 2 * (l / 4);                  // 16 uniq: b0 + bits 2..4
    case 1: return (l >> 2) * 17 + ((l >> 1) & 1) * 2;     // X-type: mg rows (stride 17) x kt
    case 2: return (l >> 2) * 9 + (l & 1) * 2;             // dY0-type
    case 3: return (l >> 2) + (l & 1) * 34;                // dYk-type
    case 4: return (l >> 1) * 17;                          // 16 rows pair-blocked
    case 5: return (l & 1) * 8 + (l >> 4) * 16;            // 4 uniq: b0 and b4
    case 6: return ((l >> 1) & 7) * 17 + (l >> 4) * 2;     // 16 uniq: bits1-3 rows, bit4 col : pair-blocked
    case 7: return (l & 1) * 17 + ((l >> 4) & 1) * 34 + 0; // y-type 4 uniq rows
    case 8: return (l >> 1) + (l & 1) * 64;                // 32 uniq, pair far apart (float4 units)
    case 9: return l * 9;                                  // 32 rows stride 9 float4
    case 10: return (l >> 3) * 17 + (l & 7);               // 32 uniq: 8 consecutive per quarter
    case 11: return (l & 15) * 17 + (l >> 4) * 4;          // 32 uniq
    case 12: return (l >> 1) * 9;                          // 16 rows stride 9
    case 13: return (l >> 2) * 17;                         // 8 rows quad-blocked
    default: return lane;
    }
}

template <int W>
__global__ void probe(int p, int iters, float *out, long long *cyc) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + pattern(p, lane) * (W * 4);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; u += 4) {
            acc0 += lds<W>(base + ((it + u) & 3) * 16);
            acc1 += lds<W>(base + ((it + u + 1) & 3) * 16);
            acc2 += lds<W>(base + ((it + u + 2) & 3) * 16);
            acc3 += lds<W>(base + ((it + u + 3) & 3) * 16);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float acc = acc0 + acc1 + acc2 + acc3;
    if (acc == 1.2345f) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char *names[] = {"(l%2)+2*(l/4)","X: (l>>2)*17+((l>>1)&1)*2","dY0: (l>>2)*9+(l&1)*2","dYk: (l>>2)+(l&1)*34","(l>>1)*17","(l&1)*8+(l>>4)*16","((l>>1)&7)*17+(l>>4)*2","(l&1)*17+((l>>4)&1)*34","(l>>1)+(l&1)*64","l*9","(l>>3)*17+(l&7)","(l&15)*17+(l>>4)*4","(l>>1)*9","(l>>2)*17"};
    for (int w = 4; w >= 4; w /= 2) {
        for (int p = 0; p < 14; ++p) {
            for (int warps = 16; warps <= 16; warps *= 2) {
                long long c = 0;
                cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); if (w == 4) probe<4><<<148, warps * 32, 65536>>>(p, iters, out, cyc);
                else if (w == 2) probe<2><<<148, warps * 32, 65536>>>(p, iters, out, cyc);
                else probe<1><<<148, warps * 32, 65536>>>(p, iters, out, cyc);
                cudaDeviceSynchronize();
                cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
                double per = (double)c / (iters * 16.0 * warps);
                printf("LDS.%-3d %-48s warps=%d  cycles per warp-instr (SM-wide) = %.2f\n", w * 32, names[p], warps, per);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
