// Dense route of the batched input-to-hidden projection (sm_100a, FP32 FFMA2).
//
// The TT-matrix x dense product may be contracted in any order.  When the input width I is small
// the cheapest order is "cores first": W_ih = G_0 x ... x G_{d-1} is formed once per call (I rows
// through the TT-matvec kernel, i.e. W^T = TTLinear(identity)), and the B*T rows of the sequence
// then need ONE dense contraction of I*G*H multiply-adds each instead of the d-stage chain
// (cfg4: 0.26 M vs 2.2 M per row; cfg3: 0.26 M vs 0.33 M; cfg5: 1.05 M vs 0.93 M but with a
// gradient pass that costs 1x instead of 3x the chain).  Gradients stay exact: dW^T = X^T delta is
// accumulated densely and pushed onto the cores by one I-row TT-matvec backward (linearity), the
// generalisation of the rank-one input mode of configs 1-2.
//
//   k_gemm_rows : C[r, :] = A[r, :] * B (+ bias)     rows x K  times  K x N   (forward xg, and dX = delta * W)
//   k_gemm_red  : C_s    = sum_{r in split s} A[r, :]^T B[r, :]                (dW^T = X^T delta, column sums = db)
//
// 256 threads, CTA tile BM x 128 (BM = 128 or 64), k-step 16, thread tile (BM/16) x 8 held as FFMA2
// pairs, operands double-buffered in shared memory, global loads prefetched one k-step ahead.  Lanes
// are arranged as in tt_static.cuh (column tile on lane bits 1..4, row tile on bit 0 + warp) so both
// operand fetches cost the 2-wavefront minimum.  No tensor cores: the 1e-5 parity bar excludes TF32.
#pragma once
#include <cuda_runtime.h>

namespace ttg {

constexpr int NT = 256, BN = 128, BK = 16;

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(f32x2 &c, float a, f32x2 w) {
    f32x2 aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(aa), "l"(w));
}
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Row r of a "ragged" row-major matrix: rows are grouped rpb per batch entry (time-chunk views of
// (B, T, width) tensors): address = p + (r / rpb) * bstride + (r % rpb) * ld
struct Rag {
    const float *p;
    long long bstride;
    int rpb, ld;
};
__device__ __forceinline__ const float *rag_row(const Rag &g, unsigned r) {
    const unsigned b = r / (unsigned)g.rpb;
    return g.p + (long long)b * g.bstride + (long long)(r - b * (unsigned)g.rpb) * g.ld;
}

// one k-step of the register tile: acc[i][jp] += a[i] * b[jp];  NG float4 column groups (64 columns apart) per thread
template <int TM, int NG = 2>
__device__ __forceinline__ void tile_step(const float *__restrict__ As, const float *__restrict__ Bs, int ty, int tx,
                                          f32x2 (&acc)[TM][2 * NG]) {
    constexpr int BM = 16 * TM, BNV = 64 * NG;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
        float a[TM];
        {
            const float4 t = *reinterpret_cast<const float4 *>(As + kk * BM + ty * 4);
            a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
        }
        if constexpr (TM == 8) {
            const float4 t = *reinterpret_cast<const float4 *>(As + kk * BM + 64 + ty * 4);
            a[4] = t.x; a[5] = t.y; a[6] = t.z; a[7] = t.w;
        }
        f32x2 w[2 * NG];
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            const float4 b = *reinterpret_cast<const float4 *>(Bs + kk * BNV + h * 64 + tx * 4);
            w[2 * h] = pk2(b.x, b.y);
            w[2 * h + 1] = pk2(b.z, b.w);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < 2 * NG; ++j) ffma2(acc[i][j], a[i], w[j]);
    }
}

__device__ __forceinline__ void thread_coords(int tid, int &ty, int &tx) {
    const int lane = tid & 31, warp = tid >> 5;
    tx = (lane >> 1) & 15;            // column tile: pair-blocked operand fetch
    ty = warp * 2 + (lane & 1);       // row tile: period-2 operand fetch
}

// ---------------------------------------------------------------------------------------------
// C[r, n] = sum_k A[r, k] B[k, n] (+ bias[n] + bias2[n]);  K % 4 == 0, N % 128 == 0
// ---------------------------------------------------------------------------------------------
struct GemmRowsArgs {
    unsigned rows;
    int K, N;
    Rag a;                   // rows x K
    const float *b;          // K x N dense
    int ldb;
    const float *bias, *bias2;
    Rag c;                   // rows x N (written)
};

// NG = 2: 128-column CTA tile, 8x8 thread tile, two CTAs per SM;  NG = 4: 256-column CTA tile, 8x16 thread tile,
// one CTA per SM (operand floats per FMA drop from 0.25 to 0.19: the shared-memory pipe stops co-limiting)
template <int TM, int NG = 2>
__global__ void __launch_bounds__(NT, NG == 2 ? 2 : 1) k_gemm_rows(const __grid_constant__ GemmRowsArgs g) {
    constexpr int BM = 16 * TM, BNV = 64 * NG;
    constexpr int AV = BM / 64;                  // float4 loads of A per thread per k-step
    constexpr int F4R = BNV / 4, RPP = NT / F4R, NPASS = BK / RPP;   // B loader: float4 per row, rows per pass, passes
    __shared__ __align__(16) float As[2][BK * BM];
    __shared__ __align__(16) float Bs[2][BK * BNV];
    const int tid = threadIdx.x;
    const int tiles_n = g.N / BNV;
    const unsigned tile_m = blockIdx.x / tiles_n;
    const int n0 = (blockIdx.x % tiles_n) * BNV;
    const unsigned r0 = tile_m * BM;
    int ty, tx;
    thread_coords(tid, ty, tx);

    // A loader (transposing): thread owns row am and k-quads akq (+ AV-1 more, 2 apart)
    const int am = tid % BM, akq = tid / BM;     // BM = 128: akq in {0,1} (+2); BM = 64: akq in 0..3
    const bool arow_ok = r0 + am < g.rows;
    const float *arow = arow_ok ? rag_row(g.a, r0 + am) : g.a.p;
    // B loader: rows bk + RPP * v; float4 column bn4
    const int bk = tid / F4R, bn4 = tid % F4R;

    float4 ra[AV], rb[NPASS];
    auto load_g = [&](int k0) {
#pragma unroll
        for (int v = 0; v < AV; ++v) {
            const int k = k0 + (akq + 2 * v) * 4;
            ra[v] = (arow_ok && k < g.K) ? ldg4(arow + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int v = 0; v < NPASS; ++v) {
            const int k = k0 + bk + RPP * v;
            rb[v] = (k < g.K) ? ldg4(g.b + (long long)k * g.ldb + n0 + bn4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_s = [&](int buf) {
#pragma unroll
        for (int v = 0; v < AV; ++v) {
            float *d = As[buf] + (akq + 2 * v) * 4 * BM + am;
            d[0] = ra[v].x; d[BM] = ra[v].y; d[2 * BM] = ra[v].z; d[3 * BM] = ra[v].w;
        }
#pragma unroll
        for (int v = 0; v < NPASS; ++v) *reinterpret_cast<float4 *>(Bs[buf] + (bk + RPP * v) * BNV + bn4 * 4) = rb[v];
    };

    f32x2 acc[TM][2 * NG];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 2 * NG; ++j) acc[i][j] = 0ull;

    load_g(0);
    store_s(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
        const bool more = k0 + BK < g.K;
        if (more) load_g(k0 + BK);
        tile_step<TM, NG>(As[buf], Bs[buf], ty, tx, acc);
        if (more) store_s(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    float4 bv[NG];
#pragma unroll
    for (int h = 0; h < NG; ++h) {
        const int n = n0 + h * 64 + tx * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias) { const float4 t = ldg4(g.bias + n); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        if (g.bias2) { const float4 t = ldg4(g.bias2 + n); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        bv[h] = s;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const unsigned r = r0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (r >= g.rows) continue;
        float *crow = const_cast<float *>(rag_row(g.c, r)) + n0 + tx * 4;
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            float4 v;
            upk2(acc[i][2 * h], v.x, v.y);
            upk2(acc[i][2 * h + 1], v.z, v.w);
            v.x += bv[h].x; v.y += bv[h].y; v.z += bv[h].z; v.w += bv[h].w;
            *reinterpret_cast<float4 *>(crow + h * 64) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Split reduction over rows:  part[s][m, n] = sum_{r in rows of split s} A[r, m] B[r, n]
//                             pbias[s][n]   = sum_r B[r, n]                        (optional)
// M % 4 == 0 (any M <= BM * tiles_m), N % 128 == 0.  grid = tiles_m * tiles_n * nsplit.
// ---------------------------------------------------------------------------------------------
struct GemmRedArgs {
    unsigned rows;
    int M, N, nsplit;
    Rag a;                   // rows x M
    Rag b;                   // rows x N
    float *part;             // [nsplit][M][N]
    float *pbias;            // [nsplit][N] or null
};

template <int TM>
__global__ void __launch_bounds__(NT, 2) k_gemm_red(const __grid_constant__ GemmRedArgs g) {
    constexpr int BM = 16 * TM;
    constexpr int AV = BM / 64;
    constexpr int AF4 = BM / 4;                  // float4 per A row of the tile
    __shared__ __align__(16) float As[2][BK * BM];
    __shared__ __align__(16) float Bs[2][BK * BN];
    const int tid = threadIdx.x;
    const int tiles_n = g.N / BN, tiles_m = (g.M + BM - 1) / BM;
    const int split = blockIdx.x / (tiles_m * tiles_n);
    const int rem = blockIdx.x % (tiles_m * tiles_n);
    const int m0 = (rem / tiles_n) * BM, n0 = (rem % tiles_n) * BN;
    // rows of this split: multiples of BK so that every k-step is full except the last of the matrix
    const unsigned per = (((g.rows + g.nsplit - 1) / g.nsplit) + BK - 1) / BK * BK;
    const unsigned rb = (unsigned)split * per;
    const unsigned re = (rb + per < g.rows) ? rb + per : g.rows;
    int ty, tx;
    thread_coords(tid, ty, tx);

    const int ar = tid / AF4, am4 = tid % AF4;   // A loader: rows ar (+ NT/AF4 per extra load), float4 column am4
    const int br = tid >> 5, bn4 = tid & 31;     // B loader: rows br, br + 8
    const bool am_ok = m0 + am4 * 4 < g.M;

    float4 ra[AV], rbv[2];
    auto load_g = [&](unsigned k0) {
#pragma unroll
        for (int v = 0; v < AV; ++v) {
            const unsigned r = k0 + ar + v * (NT / AF4);
            ra[v] = (am_ok && r < re) ? ldg4(rag_row(g.a, r) + m0 + am4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const unsigned r = k0 + br + 8 * v;
            rbv[v] = (r < re) ? ldg4(rag_row(g.b, r) + n0 + bn4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_s = [&](int buf) {
#pragma unroll
        for (int v = 0; v < AV; ++v)
            *reinterpret_cast<float4 *>(As[buf] + (ar + v * (NT / AF4)) * BM + am4 * 4) = ra[v];
#pragma unroll
        for (int v = 0; v < 2; ++v) *reinterpret_cast<float4 *>(Bs[buf] + (br + 8 * v) * BN + bn4 * 4) = rbv[v];
    };

    f32x2 acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
    // column sums of B (bias gradient): every thread adds up the float4s it stages (rows br, br + 8 of each
    // k-step); the 8 row groups are combined through shared memory at the end.  Only the m0 == 0 tiles do it.
    const bool want_bias = g.pbias && m0 == 0;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);

    if (rb < re) {
        load_g(rb);
        store_s(0);
    }
    __syncthreads();
    int buf = 0;
    for (unsigned k0 = rb; k0 < re; k0 += BK) {
        if (want_bias) {
            bsum.x += rbv[0].x + rbv[1].x; bsum.y += rbv[0].y + rbv[1].y;
            bsum.z += rbv[0].z + rbv[1].z; bsum.w += rbv[0].w + rbv[1].w;
        }
        const bool more = k0 + BK < re;
        if (more) load_g(k0 + BK);
        tile_step<TM>(As[buf], Bs[buf], ty, tx, acc);
        if (more) store_s(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    float *pc = g.part + (long long)split * g.M * g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (m >= g.M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 v;
            upk2(acc[i][2 * h], v.x, v.y);
            upk2(acc[i][2 * h + 1], v.z, v.w);
            *reinterpret_cast<float4 *>(pc + (long long)m * g.N + n0 + h * 64 + tx * 4) = v;
        }
    }
    if (want_bias) {
        float *red = Bs[0];                       // free after the final barrier of the loop
        *reinterpret_cast<float4 *>(red + br * BN + bn4 * 4) = bsum;
        __syncthreads();
        if (tid < 32) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = *reinterpret_cast<const float4 *>(red + q * BN + tid * 4);
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            }
            *reinterpret_cast<float4 *>(g.pbias + (long long)split * g.N + n0 + tid * 4) = t;
        }
    }
}

// out[e] (+)= sum_s part[s * stride + e]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) k_sum_splits(const float *__restrict__ part, int nsplit, long long stride,
                                                     long long n, float *__restrict__ out, int accumulate) {
    for (long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; e < n;
         e += (long long)gridDim.x * blockDim.x * 4) {
        float4 s = accumulate ? *reinterpret_cast<const float4 *>(out + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < nsplit; ++k) {
            const float4 v = *reinterpret_cast<const float4 *>(part + (long long)k * stride + e);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4 *>(out + e) = s;
    }
}

// n x n identity
__global__ void __launch_bounds__(256) k_eye(float *__restrict__ p, int n) {
    const long long tot = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x)
        p[e] = (e / n == e % n) ? 1.0f : 0.0f;
}

// dst (cols x rows) = src (rows x cols)^T
__global__ void __launch_bounds__(256) k_transpose(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;     // 32 x 8
    for (int j = ly; j < 32; j += 8)
        if (by + j < rows && bx + lx < cols) tile[j][lx] = src[(long long)(by + j) * cols + bx + lx];
    __syncthreads();
    for (int j = ly; j < 32; j += 8)
        if (bx + j < cols && by + lx < rows) dst[(long long)(bx + j) * rows + by + lx] = tile[lx][j];
}

}  // namespace ttg
