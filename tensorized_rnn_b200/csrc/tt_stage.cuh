// Device building blocks: the per-core GEMM stages of a TT chain, operating on
// shared-memory tiles, FP32 FFMA with register micro-tiles.
//
// All shape arguments are plain ints and every function is force-inlined, so a
// caller that passes compile-time constants gets a fully specialised stage and
// a caller that passes a runtime plan gets the generic one from the same source.
#pragma once
#include "tt_plan.h"

#define TT_DEV __device__ __forceinline__

TT_DEV float4 tt_ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
TT_DEV void tt_st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
TT_DEV float tt_get(const float4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// tile codes chosen on the host (StagePlan has no room for templates)
enum { TT_TILE_GEN = 0, TT_TILE_1x4, TT_TILE_2x4, TT_TILE_4x4, TT_TILE_8x4, TT_TILE_4x8, TT_TILE_8x8 };

// ---------------------------------------------------------------------------
// Core blob (r_k, i_k, j_k, r_{k+1}) -> shared-memory W_k[kappa=(j,a')][n=(i,a)]
// ---------------------------------------------------------------------------
TT_DEV void tt_stage_weights(const ChainPlan &p, const float *__restrict__ cores, float *__restrict__ wsm,
                             int tid, int nthr) {
    for (int k = 0; k < p.d; ++k) {
        const StagePlan &s = p.st[k];
        const int total = s.r * s.I * s.J * s.rn;
        for (int e = tid; e < total; e += nthr) {
            int ap = e % s.rn;
            int t = e / s.rn;
            int j = t % s.J;
            t /= s.J;
            int i = t % s.I;
            int a = t / s.I;
            wsm[s.w_off + (j * s.rn + ap) * s.NS + i * s.r + a] = __ldg(cores + s.c_off + e);
        }
    }
}

// Address of output element (row m of the stage, column n = (i, a)) inside X_{k-1}:
//   b*BSo + i*ISo + (mr / Jp)*KSo + (mr % Jp)*r + a,   m = b*Mrow + mr
TT_DEV int tt_out_row_base(const StagePlan &s, int m) {
    const int b = m / s.Mrow;
    const int mr = m - b * s.Mrow;
    const int hi = mr / s.Jp;
    const int lo = mr - hi * s.Jp;
    return b * s.BSo + hi * s.KSo + lo * s.r;
}

// ---------------------------------------------------------------------------
// Forward stage, vector path.  Requires K % 4 == 0, N % TN == 0, TN % 4 == 0.
//   Y[(i,m,a)] = sum_kappa X[m][kappa] * W[kappa][(i,a)]
// Thread (tm, tn): n-tile tn (TN consecutive columns), rows m0 + q*MTH.
// ---------------------------------------------------------------------------
template <int TM, int TN>
TT_DEV void tt_stage_fwd_vec(const StagePlan &s, int R, const float *__restrict__ X,
                             const float *__restrict__ W, float *__restrict__ Y, int tid, int nthr) {
    const int M = R * s.Mrow;
    const int NT = s.N / TN;
    const int NTt = NT < nthr ? NT : nthr;
    const int MTH = nthr / NTt;
    const int tn = tid % NTt, tm = tid / NTt;
    if (tm >= MTH) return;
    const bool vec_out = (s.r % 4 == 0);
    for (int nb = tn; nb < NT; nb += NTt) {
        const int n0 = nb * TN;
        const float *wp = W + n0;
        for (int m0 = tm; m0 < M; m0 += MTH * TM) {
            float acc[TM][TN];
            const float *xr[TM];
#pragma unroll
            for (int q = 0; q < TM; ++q) {
                int m = m0 + q * MTH;
                xr[q] = X + (m < M ? m : m0) * s.KS;
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[q][j] = 0.f;
            }
#pragma unroll 2
            for (int k = 0; k < s.K; k += 4) {
                float4 a[TM];
#pragma unroll
                for (int q = 0; q < TM; ++q) a[q] = tt_ld4(xr[q] + k);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float w[TN];
#pragma unroll
                    for (int j = 0; j < TN; j += 4) {
                        float4 t = tt_ld4(wp + (k + kk) * s.NS + j);
                        w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
                    }
#pragma unroll
                    for (int q = 0; q < TM; ++q) {
                        const float av = tt_get(a[q], kk);
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[q][j] = fmaf(av, w[j], acc[q][j]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < TM; ++q) {
                const int m = m0 + q * MTH;
                if (m < M) {
                    float *yb = Y + tt_out_row_base(s, m);
                    if (vec_out) {
#pragma unroll
                        for (int j = 0; j < TN; j += 4) {
                            const int n = n0 + j;
                            const int i = n / s.r;
                            const int a0 = n - i * s.r;
                            tt_st4(yb + i * s.ISo + a0, make_float4(acc[q][j], acc[q][j + 1], acc[q][j + 2], acc[q][j + 3]));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < TN; ++j) {
                            const int n = n0 + j;
                            const int i = n / s.r;
                            const int a0 = n - i * s.r;
                            yb[i * s.ISo + a0] = acc[q][j];
                        }
                    }
                }
            }
        }
    }
}

// Forward stage, generic scalar path: any K, N (bounds-checked).
template <int TM, int TN>
TT_DEV void tt_stage_fwd_gen(const StagePlan &s, int R, const float *__restrict__ X,
                             const float *__restrict__ W, float *__restrict__ Y, int tid, int nthr) {
    const int M = R * s.Mrow;
    const int NT = (s.N + TN - 1) / TN;
    const int NTt = NT < nthr ? NT : nthr;
    const int MTH = nthr / NTt;
    const int tn = tid % NTt, tm = tid / NTt;
    if (tm >= MTH) return;
    for (int nb = tn; nb < NT; nb += NTt) {
        const int n0 = nb * TN;
        for (int m0 = tm; m0 < M; m0 += MTH * TM) {
            float acc[TM][TN];
            const float *xr[TM];
#pragma unroll
            for (int q = 0; q < TM; ++q) {
                int m = m0 + q * MTH;
                xr[q] = X + (m < M ? m : m0) * s.KS;
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[q][j] = 0.f;
            }
            for (int k = 0; k < s.K; ++k) {
                float w[TN];
#pragma unroll
                for (int j = 0; j < TN; ++j) w[j] = (n0 + j < s.N) ? W[k * s.NS + n0 + j] : 0.f;
#pragma unroll
                for (int q = 0; q < TM; ++q) {
                    const float av = xr[q][k];
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[q][j] = fmaf(av, w[j], acc[q][j]);
                }
            }
#pragma unroll
            for (int q = 0; q < TM; ++q) {
                const int m = m0 + q * MTH;
                if (m < M) {
                    float *yb = Y + tt_out_row_base(s, m);
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        const int n = n0 + j;
                        if (n < s.N) {
                            const int i = n / s.r;
                            const int a0 = n - i * s.r;
                            yb[i * s.ISo + a0] = acc[q][j];
                        }
                    }
                }
            }
        }
    }
}

TT_DEV void tt_stage_fwd(const StagePlan &s, int tile, int R, const float *X, const float *W, float *Y,
                         int tid, int nthr) {
    switch (tile) {
    case TT_TILE_8x8: tt_stage_fwd_vec<8, 8>(s, R, X, W, Y, tid, nthr); break;
    case TT_TILE_4x8: tt_stage_fwd_vec<4, 8>(s, R, X, W, Y, tid, nthr); break;
    case TT_TILE_8x4: tt_stage_fwd_vec<8, 4>(s, R, X, W, Y, tid, nthr); break;
    case TT_TILE_4x4: tt_stage_fwd_vec<4, 4>(s, R, X, W, Y, tid, nthr); break;
    case TT_TILE_2x4: tt_stage_fwd_vec<2, 4>(s, R, X, W, Y, tid, nthr); break;
    case TT_TILE_1x4: tt_stage_fwd_vec<1, 4>(s, R, X, W, Y, tid, nthr); break;
    default: tt_stage_fwd_gen<2, 2>(s, R, X, W, Y, tid, nthr); break;
    }
}

// ---------------------------------------------------------------------------
// Backward-data stage:  dX[m][kappa] = sum_n dY[(i,m,a)] * W[kappa][(i,a)]
// dY is read from the X_{k-1}-shaped buffer the forward stage wrote into.
// Thread (tm, tk): kappa-tile of TK consecutive kappas, rows m0 + q*MTH.
// Vector path requires K % TK == 0, TK % 4 == 0 (stores) and r % 4 == 0 or r == 1.
// ---------------------------------------------------------------------------
template <int TM, int TK, bool VEC>
TT_DEV void tt_stage_bwd_data(const StagePlan &s, int R, const float *__restrict__ dY,
                              const float *__restrict__ W, float *__restrict__ dX, int tid, int nthr) {
    const int M = R * s.Mrow;
    const int KT = (s.K + TK - 1) / TK;
    const int KTt = KT < nthr ? KT : nthr;
    const int MTH = nthr / KTt;
    const int tk = tid % KTt, tm = tid / KTt;
    if (tm >= MTH) return;
    for (int kb = tk; kb < KT; kb += KTt) {
        const int k0 = kb * TK;
        for (int m0 = tm; m0 < M; m0 += MTH * TM) {
            float acc[TM][TK];
            const float *yr[TM];
#pragma unroll
            for (int q = 0; q < TM; ++q) {
                int m = m0 + q * MTH;
                yr[q] = dY + tt_out_row_base(s, m < M ? m : m0);
#pragma unroll
                for (int j = 0; j < TK; ++j) acc[q][j] = 0.f;
            }
            if (VEC && (s.r % 4 == 0)) {
                // n = (i, a): four consecutive a per load
                for (int i = 0; i < s.I; ++i) {
                    for (int a0 = 0; a0 < s.r; a0 += 4) {
                        const int n = i * s.r + a0;
                        float4 y[TM];
#pragma unroll
                        for (int q = 0; q < TM; ++q) y[q] = tt_ld4(yr[q] + i * s.ISo + a0);
#pragma unroll
                        for (int j = 0; j < TK; ++j) {
                            const float4 w = tt_ld4(W + (k0 + j) * s.NS + n);
#pragma unroll
                            for (int q = 0; q < TM; ++q) {
                                acc[q][j] = fmaf(y[q].x, w.x, acc[q][j]);
                                acc[q][j] = fmaf(y[q].y, w.y, acc[q][j]);
                                acc[q][j] = fmaf(y[q].z, w.z, acc[q][j]);
                                acc[q][j] = fmaf(y[q].w, w.w, acc[q][j]);
                            }
                        }
                    }
                }
            } else {
                for (int i = 0; i < s.I; ++i) {
                    for (int a0 = 0; a0 < s.r; ++a0) {
                        const int n = i * s.r + a0;
                        float y[TM];
#pragma unroll
                        for (int q = 0; q < TM; ++q) y[q] = yr[q][i * s.ISo + a0];
#pragma unroll
                        for (int j = 0; j < TK; ++j) {
                            const float w = (k0 + j < s.K) ? W[(k0 + j) * s.NS + n] : 0.f;
#pragma unroll
                            for (int q = 0; q < TM; ++q) acc[q][j] = fmaf(y[q], w, acc[q][j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < TM; ++q) {
                const int m = m0 + q * MTH;
                if (m < M) {
                    float *xo = dX + m * s.KS + k0;
                    if (VEC) {
#pragma unroll
                        for (int j = 0; j < TK; j += 4)
                            tt_st4(xo + j, make_float4(acc[q][j], acc[q][j + 1], acc[q][j + 2], acc[q][j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < TK; ++j)
                            if (k0 + j < s.K) xo[j] = acc[q][j];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Backward-weight stage:  dW[kappa][n] += sum_m X[m][kappa] * dY[(i,m,a)]
// accumulated into a shared-memory copy of the W layout (dWs, same strides).
// Work item = (m-group, kappa-tile, n-tile); each item reduces its rows in
// registers, then adds its TKxTN partial into shared memory.  When several m-groups share an
// entry (mgroups > 1) the groups add one after the other between block barriers: a fixed summation
// order, so the gradients are bit-reproducible (no atomics).  Must be called by every thread of the CTA.
//   VX: K % 4 == 0  -> X read as float4 (TK = 4), else TK = 2 scalar
//   VY: r % 4 == 0  -> dY read as float4 along a (TN = 4), else TN scalar loads
// ---------------------------------------------------------------------------
template <bool VX, bool VY>
TT_DEV void tt_stage_bwd_weight(const StagePlan &s, int R, const float *__restrict__ X,
                                const float *__restrict__ dY, float *__restrict__ dWs, int mgroups,
                                int tid, int nthr) {
    constexpr int TK = VX ? 4 : 2;
    constexpr int TN = (VX || VY) ? 4 : 2;
    const int M = R * s.Mrow;
    const int KT = (s.K + TK - 1) / TK;
    const int NT = (s.N + TN - 1) / TN;
    const int items = KT * NT * mgroups;
    const int rows_per_group = (M + mgroups - 1) / mgroups;
    // uniform trip count: every thread walks the same number of item rounds so that the serialised
    // group adds below can sit between block barriers
    for (int it0 = 0; it0 < items; it0 += nthr) {
        const int it = it0 + tid;
        const bool has = it < items;
        const int tn = it % NT;
        int t = it / NT;
        const int tk = t % KT;
        const int g = t / KT;
        const int k0 = tk * TK, n0 = tn * TN;
        const int mbeg = g * rows_per_group;
        const int mend = (mbeg + rows_per_group < M) ? mbeg + rows_per_group : M;
        float acc[TK][TN];
#pragma unroll
        for (int a = 0; a < TK; ++a)
#pragma unroll
            for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
        // column offsets of this n-tile inside one output row group
        int ycol[TN];
#pragma unroll
        for (int b = 0; b < TN; ++b) {
            const int n = (n0 + b < s.N) ? n0 + b : n0;
            const int i = n / s.r;
            ycol[b] = i * s.ISo + (n - i * s.r);
        }
        int bb = mbeg / s.Mrow;
        int mr = mbeg - bb * s.Mrow;
        for (int m = mbeg; has && m < mend; ++m) {
            const int hi = mr / s.Jp;
            const int lo = mr - hi * s.Jp;
            const float *yp = dY + bb * s.BSo + hi * s.KSo + lo * s.r;
            float x[TK], y[TN];
            if (VX) {
                const float4 xv = tt_ld4(X + m * s.KS + k0);
                x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
            } else {
#pragma unroll
                for (int a = 0; a < TK; ++a) x[a] = (k0 + a < s.K) ? X[m * s.KS + k0 + a] : 0.f;
            }
            if (VY) {
                const float4 yv = tt_ld4(yp + ycol[0]);
                y[0] = yv.x; y[1] = yv.y; y[2] = yv.z; y[3] = yv.w;
            } else {
#pragma unroll
                for (int b = 0; b < TN; ++b) y[b] = (n0 + b < s.N) ? yp[ycol[b]] : 0.f;
            }
#pragma unroll
            for (int a = 0; a < TK; ++a)
#pragma unroll
                for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(x[a], y[b], acc[a][b]);
            if (++mr == s.Mrow) { mr = 0; ++bb; }
        }
        if (mgroups == 1) {
            if (has) {
#pragma unroll
                for (int a = 0; a < TK; ++a)
#pragma unroll
                    for (int b = 0; b < TN; ++b)
                        if (k0 + a < s.K && n0 + b < s.N) dWs[(k0 + a) * s.NS + n0 + b] += acc[a][b];
            }
        } else {
            // the m-groups of one (kappa, n) tile add one after the other: fixed summation order
            for (int gg = 0; gg < mgroups; ++gg) {
                if (has && g == gg) {
#pragma unroll
                    for (int a = 0; a < TK; ++a)
#pragma unroll
                        for (int b = 0; b < TN; ++b)
                            if (k0 + a < s.K && n0 + b < s.N) dWs[(k0 + a) * s.NS + n0 + b] += acc[a][b];
                }
                __syncthreads();
            }
        }
    }
}

// Shared-memory dW (W layout) -> gradient blob layout (r_k, i_k, j_k, r_{k+1}); plain store
TT_DEV void tt_unstage_weights(const ChainPlan &p, const float *__restrict__ dws, float *__restrict__ dcores,
                               int tid, int nthr) {
    for (int k = 0; k < p.d; ++k) {
        const StagePlan &s = p.st[k];
        const int total = s.r * s.I * s.J * s.rn;
        for (int e = tid; e < total; e += nthr) {
            int ap = e % s.rn;
            int t = e / s.rn;
            int j = t % s.J;
            t /= s.J;
            int i = t % s.I;
            int a = t / s.I;
            dcores[s.c_off + e] = dws[s.w_off + (j * s.rn + ap) * s.NS + i * s.r + a];
        }
    }
}

// accurate logistic; expf / tanhf are the full-accuracy libdevice versions (no fast-math)
TT_DEV float tt_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
