// Statically specialised TT chain engine (sm_100a): every mode size, rank, row count and thread
// mapping is a compile-time constant, so index arithmetic folds away, the k-loops unroll and the
// gate math fuses into the epilogue of the last contraction stage.  Instantiated for the shapes
// registered in tt_static_inst.cu; every other shape runs on the runtime-shape kernels of
// tt_kernels.cuh (same maths, same C ABI).
//
// Restrictions of the static path (checked by static_assert / by the registry):
//   inner ranks r_1..r_{d-1} multiples of 4, every K_k = j_k * r_{k+1} a multiple of 4,
//   gates aligned with the first output mode (i_0 % G == 0), final-stage tiles <= one per thread.
#pragma once
#include <cuda_runtime.h>
#include "tt_plan.h"

namespace tts {

constexpr int NTHR = 256;

constexpr int cpad(int k) { return (k % 4 != 0) ? k : (((k / 4) % 2 == 1) ? k : k + 4); }
constexpr int cr4(int v) { return (v + 3) & ~3; }
constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int p2floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
// largest power of two <= want that divides n (>= 1)
constexpr int p2div(int n, int want) {
    int p = p2floor(cmax(want, 1));
    while (p > 1 && n % p != 0) p /= 2;
    return p;
}

#define TTS_DEV __device__ __forceinline__

TTS_DEV float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
TTS_DEV void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// ---------------------------------------------------------------------------------------------
// Shape: struct with static constexpr D, G, J[], I[], RK[] (ranks r_0..r_d)
// ---------------------------------------------------------------------------------------------
template <class S, int k>
struct St {
    static_assert(k >= 0 && k < S::D, "stage index");
    static constexpr int J = S::J[k], I = S::I[k], r = S::RK[k], rn = S::RK[k + 1];
    static constexpr int K = J * rn;
    static constexpr int N = I * r;
    static constexpr int mrow() {
        int m = 1;
        for (int q = k + 1; q < S::D; ++q) m *= S::I[q];
        for (int q = 0; q < k; ++q) m *= S::J[q];
        return m;
    }
    static constexpr int Mrow = mrow();
    static constexpr int KS = cpad(K);
    static constexpr int BS = Mrow * KS;              // floats per batch row of X_k
    // shared-memory weight layout: k >= 1: [kappa][n] stride NS;  k == 0: [kappa][i0'][4 gates]
    static constexpr int NW = (k == 0) ? (I / S::G) * 4 : N;
    static constexpr int NS = cpad(NW);
    static constexpr int WFLOATS = cr4(K * NS);
    static constexpr int CORE = r * I * J * rn;       // floats of core k in the blob
};

template <class S> constexpr int n_in() { int v = 1; for (int k = 0; k < S::D; ++k) v *= S::J[k]; return v; }
template <class S> constexpr int n_out() { int v = 1; for (int k = 0; k < S::D; ++k) v *= S::I[k]; return v; }

template <class S, int k> struct WOff { static constexpr int v = WOff<S, k - 1>::v + St<S, k - 1>::WFLOATS; };
template <class S> struct WOff<S, 0> { static constexpr int v = 0; };
template <class S> constexpr int w_floats() { return WOff<S, S::D - 1>::v + St<S, S::D - 1>::WFLOATS; }
template <class S, int k> struct COff { static constexpr int v = COff<S, k - 1>::v + St<S, k - 1>::CORE; };
template <class S> struct COff<S, 0> { static constexpr int v = 0; };
template <class S> constexpr int core_floats() { return COff<S, S::D - 1>::v + St<S, S::D - 1>::CORE; }

// ---- stage the cores: blob (r,i,j,r') -> shared W layouts --------------------------------------
template <class S, int k>
TTS_DEV void stage_weights_k(const float *__restrict__ cores, float *__restrict__ wsm, int tid) {
    using T = St<S, k>;
    constexpr int I0p = T::I / S::G;
    for (int e = tid; e < T::CORE; e += NTHR) {
        const int ap = e % T::rn;
        int t = e / T::rn;
        const int j = t % T::J;
        t /= T::J;
        const int i = t % T::I;
        const int a = t / T::I;
        int col;
        if (k == 0) col = (i % I0p) * 4 + (i / I0p);      // gate index = i / I0p (gates are the high part of i_0)
        else col = i * T::r + a;
        wsm[WOff<S, k>::v + (j * T::rn + ap) * T::NS + col] = __ldg(cores + COff<S, k>::v + e);
    }
    if (k == 0 && S::G == 3) {   // zero the padded 4th gate column so it can be read harmlessly
        for (int e = tid; e < T::K * I0p; e += NTHR) wsm[WOff<S, 0>::v + (e / I0p) * T::NS + (e % I0p) * 4 + 3] = 0.f;
    }
    if constexpr (k + 1 < S::D) stage_weights_k<S, k + 1>(cores, wsm, tid);
}

// ---------------------------------------------------------------------------------------------
// Forward stage k >= 1:  X_{k-1}[(i,m,a)] = sum_kappa X_k[m][kappa] W_k[kappa][(i,a)]
// thread tile = R batch rows x TMr rows x TN columns
// ---------------------------------------------------------------------------------------------
template <class S, int k, int R>
struct FwdMap {
    using T = St<S, k>;
    static constexpr int TN = (T::N % 8 == 0) ? 8 : 4;
    static constexpr int NTl = T::N / TN;
    static constexpr int TMr = p2div(T::Mrow, cmin(cmax(T::Mrow * NTl / NTHR, 1), cmax(64 / (R * TN), 1)));
    static constexpr int MTl = T::Mrow / TMr;
    static constexpr int TT = MTl * NTl;
    static constexpr int ITER = (TT + NTHR - 1) / NTHR;
    static_assert(T::K % 4 == 0 && T::N % 4 == 0 && T::r % 4 == 0, "static path needs K, N, r multiples of 4");
};

template <class S, int k, int R>
TTS_DEV void fwd_stage(const float *__restrict__ X, const float *__restrict__ W, float *__restrict__ Y, int tid) {
    using T = St<S, k>;
    using To = St<S, k - 1>;
    using M = FwdMap<S, k, R>;
    constexpr int TN = M::TN, TMr = M::TMr;
    constexpr int Jp = To::J, KSo = To::KS, BSo = To::BS;
    constexpr int ISo = (T::Mrow / Jp) * KSo;
#pragma unroll 1
    for (int it = 0; it < M::ITER; ++it) {
        const int u = tid + it * NTHR;
        if (M::TT % NTHR != 0 && u >= M::TT) break;
        const int tn = u % M::NTl, mt = u / M::NTl;
        float acc[R][TMr][TN];
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int q = 0; q < TMr; ++q)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[b][q][j] = 0.f;
        const float *xb = X + mt * TMr * T::KS;
        const float *wb = W + tn * TN;
#pragma unroll 4
        for (int k4 = 0; k4 < T::K; k4 += 4) {
            float4 a[R][TMr];
#pragma unroll
            for (int b = 0; b < R; ++b)
#pragma unroll
                for (int q = 0; q < TMr; ++q) a[b][q] = ld4(xb + b * T::BS + q * T::KS + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float w[TN];
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 t = ld4(wb + (k4 + kk) * T::NS + j);
                    w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
                }
#pragma unroll
                for (int b = 0; b < R; ++b)
#pragma unroll
                    for (int q = 0; q < TMr; ++q) {
                        const float av = kk == 0 ? a[b][q].x : (kk == 1 ? a[b][q].y : (kk == 2 ? a[b][q].z : a[b][q].w));
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[b][q][j] = fmaf(av, w[j], acc[b][q][j]);
                    }
            }
        }
#pragma unroll
        for (int q = 0; q < TMr; ++q) {
            const int mr = mt * TMr + q;
            const int base = (mr / Jp) * KSo + (mr % Jp) * T::r;
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const int n = tn * TN + j;
                const int off = base + (n / T::r) * ISo + (n % T::r);
#pragma unroll
                for (int b = 0; b < R; ++b)
                    st4(Y + b * BSo + off, make_float4(acc[b][q][j], acc[b][q][j + 1], acc[b][q][j + 2], acc[b][q][j + 3]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Final stage (k = 0) with every gate of a hidden unit in one thread.
// thread tile = R batch rows x TMr rows x TI first-mode slices x G gates; lanes: 8 row tiles x 4 slices
// hidden unit of (mr, i0') is  h = i0' * Mrow_0 + mr
// ---------------------------------------------------------------------------------------------
template <class S, int R>
struct FinMap {
    using T = St<S, 0>;
    static_assert(T::r == 1 && T::I % S::G == 0, "gates must align with the first output mode");
    static constexpr int I0p = T::I / S::G;
    // grow the tile until one tile per thread suffices
    static constexpr int tiles1 = T::Mrow * I0p;                       // with TMr = TI = 1
    static constexpr int need = (tiles1 + NTHR - 1) / NTHR;            // elements per thread
    static constexpr int TI = cmin(p2div(I0p, need), I0p);
    static constexpr int TMr = p2div(T::Mrow, cmax(need / TI, 1));
    static constexpr int ITl = I0p / TI;
    static constexpr int MTl = T::Mrow / TMr;
    static constexpr int TT = MTl * ITl;
    static_assert(TT <= NTHR, "final stage does not fit one tile per thread");
    static constexpr int LM = (MTl % 8 == 0) ? 8 : p2div(MTl, 8);      // row tiles that are lane-adjacent
    TTS_DEV static void coords(int u, int &mt, int &it) {
        const int lo = u % LM;
        const int rest = u / LM;
        it = rest % ITl;
        mt = (rest / ITl) * LM + lo;
    }
};

template <class S, int R>
TTS_DEV void final_stage(const float *__restrict__ X, const float *__restrict__ W, int mt, int it,
                         float (&acc)[R][FinMap<S, R>::TMr][FinMap<S, R>::TI][4]) {
    using T = St<S, 0>;
    using M = FinMap<S, R>;
    constexpr int TMr = M::TMr, TI = M::TI;
#pragma unroll
    for (int b = 0; b < R; ++b)
#pragma unroll
        for (int q = 0; q < TMr; ++q)
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[b][q][i][g] = 0.f;
    const float *xb = X + mt * TMr * T::KS;
    const float *wb = W + it * TI * 4;
#pragma unroll 4
    for (int k4 = 0; k4 < T::K; k4 += 4) {
        float4 a[R][TMr];
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int q = 0; q < TMr; ++q) a[b][q] = ld4(xb + b * T::BS + q * T::KS + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 w[TI];
#pragma unroll
            for (int i = 0; i < TI; ++i) w[i] = ld4(wb + (k4 + kk) * T::NS + i * 4);
#pragma unroll
            for (int b = 0; b < R; ++b)
#pragma unroll
                for (int q = 0; q < TMr; ++q) {
                    const float av = kk == 0 ? a[b][q].x : (kk == 1 ? a[b][q].y : (kk == 2 ? a[b][q].z : a[b][q].w));
#pragma unroll
                    for (int i = 0; i < TI; ++i) {
                        acc[b][q][i][0] = fmaf(av, w[i].x, acc[b][q][i][0]);
                        acc[b][q][i][1] = fmaf(av, w[i].y, acc[b][q][i][1]);
                        acc[b][q][i][2] = fmaf(av, w[i].z, acc[b][q][i][2]);
                        if (S::G == 4) acc[b][q][i][3] = fmaf(av, w[i].w, acc[b][q][i][3]);
                    }
                }
        }
    }
}

TTS_DEV float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// shared-memory floats of the forward recurrent kernel
template <class S, int R>
struct FwdSmem {
    static constexpr int D = S::D;
    static constexpr int W = w_floats<S>();
    static constexpr int HS = cr4(R * St<S, D - 1>::BS);
    static constexpr int slot(int k) { return 0; }
    // ping-pong slots P (X_{d-2}, X_{d-4}, ..) and Q (X_{d-3}, ..)
    template <int k> static constexpr int bs() { return St<S, k>::BS; }
    static constexpr int pfloats() {
        int m = 0;
        if (D >= 2) m = cmax(m, St<S, (D >= 2 ? D - 2 : 0)>::BS);
        if (D >= 4) m = cmax(m, St<S, (D >= 4 ? D - 4 : 0)>::BS);
        if (D >= 6) m = cmax(m, St<S, (D >= 6 ? D - 6 : 0)>::BS);
        return cr4(R * m);
    }
    static constexpr int qfloats() {
        int m = 0;
        if (D >= 3) m = cmax(m, St<S, (D >= 3 ? D - 3 : 0)>::BS);
        if (D >= 5) m = cmax(m, St<S, (D >= 5 ? D - 5 : 0)>::BS);
        return cr4(R * m);
    }
    static constexpr int P = pfloats(), Q = qfloats();
    static constexpr int TOTAL = W + HS + P + Q;
    static constexpr size_t BYTES = (size_t)TOTAL * 4;
};

struct RnnFwdSArgs {
    int steps;
    long long B;
    const float *xg;         // (B, steps, G*H) ih projection (+ biases folded for LSTM); MODE_XG
    long long xg_bstride;
    const float *x1;         // rank-one input mode: x (B, T, 1) pre-offset to this chunk; row stride x1_bstride
    long long x1_bstride;
    const float *w_eff;      // rank-one input mode: (G*H) dense column of W_ih
    const float *bias_ih;    // rank-one mode: (G*H) or null
    const float *cores;      // hh core blob
    const float *bias_hh;    // (G*H) or null: GRU always; LSTM only in rank-one mode (else folded into xg)
    const float *h_in, *c_in;
    float *out;
    long long out_bstride;
    float *c_save;
    float *h_out, *c_out;
};

enum { MODE_XG = 0, MODE_RANK1 = 1 };

// run stages D-1 .. 1 (each followed by a barrier); returns the slot holding X_0
template <class S, int R, int k>
TTS_DEV const float *fwd_chain_pp(float *hs, float *P, float *Q, const float *wsm, int tid) {
    if constexpr (k == 0) {
        return (S::D == 1) ? hs : (((S::D - 2) % 2 == 0) ? P : Q);
    } else {
        const float *X = (k == S::D - 1) ? hs : (((S::D - 2 - k) % 2 == 0) ? P : Q);
        float *Y = ((S::D - 2 - (k - 1)) % 2 == 0) ? P : Q;
        fwd_stage<S, k, R>(X, wsm + WOff<S, k>::v, Y, tid);
        __syncthreads();
        return fwd_chain_pp<S, R, k - 1>(hs, P, Q, wsm, tid);
    }
}

template <class S, int CELL, int R, int MODE>
__global__ void __launch_bounds__(NTHR, 1) k_rnn_fwd_s(const __grid_constant__ RnnFwdSArgs a) {
    extern __shared__ __align__(16) float smem[];
    using SM = FwdSmem<S, R>;
    using FM = FinMap<S, R>;
    using TL = St<S, S::D - 1>;
    using T0 = St<S, 0>;
    constexpr int G = S::G, H = n_in<S>(), GH = G * H;
    constexpr int TMr = FM::TMr, TI = FM::TI;
    constexpr bool LSTM = (CELL == TTRNN_CELL_LSTM);
    static_assert(G == (LSTM ? 4 : 3), "gate count");
    const int tid = threadIdx.x;
    float *wsm = smem;
    float *hs = wsm + SM::W;
    float *P = hs + SM::HS;
    float *Q = P + SM::P;

    stage_weights_k<S, 0>(a.cores, wsm, tid);

    const bool active = tid < FM::TT;
    int mt = 0, it = 0;
    FM::coords(active ? tid : 0, mt, it);
    // hidden units owned by this thread: h(q, i) = (it*TI + i) * Mrow_0 + mt*TMr + q
    float bhh[TI][TMr][4], weff[TI][TMr][4], bih[TI][TMr][4];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int q = 0; q < TMr; ++q)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int col = g * H + (it * TI + i) * T0::Mrow + mt * TMr + q;
                bhh[i][q][g] = a.bias_hh ? __ldg(a.bias_hh + col) : 0.f;
                weff[i][q][g] = (MODE == MODE_RANK1) ? __ldg(a.w_eff + col) : 0.f;
                bih[i][q][g] = (MODE == MODE_RANK1 && a.bias_ih) ? __ldg(a.bias_ih + col) : 0.f;
            }

    const long long ntiles = (a.B + R - 1) / R;
    for (long long tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const long long row0 = tile_i * R;
        __syncthreads();
        // initial state: h -> shared slot (X_{d-1} layout), c / h_prev -> registers of the owner thread
        for (int e = tid; e < R * H; e += NTHR) {
            const int b = e / H, h = e % H;
            float hv = 0.f;
            if (row0 + b < a.B && a.h_in) hv = __ldg(a.h_in + (row0 + b) * H + h);
            hs[b * TL::BS + (h / TL::K) * TL::KS + (h % TL::K)] = hv;
        }
        float cst[R][TMr][TI], hpr[R][TMr][TI];
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int q = 0; q < TMr; ++q)
#pragma unroll
                for (int i = 0; i < TI; ++i) {
                    const int h = (it * TI + i) * T0::Mrow + mt * TMr + q;
                    const bool ok = active && (row0 + b < a.B);
                    cst[b][q][i] = (LSTM && ok && a.c_in) ? __ldg(a.c_in + (row0 + b) * H + h) : 0.f;
                    hpr[b][q][i] = (ok && a.h_in) ? __ldg(a.h_in + (row0 + b) * H + h) : 0.f;
                }
        __syncthreads();

        for (int t = 0; t < a.steps; ++t) {
            // ---- operands of the gate phase, requested before the chain so their latency hides
            float xin[R][TMr][TI][4];
            float x1[R];
            if (MODE == MODE_XG) {
#pragma unroll
                for (int b = 0; b < R; ++b)
#pragma unroll
                    for (int q = 0; q < TMr; ++q)
#pragma unroll
                        for (int i = 0; i < TI; ++i)
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                const int col = g * H + (it * TI + i) * T0::Mrow + mt * TMr + q;
                                xin[b][q][i][g] = (active && row0 + b < a.B)
                                    ? __ldg(a.xg + (row0 + b) * a.xg_bstride + (long long)t * GH + col) : 0.f;
                            }
            } else {
#pragma unroll
                for (int b = 0; b < R; ++b)
                    x1[b] = (row0 + b < a.B) ? __ldg(a.x1 + (row0 + b) * a.x1_bstride + t) : 0.f;
            }
            // ---- stages d-1 .. 1
            const float *X0 = fwd_chain_pp<S, R, S::D - 1>(hs, P, Q, wsm, tid);
            // ---- stage 0 fused with the gate math and the state update
            if (active) {
                float acc[R][TMr][TI][4];
                final_stage<S, R>(X0, wsm + WOff<S, 0>::v, mt, it, acc);
#pragma unroll
                for (int b = 0; b < R; ++b)
#pragma unroll
                    for (int q = 0; q < TMr; ++q)
#pragma unroll
                        for (int i = 0; i < TI; ++i) {
                            const int h = (it * TI + i) * T0::Mrow + mt * TMr + q;
                            float ain[4];
#pragma unroll
                            for (int g = 0; g < G; ++g)
                                ain[g] = (MODE == MODE_XG) ? xin[b][q][i][g] : fmaf(x1[b], weff[i][q][g], bih[i][q][g]);
                            float hnew;
                            if (LSTM) {
                                const float ig = sigmoidf_acc(acc[b][q][i][0] + bhh[i][q][0] + ain[0]);
                                const float fg = sigmoidf_acc(acc[b][q][i][1] + bhh[i][q][1] + ain[1]);
                                const float gg = tanhf(acc[b][q][i][2] + bhh[i][q][2] + ain[2]);
                                const float og = sigmoidf_acc(acc[b][q][i][3] + bhh[i][q][3] + ain[3]);
                                const float cn = fg * cst[b][q][i] + ig * gg;
                                cst[b][q][i] = cn;
                                hnew = og * tanhf(cn);
                            } else {
                                const float rg = sigmoidf_acc(ain[0] + (acc[b][q][i][0] + bhh[i][q][0]));
                                const float zg = sigmoidf_acc(ain[1] + (acc[b][q][i][1] + bhh[i][q][1]));
                                const float ng = tanhf(ain[2] + rg * (acc[b][q][i][2] + bhh[i][q][2]));
                                hnew = (1.0f - zg) * ng + zg * hpr[b][q][i];
                            }
                            hpr[b][q][i] = hnew;
                            hs[b * TL::BS + (h / TL::K) * TL::KS + (h % TL::K)] = hnew;
                            if (row0 + b < a.B) {
                                a.out[(row0 + b) * a.out_bstride + (long long)t * H + h] = hnew;
                                if (LSTM && a.c_save) a.c_save[(row0 + b) * a.out_bstride + (long long)t * H + h] = cst[b][q][i];
                            }
                        }
            }
            __syncthreads();
        }
        if (active) {
#pragma unroll
            for (int b = 0; b < R; ++b)
#pragma unroll
                for (int q = 0; q < TMr; ++q)
#pragma unroll
                    for (int i = 0; i < TI; ++i) {
                        const int h = (it * TI + i) * T0::Mrow + mt * TMr + q;
                        if (row0 + b < a.B) {
                            if (a.h_out) a.h_out[(row0 + b) * H + h] = hpr[b][q][i];
                            if (LSTM && a.c_out) a.c_out[(row0 + b) * H + h] = cst[b][q][i];
                        }
                    }
        }
    }
}

}  // namespace tts
