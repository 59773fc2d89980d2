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
// Lane arrangement.  Measured on B200 (tools/lds_probe.cu): an LDS.128 costs 2 wavefronts when the
// lanes that share an address are adjacent pairs (address = f(lane/2)) or when the address
// depends only on lane%2 and on the high lane bits; period-4/8/16 interleaved duplicates cost 4.
// So a warp is laid out as LX "x" positions (lane bits 1..lgLX) times LY = 32/LX "y" positions
// (lane bit 0 and the top bits): the x operand is read pair-blocked, the y operand with period 2.
// ---------------------------------------------------------------------------------------------
template <int LX>
struct Lanes {
    static constexpr int LY = 32 / LX;
    static constexpr int lg = (LX == 32) ? 5 : (LX == 16) ? 4 : (LX == 8) ? 3 : (LX == 4) ? 2 : (LX == 2) ? 1 : 0;
    TTS_DEV static int x(int lane) { return (LX == 32) ? lane : ((lane >> 1) & (LX - 1)); }
    TTS_DEV static int y(int lane) { return (LX == 32) ? 0 : ((lane & 1) | ((lane >> (1 + lg)) << 1)); }
};

// ---------------------------------------------------------------------------------------------
// Forward stage k >= 1:  X_{k-1}[(i,m,a)] = sum_kappa X_k[m][kappa] W_k[kappa][(i,a)]
// thread tile = (R batch rows x TMr rows) x TN columns; TMr rows are MTl apart (lanes read
// neighbouring rows), the TN = 8 columns are two float4 groups N/2 apart (lanes read neighbouring
// float4s of W).  x = column tile, y = row tile.
// ---------------------------------------------------------------------------------------------
template <class S, int k, int R, int TMr_, int TN_>
struct FwdMap {
    using T = St<S, k>;
    static constexpr int TN = TN_, TMr = TMr_;
    static_assert(TN == 4 || TN == 8, "TN");
    static_assert(T::K % 4 == 0 && T::N % TN == 0 && T::r % 4 == 0, "static path needs K, N, r multiples of 4");
    static_assert(T::Mrow % TMr == 0, "TMr must divide Mrow");
    static constexpr int NTl = T::N / TN;
    static constexpr int MTl = T::Mrow / TMr;
    static constexpr int LX = NTl >= 16 ? 16 : NTl;            // column tiles across a warp
    static constexpr int LY = 32 / LX;
    static_assert(NTl % LX == 0 && MTl % LY == 0, "tile grid must be a multiple of the lane grid");
    static constexpr int WX = NTl / LX;                         // warps along columns
    static constexpr int TT = MTl * NTl;
    static constexpr int ITER = (TT + NTHR - 1) / NTHR;
    static_assert(TT % NTHR == 0 || TT < NTHR, "tile count");
};

template <class S, int k, int R, int TMr, int TN>
TTS_DEV void fwd_stage(const float *__restrict__ X, const float *__restrict__ W, float *__restrict__ Y, int tid) {
    using T = St<S, k>;
    using To = St<S, k - 1>;
    using M = FwdMap<S, k, R, TMr, TN>;
    using L = Lanes<M::LX>;
    constexpr int Jp = To::J, KSo = To::KS, BSo = To::BS;
    constexpr int ISo = (T::Mrow / Jp) * KSo;
    constexpr int NG = TN / 4;                 // float4 column groups per thread
    constexpr int GSTR = T::N / NG;            // distance between the groups (columns)
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int it = 0; it < M::ITER; ++it) {
        const int wv = warp + it * (NTHR / 32);
        const int tn = (wv % M::WX) * M::LX + L::x(lane);
        const int mt = (wv / M::WX) * M::LY + L::y(lane);
        if (M::TT < NTHR && mt >= M::MTl) break;
        float acc[R][TMr][TN];
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int q = 0; q < TMr; ++q)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[b][q][j] = 0.f;
        const float *xb = X + mt * T::KS;
        const float *wb = W + tn * 4;
#pragma unroll 2
        for (int k4 = 0; k4 < T::K; k4 += 4) {
            float4 a[R][TMr];
#pragma unroll
            for (int b = 0; b < R; ++b)
#pragma unroll
                for (int q = 0; q < TMr; ++q) a[b][q] = ld4(xb + b * T::BS + q * M::MTl * T::KS + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float w[TN];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float4 t = ld4(wb + (k4 + kk) * T::NS + g * GSTR);
                    w[4 * g] = t.x; w[4 * g + 1] = t.y; w[4 * g + 2] = t.z; w[4 * g + 3] = t.w;
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
            const int mr = q * M::MTl + mt;
            const int base = (mr / Jp) * KSo + (mr % Jp) * T::r;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int n = g * GSTR + tn * 4;
                const int off = base + (n / T::r) * ISo + (n % T::r);
#pragma unroll
                for (int b = 0; b < R; ++b)
                    st4(Y + b * BSo + off,
                        make_float4(acc[b][q][4 * g], acc[b][q][4 * g + 1], acc[b][q][4 * g + 2], acc[b][q][4 * g + 3]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Final stage (k = 0): every gate of a hidden unit ends up in one thread.
//   thread tile  = (R batch rows x TMr rows) x (TI first-mode slices x G gates), over 1/SK of K
//   thread grid  = MTl row tiles x ITl slice tiles x SK k-splits = NTHR threads
//   lanes        = 16 row tiles (x) x 2 slice tiles (y)   [ITl >= 2]   or 32 row tiles [ITl == 1]
//   after the k-loop the SK partial tiles are reduce-scattered through shared memory: thread kh
//   keeps the elements e = q*TI + i with e % SK == kh (NE = TMr*TI/SK of them).
// hidden unit of (mr, i0') is  h = i0' * Mrow_0 + mr
// ---------------------------------------------------------------------------------------------
template <class S, int R, int TMr_, int TI_, int SK_>
struct FinMap {
    using T = St<S, 0>;
    static_assert(T::r == 1 && T::I % S::G == 0, "gates must align with the first output mode");
    static constexpr int TMr = TMr_, TI = TI_, SK = SK_;
    static constexpr int I0p = T::I / S::G;
    static_assert(I0p % TI == 0 && T::Mrow % TMr == 0 && (T::K / 4) % SK == 0, "final tile shape");
    static_assert((TMr * TI) % SK == 0, "k-split must divide the tile elements");
    static constexpr int ITl = I0p / TI;
    static constexpr int MTl = T::Mrow / TMr;
    static constexpr int TT = MTl * ITl * SK;
    static_assert(TT == NTHR, "final stage must use exactly one tile per thread");
    static constexpr int NE = TMr * TI / SK;            // elements (hidden units per batch row) owned per thread
    static constexpr int LX = (ITl >= 2) ? 16 : 32;
    static constexpr int LYI = (ITl >= 2) ? 2 : 1;      // slice tiles per warp
    static_assert(MTl % LX == 0 && ITl % LYI == 0, "lane grid");
    static constexpr int WM = MTl / LX;                 // warps along rows
    static constexpr int WI = ITl / LYI;                // warps along slices
    static constexpr int KPART = T::K / SK;
    static constexpr int XCH = (SK > 1) ? cr4((SK - 1) * NE * R * 4) : 0;   // exchange floats per thread
    static constexpr int XCH_FLOATS = XCH * NTHR;
    TTS_DEV static void coords(int tid, int &mt, int &itg, int &kh) {
        const int lane = tid & 31, warp = tid >> 5;
        using L = Lanes<LX>;
        mt = (warp % WM) * LX + L::x(lane);
        itg = ((warp / WM) % WI) * LYI + L::y(lane);
        kh = warp / (WM * WI);
    }
    // element e = q*TI + i  ->  hidden unit
    TTS_DEV static int hidden(int mt, int itg, int e) {
        const int q = e / TI, i = e % TI;
        return (itg * TI + i) * T::Mrow + q * MTl + mt;
    }
};

// accumulates the partial tile of this thread; acc[b][q][i][g]
template <class S, int R, class FM>
TTS_DEV void final_partial(const float *__restrict__ X, const float *__restrict__ W, int mt, int itg, int kh,
                           float (&acc)[R][FM::TMr][FM::TI][4]) {
    using T = St<S, 0>;
    constexpr int TMr = FM::TMr, TI = FM::TI;
#pragma unroll
    for (int b = 0; b < R; ++b)
#pragma unroll
        for (int q = 0; q < TMr; ++q)
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[b][q][i][g] = 0.f;
    const float *xb = X + mt * T::KS + kh * FM::KPART;
    const float *wb = W + kh * FM::KPART * T::NS + itg * TI * 4;
#pragma unroll 2
    for (int k4 = 0; k4 < FM::KPART; k4 += 4) {
        float4 a[R][TMr];
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int q = 0; q < TMr; ++q) a[b][q] = ld4(xb + b * T::BS + q * FM::MTl * T::KS + k4);
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

// reduce-scatter of the SK partial tiles: on return pre[b][n][g] holds the full sums of the NE
// elements this thread owns (element index e = n*SK + kh).  xch: FM::XCH_FLOATS floats of shared
// memory; contains two block barriers when SK > 1.
template <class S, int R, class FM>
TTS_DEV void final_reduce(float (&acc)[R][FM::TMr][FM::TI][4], float (&pre)[R][FM::NE][4], float *xch, int tid, int kh) {
    constexpr int TI = FM::TI, SK = FM::SK, NE = FM::NE;
    constexpr int PER = NTHR / SK;                       // threads per k-split group
    if constexpr (SK > 1) {
        // slot layout: [(d-1)][n][b] float4 per thread, thread-minor: xch4[slot * NTHR + owner_tid]
        float4 *x4 = reinterpret_cast<float4 *>(xch);
        const int tprime = tid % PER;
#pragma unroll
        for (int e = 0; e < FM::TMr * TI; ++e) {
            const int owner = e % SK;                    // compile-time after unrolling
            const int n = e / SK;
            const int q = e / TI, i = e % TI;
#pragma unroll
            for (int dlt = 1; dlt < SK; ++dlt) {
                // this thread is the dlt-th partner of `owner` iff (kh - owner) mod SK == dlt
                if (((kh - owner + SK) % SK) == dlt) {
#pragma unroll
                    for (int b = 0; b < R; ++b)
                        x4[((dlt - 1) * NE * R + n * R + b) * NTHR + owner * PER + tprime] =
                            make_float4(acc[b][q][i][0], acc[b][q][i][1], acc[b][q][i][2], acc[b][q][i][3]);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < FM::TMr * TI; ++e) {
            const int owner = e % SK;
            const int n = e / SK;
            const int q = e / TI, i = e % TI;
            if (owner == kh) {
#pragma unroll
                for (int b = 0; b < R; ++b) {
                    float s0 = acc[b][q][i][0], s1 = acc[b][q][i][1], s2 = acc[b][q][i][2], s3 = acc[b][q][i][3];
#pragma unroll
                    for (int dlt = 1; dlt < SK; ++dlt) {
                        const float4 v = x4[((dlt - 1) * NE * R + n * R + b) * NTHR + tid];
                        s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
                    }
                    pre[b][n][0] = s0; pre[b][n][1] = s1; pre[b][n][2] = s2; pre[b][n][3] = s3;
                }
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < FM::TMr * TI; ++e) {
            const int q = e / TI, i = e % TI;
#pragma unroll
            for (int b = 0; b < R; ++b)
#pragma unroll
                for (int g = 0; g < 4; ++g) pre[b][e][g] = acc[b][q][i][g];
        }
    }
}

TTS_DEV float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- per (shape, R) tuning table ----------------------------------------------------------------
// TMr[k], TN[k] for the forward stages k >= 1; FTMr / FTI / FSK for the final stage
template <int FTMr_, int FTI_, int FSK_, int TM1 = 1, int TN1 = 8, int TM2 = 1, int TN2 = 8, int TM3 = 1, int TN3 = 8>
struct Tune {
    static constexpr int FTMr = FTMr_, FTI = FTI_, FSK = FSK_;
    static constexpr int TMr[4] = {0, TM1, TM2, TM3};
    static constexpr int TN[4] = {0, TN1, TN2, TN3};
};

// shared-memory floats of the forward recurrent kernel
template <class S, int R, class TU>
struct FwdSmem {
    static constexpr int D = S::D;
    using FM = FinMap<S, R, TU::FTMr, TU::FTI, TU::FSK>;
    static constexpr int W = w_floats<S>();
    static constexpr int HS = cr4(R * St<S, D - 1>::BS);
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
    static constexpr int XCH = FM::XCH_FLOATS;
    static constexpr int TOTAL = W + HS + P + Q + XCH;
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
template <class S, int R, class TU, int k>
TTS_DEV const float *fwd_chain_pp(float *hs, float *P, float *Q, const float *wsm, int tid) {
    if constexpr (k == 0) {
        return (S::D == 1) ? hs : (((S::D - 2) % 2 == 0) ? P : Q);
    } else {
        const float *X = (k == S::D - 1) ? hs : (((S::D - 2 - k) % 2 == 0) ? P : Q);
        float *Y = ((S::D - 2 - (k - 1)) % 2 == 0) ? P : Q;
        fwd_stage<S, k, R, TU::TMr[k], TU::TN[k]>(X, wsm + WOff<S, k>::v, Y, tid);
        __syncthreads();
        return fwd_chain_pp<S, R, TU, k - 1>(hs, P, Q, wsm, tid);
    }
}

template <class S, int CELL, int R, int MODE, class TU>
__global__ void __launch_bounds__(NTHR, 1) k_rnn_fwd_s(const __grid_constant__ RnnFwdSArgs a) {
    extern __shared__ __align__(16) float smem[];
    using SM = FwdSmem<S, R, TU>;
    using FM = typename SM::FM;
    using TL = St<S, S::D - 1>;
    constexpr int G = S::G, H = n_in<S>(), GH = G * H;
    constexpr int NE = FM::NE;
    constexpr bool LSTM = (CELL == TTRNN_CELL_LSTM);
    static_assert(G == (LSTM ? 4 : 3), "gate count");
    const int tid = threadIdx.x;
    float *wsm = smem;
    float *hs = wsm + SM::W;
    float *P = hs + SM::HS;
    float *Q = P + SM::P;
    float *xch = Q + SM::Q;

    stage_weights_k<S, 0>(a.cores, wsm, tid);

    int mt, itg, kh;
    FM::coords(tid, mt, itg, kh);
    // hidden units owned by this thread: element e = n*SK + kh
    int hid[NE];
    float bhh[NE][4], weff[NE][4], bih[NE][4];
#pragma unroll
    for (int n = 0; n < NE; ++n) {
        hid[n] = FM::hidden(mt, itg, n * FM::SK + kh);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int col = g * H + hid[n];
            bhh[n][g] = a.bias_hh ? __ldg(a.bias_hh + col) : 0.f;
            weff[n][g] = (MODE == MODE_RANK1) ? __ldg(a.w_eff + col) : 0.f;
            bih[n][g] = (MODE == MODE_RANK1 && a.bias_ih) ? __ldg(a.bias_ih + col) : 0.f;
        }
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
        float cst[R][NE], hpr[R][NE];
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int n = 0; n < NE; ++n) {
                const bool ok = (row0 + b < a.B);
                cst[b][n] = (LSTM && ok && a.c_in) ? __ldg(a.c_in + (row0 + b) * H + hid[n]) : 0.f;
                hpr[b][n] = (ok && a.h_in) ? __ldg(a.h_in + (row0 + b) * H + hid[n]) : 0.f;
            }
        __syncthreads();

        for (int t = 0; t < a.steps; ++t) {
            // ---- operands of the gate phase, requested before the chain so their latency hides
            float xin[R][NE][4];
            float x1[R];
            if (MODE == MODE_XG) {
#pragma unroll
                for (int b = 0; b < R; ++b)
#pragma unroll
                    for (int n = 0; n < NE; ++n)
#pragma unroll
                        for (int g = 0; g < G; ++g)
                            xin[b][n][g] = (row0 + b < a.B)
                                ? __ldg(a.xg + (row0 + b) * a.xg_bstride + (long long)t * GH + g * H + hid[n]) : 0.f;
            } else {
#pragma unroll
                for (int b = 0; b < R; ++b)
                    x1[b] = (row0 + b < a.B) ? __ldg(a.x1 + (row0 + b) * a.x1_bstride + t) : 0.f;
            }
            // ---- stages d-1 .. 1
            const float *X0 = fwd_chain_pp<S, R, TU, S::D - 1>(hs, P, Q, wsm, tid);
            // ---- stage 0 (split over K) + reduce-scatter + gate math + state update
            float pre[R][NE][4];
            {
                float acc[R][FM::TMr][FM::TI][4];
                final_partial<S, R, FM>(X0, wsm + WOff<S, 0>::v, mt, itg, kh, acc);
                final_reduce<S, R, FM>(acc, pre, xch, tid, kh);
            }
#pragma unroll
            for (int b = 0; b < R; ++b)
#pragma unroll
                for (int n = 0; n < NE; ++n) {
                    const int h = hid[n];
                    float ain[4];
#pragma unroll
                    for (int g = 0; g < G; ++g)
                        ain[g] = (MODE == MODE_XG) ? xin[b][n][g] : fmaf(x1[b], weff[n][g], bih[n][g]);
                    float hnew;
                    if (LSTM) {
                        const float ig = sigmoidf_acc(pre[b][n][0] + bhh[n][0] + ain[0]);
                        const float fg = sigmoidf_acc(pre[b][n][1] + bhh[n][1] + ain[1]);
                        const float gg = tanhf(pre[b][n][2] + bhh[n][2] + ain[2]);
                        const float og = sigmoidf_acc(pre[b][n][3] + bhh[n][3] + ain[3]);
                        const float cn = fg * cst[b][n] + ig * gg;
                        cst[b][n] = cn;
                        hnew = og * tanhf(cn);
                    } else {
                        const float rg = sigmoidf_acc(ain[0] + (pre[b][n][0] + bhh[n][0]));
                        const float zg = sigmoidf_acc(ain[1] + (pre[b][n][1] + bhh[n][1]));
                        const float ng = tanhf(ain[2] + rg * (pre[b][n][2] + bhh[n][2]));
                        hnew = (1.0f - zg) * ng + zg * hpr[b][n];
                    }
                    hpr[b][n] = hnew;
                    hs[b * TL::BS + (h / TL::K) * TL::KS + (h % TL::K)] = hnew;
                    if (row0 + b < a.B) {
                        a.out[(row0 + b) * a.out_bstride + (long long)t * H + h] = hnew;
                        if (LSTM && a.c_save) a.c_save[(row0 + b) * a.out_bstride + (long long)t * H + h] = cst[b][n];
                    }
                }
            __syncthreads();
        }
#pragma unroll
        for (int b = 0; b < R; ++b)
#pragma unroll
            for (int n = 0; n < NE; ++n)
                if (row0 + b < a.B) {
                    if (a.h_out) a.h_out[(row0 + b) * H + hid[n]] = hpr[b][n];
                    if (LSTM && a.c_out) a.c_out[(row0 + b) * H + hid[n]] = cst[b][n];
                }
    }
}

}  // namespace tts
