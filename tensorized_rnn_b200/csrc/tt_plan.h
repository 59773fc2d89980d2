// Contraction plan of one TT matrix, shared by host launch code and kernels.
//
// A TT matrix W (M x N) with cores G_k (r_k, i_k, j_k, r_{k+1}) is applied to a
// tile of R batch rows as a chain of d small GEMMs, k = d-1 .. 0, the sweep
// order of the reference's tt_dense_matmul (t3nsor/ops.py:81-90):
//
//     X_{k-1}[b] (as [i_k][m][a_k])  =  X_k[b] ([m][kappa=(j_k,a_{k+1})])  *  W_k ([kappa][n=(i_k,a_k)])
//
// where m runs over (i_{k+1..d-1}, j_{0..k-1}).  X_k is kept in shared memory as
// a row-major matrix of Mrow_k rows x K_k columns with a padded row stride KS_k;
// the [i_k][m][a_k] order of a stage's output IS the row-major order of the next
// stage's input, so the reference's .contiguous() reshuffle (ops.py:89-90)
// becomes an addressing rule of the epilogue and costs nothing.
#pragma once
#include <stdint.h>
#include "../../include/ttrnn_b200.h"

#define TT_MAX_D TTRNN_MAX_CORES

#ifdef __CUDACC__
#define TT_HD __host__ __device__
#else
#define TT_HD
#endif

struct StagePlan {
    int K;      // j_k * r_{k+1}: contraction length
    int N;      // i_k * r_k:     output columns
    int Mrow;   // rows of X_k per batch row
    int KS;     // physical row stride of X_k (floats), KS >= K
    int NS;     // physical row stride of W_k in shared memory, NS >= N
    int r;      // r_k: length of a contiguous a-run in the output
    int Jp;     // j_{k-1} (1 for k == 0)
    int KSo;    // row stride of the output buffer X_{k-1} (1 for k == 0)
    int ISo;    // distance between consecutive i_k in the output buffer
    int BSo;    // batch-row stride of the output buffer
    int BS;     // batch-row stride of X_k  (= Mrow * KS)
    int I, J;   // i_k, j_k
    int rn;     // r_{k+1}
    int w_off;  // offset of W_k inside the shared-memory weight area (floats)
    int c_off;  // offset of core k inside the core blob (floats)
    int xpp;    // ping-pong slot of X_k for forward-only chains: 0 = input slot, 1 = P, 2 = Q
    int xall;   // per-batch-row offset of X_k when every X_k is kept (backward), inside its space
    int xsp;    // backward only: 0 = X_k slot in shared memory, 1 = in the global (L2-resident) spill area
};

struct ChainPlan {
    int d;
    int n_in;        // N = prod j_k
    int n_out;       // M = prod i_k
    int w_floats;    // shared-memory floats of all W_k (padded strides)
    int core_floats; // floats of all cores in the blob
    int g_IS;        // stage-0 output buffer G: stride between consecutive i_0
    int g_BS;        // stage-0 output buffer G: batch-row stride
    int in_BS;       // per-batch-row floats of X_{d-1} (the chain input)
    int pp_floats[2];// per-batch-row floats of ping-pong slots P, Q
    int all_floats;  // per-batch-row shared-memory floats of the kept X_k slots (backward)
    int spill_floats;// per-batch-row floats of the X_k slots placed in the global spill area
    StagePlan st[TT_MAX_D];
};

TT_HD static inline int tt_round4(int v) { return (v + 3) & ~3; }
// row stride >= k, multiple of 4 with an odd number of float4s so that rows that are read
// by different lanes at the same column fall into different bank groups
static inline int tt_pad_stride(int k) {
    if (k % 4 != 0) return k;
    int q = k / 4;
    return (q % 2 == 1) ? k : k + 4;
}

// returns 0 on success
static inline int tt_build_plan(const ttrnn_tt_shape *s, ChainPlan *p) {
    const int d = s->d;
    if (d < 1 || d > TT_MAX_D) return 1;
    if (s->ranks[0] != 1 || s->ranks[d] != 1) return 2;
    int64_t n_in = 1, n_out = 1;
    for (int k = 0; k < d; ++k) {
        if (s->in_modes[k] < 1 || s->out_modes[k] < 1 || s->ranks[k] < 1) return 3;
        n_in *= s->in_modes[k];
        n_out *= s->out_modes[k];
    }
    if (n_in > (1 << 20) || n_out > (1 << 22)) return 4;
    p->d = d;
    p->n_in = (int)n_in;
    p->n_out = (int)n_out;
    int w_off = 0, c_off = 0;
    for (int k = 0; k < d; ++k) {
        StagePlan &t = p->st[k];
        t.I = s->out_modes[k];
        t.J = s->in_modes[k];
        t.r = s->ranks[k];
        t.rn = s->ranks[k + 1];
        t.K = t.J * t.rn;
        t.N = t.I * t.r;
        int64_t mrow = 1;
        for (int m = k + 1; m < d; ++m) mrow *= s->out_modes[m];
        for (int m = 0; m < k; ++m) mrow *= s->in_modes[m];
        t.Mrow = (int)mrow;
        t.KS = tt_pad_stride(t.K);
        t.NS = tt_pad_stride(t.N);
        t.BS = t.Mrow * t.KS;
        t.w_off = w_off;
        w_off += tt_round4(t.K * t.NS);
        t.c_off = c_off;
        c_off += t.r * t.I * t.J * t.rn;
    }
    p->w_floats = w_off;
    p->core_floats = c_off;
    p->g_IS = p->st[0].Mrow | 1;
    p->g_BS = tt_round4(p->st[0].I * p->g_IS);
    for (int k = 0; k < d; ++k) {
        StagePlan &t = p->st[k];
        if (k > 0) {
            t.Jp = p->st[k - 1].J;
            t.KSo = p->st[k - 1].KS;
            t.ISo = (t.Mrow / t.Jp) * t.KSo;
            t.BSo = p->st[k - 1].BS;
        } else {
            t.Jp = 1;
            t.KSo = 1;
            t.ISo = p->g_IS;
            t.BSo = p->g_BS;
        }
    }
    // forward-only chains: X_{d-1} lives in the input slot, then P, Q, P, ...
    p->in_BS = tt_round4(p->st[d - 1].BS);
    p->pp_floats[0] = p->pp_floats[1] = 0;
    int all = 0;
    for (int k = d - 1; k >= 0; --k) {
        StagePlan &t = p->st[k];
        t.xall = all;
        all += tt_round4(t.BS);
        if (k == d - 1) {
            t.xpp = 0;
        } else {
            int slot = ((d - 2 - k) % 2);          // X_{d-2} -> P, X_{d-3} -> Q, ...
            t.xpp = 1 + slot;
            if (tt_round4(t.BS) > p->pp_floats[slot]) p->pp_floats[slot] = tt_round4(t.BS);
        }
    }
    p->all_floats = all;
    p->spill_floats = 0;
    for (int k = 0; k < d; ++k) p->st[k].xsp = 0;
    return 0;
}

// Backward keeps every X_k alive at once.  When they do not fit the shared-memory budget
// (floats per batch row), the largest slots are moved to a per-CTA global scratch area that
// stays L1/L2 resident.  Recomputes xall offsets inside each space.
static inline void tt_place_slots(ChainPlan *p, long long budget_floats_per_row) {
    const int d = p->d;
    for (int k = 0; k < d; ++k) p->st[k].xsp = 0;
    long long in_smem = 0;
    for (int k = 0; k < d; ++k) in_smem += tt_round4(p->st[k].BS);
    while (in_smem > budget_floats_per_row) {
        int big = -1;
        for (int k = 0; k < d; ++k)
            if (!p->st[k].xsp && (big < 0 || p->st[k].BS > p->st[big].BS)) big = k;
        if (big < 0) break;
        p->st[big].xsp = 1;
        in_smem -= tt_round4(p->st[big].BS);
    }
    int a = 0, sp = 0;
    for (int k = d - 1; k >= 0; --k) {
        StagePlan &t = p->st[k];
        if (t.xsp) { t.xall = sp; sp += tt_round4(t.BS); }
        else       { t.xall = a;  a += tt_round4(t.BS); }
    }
    p->all_floats = a;
    p->spill_floats = sp;
}
