// Stand-alone cell step (gate math + state update) for the "cell-step" mode of the module mirror:
// the reference's per-timestep module call (LSTMCell.forward lstm.py:23-41, GRUCell.forward
// gru.py:25-50) with the two pre-activation blocks already computed by TTLinear / TTLinearSet.
// Used when the fused sequence kernels cannot apply: is_naive=True (one TT matrix per gate,
// tt_linearset.py:5-38) and log_grads=True (per-step hooks, rnn_utils.py:42-215).
// One thread per (batch row, hidden unit); every access is coalesced along the hidden index.
#pragma once
#include <cuda_runtime.h>
#include "tt_stage.cuh"

struct CellStepArgs {
    long long B;
    int H;
    const float *a;        // (B, G*H)  W_ih x + b_ih
    const float *u;        // (B, G*H)  W_hh h + b_hh
    const float *h_prev;   // (B, H)
    const float *c_prev;   // (B, H)   LSTM only
    float *h, *c;          // forward outputs
    const float *dh, *dc;  // backward inputs (null = zero)
    float *da, *du;        // (B, G*H)
    float *dh_prev;        // (B, H)   direct term only (GRU: dh * z; LSTM: zero)
    float *dc_prev;        // (B, H)   LSTM only
    float *dc_total;       // (B, H)   LSTM only, optional: dL/dc_t including the path through h_t (what a tensor
                           //          hook on the reference's `cy` observes, lstm.py:38-39)
};

template <bool LSTM>
__global__ void __launch_bounds__(256) k_cell_fwd(const CellStepArgs p) {
    const long long n = p.B * p.H;
    const int H = p.H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int h = (int)(e - b * H);
        if (LSTM) {
            const float *ar = p.a + b * 4 * H, *ur = p.u + b * 4 * H;
            const float ig = tt_sigmoid(ar[h] + ur[h]);
            const float fg = tt_sigmoid(ar[H + h] + ur[H + h]);
            const float gg = tanhf(ar[2 * H + h] + ur[2 * H + h]);
            const float og = tt_sigmoid(ar[3 * H + h] + ur[3 * H + h]);
            const float cn = fg * p.c_prev[e] + ig * gg;
            p.c[e] = cn;
            p.h[e] = og * tanhf(cn);
        } else {
            const float *ar = p.a + b * 3 * H, *ur = p.u + b * 3 * H;
            const float rg = tt_sigmoid(ar[h] + ur[h]);
            const float zg = tt_sigmoid(ar[H + h] + ur[H + h]);
            const float ng = tanhf(ar[2 * H + h] + rg * ur[2 * H + h]);
            p.h[e] = (1.0f - zg) * ng + zg * p.h_prev[e];
        }
    }
}

// analytic backward of one step; gates are recomputed from a, u
template <bool LSTM>
__global__ void __launch_bounds__(256) k_cell_bwd(const CellStepArgs p) {
    const long long n = p.B * p.H;
    const int H = p.H;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int h = (int)(e - b * H);
        const float dh = p.dh ? p.dh[e] : 0.f;
        if (LSTM) {
            const float *ar = p.a + b * 4 * H, *ur = p.u + b * 4 * H;
            const float ig = tt_sigmoid(ar[h] + ur[h]);
            const float fg = tt_sigmoid(ar[H + h] + ur[H + h]);
            const float gg = tanhf(ar[2 * H + h] + ur[2 * H + h]);
            const float og = tt_sigmoid(ar[3 * H + h] + ur[3 * H + h]);
            const float cp = p.c_prev[e];
            const float cn = fg * cp + ig * gg;
            const float tc = tanhf(cn);
            const float dcn = (p.dc ? p.dc[e] : 0.f) + dh * og * (1.0f - tc * tc);
            const float d0 = dcn * gg * ig * (1.0f - ig);
            const float d1 = dcn * cp * fg * (1.0f - fg);
            const float d2 = dcn * ig * (1.0f - gg * gg);
            const float d3 = dh * tc * og * (1.0f - og);
            float *dar = p.da + b * 4 * H, *dur = p.du + b * 4 * H;
            dar[h] = d0; dar[H + h] = d1; dar[2 * H + h] = d2; dar[3 * H + h] = d3;
            dur[h] = d0; dur[H + h] = d1; dur[2 * H + h] = d2; dur[3 * H + h] = d3;
            p.dc_prev[e] = dcn * fg;
            if (p.dc_total) p.dc_total[e] = dcn;
            p.dh_prev[e] = 0.f;
        } else {
            const float *ar = p.a + b * 3 * H, *ur = p.u + b * 3 * H;
            const float un = ur[2 * H + h];
            const float rg = tt_sigmoid(ar[h] + ur[h]);
            const float zg = tt_sigmoid(ar[H + h] + ur[H + h]);
            const float ng = tanhf(ar[2 * H + h] + rg * un);
            const float d_n = dh * (1.0f - zg) * (1.0f - ng * ng);
            const float d_z = dh * (p.h_prev[e] - ng) * zg * (1.0f - zg);
            const float d_r = d_n * un * rg * (1.0f - rg);
            float *dar = p.da + b * 3 * H, *dur = p.du + b * 3 * H;
            dar[h] = d_r; dar[H + h] = d_z; dar[2 * H + h] = d_n;
            dur[h] = d_r; dur[H + h] = d_z; dur[2 * H + h] = d_n * rg;
            p.dh_prev[e] = dh * zg;
        }
    }
}
