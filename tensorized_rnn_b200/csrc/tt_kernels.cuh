// Kernels of the TT recurrent engine (sm_100a).  Launch code is in ttrnn_capi.cu.
//
//   k_ttlinear_fwd   batched TT matvec over many rows (the ih projection of a whole
//                    time chunk, and stand-alone TTLinear): persistent CTAs, cores
//                    staged once in shared memory, coalesced float4 row loads/stores
//   k_rnn_fwd        persistent recurrent kernel: one CTA owns R batch rows for all
//                    timesteps of the launch; hh chain core by core in shared memory,
//                    gate math + state update fused, h_t written once per step
//   k_ttlinear_bwd / k_rnn_bwd   the mirrors (reverse time for the recurrence)
//   k_reduce_partials            deterministic sum of per-CTA gradient partials
#pragma once
#include <cuda_runtime.h>
#include "tt_stage.cuh"

#define TT_NTHREADS 256

struct TTLinFwdArgs {
    ChainPlan p;
    int tile[TT_MAX_D];
    int R;                 // rows per tile
    long long rows;
    int rows_per_b;        // row -> (b = row / rows_per_b, t = row % rows_per_b)
    long long x_bstride;   // floats between consecutive b in x (row stride inside b is n_in)
    long long y_bstride;   // same for y (row stride n_out)
    const float *x;
    const float *cores;    // core blob
    const float *bias;     // (n_out) or null
    const float *bias2;    // (n_out) or null  (LSTM: hh bias folded into the projection)
    float *y;
};

struct RnnFwdArgs {
    ChainPlan p;           // hh chain
    int tile[TT_MAX_D];
    int cell;              // TTRNN_CELL_*
    int R;                 // batch rows per CTA
    int H, G;
    int steps;             // timesteps in this launch
    long long B;
    const float *xg;       // ih projection of this chunk: row b, step t at xg + (b*xg_bstride + t*G*H)
    long long xg_bstride;
    const float *cores;    // hh core blob
    const float *bias_hh;  // (G*H) or null; GRU only (LSTM folds it into xg)
    const float *h_in;     // (B,H) or null = zeros
    const float *c_in;     // (B,H) or null = zeros (LSTM)
    float *out;            // h_t of this chunk: out + b*out_bstride + t*H
    long long out_bstride;
    float *c_save;         // c_t, same addressing as out, or null (inference)
    float *h_out;          // (B,H) state after the last step of this launch
    float *c_out;          // (B,H) (LSTM)
};

TT_DEV void tt_cp_async16(float *smem_dst, const float *gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
TT_DEV void tt_cp_async4(float *smem_dst, const float *gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc));
}
TT_DEV void tt_cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Run the forward chain on R rows.  X_{d-1} is in `xin`; ping-pong slots P, Q; the last
// stage writes into `g` (layout [b][i_0*g_IS + m]).  One __syncthreads after every stage.
TT_DEV void tt_chain_fwd_pp(const ChainPlan &p, const int *tile, int R, const float *xin, float *P, float *Q,
                            float *g, const float *wsm, int tid) {
    for (int k = p.d - 1; k >= 0; --k) {
        const StagePlan &s = p.st[k];
        const float *X = (s.xpp == 0) ? xin : (s.xpp == 1 ? P : Q);
        float *Y = (k == 0) ? g : (p.st[k - 1].xpp == 1 ? P : Q);
        tt_stage_fwd(s, tile[k], R, X, wsm + s.w_off, Y, tid, TT_NTHREADS);
        __syncthreads();
    }
}

// position of pre-activation column pcol (0 .. n_out) inside the G buffer of one batch row
TT_DEV int tt_g_index(const ChainPlan &p, int pcol) {
    const int mrow0 = p.st[0].Mrow;
    const int i0 = pcol / mrow0;
    return i0 * p.g_IS + (pcol - i0 * mrow0);
}
// position of input column c inside the X_{d-1} slot of one batch row
TT_DEV int tt_in_index(const ChainPlan &p, int c) {
    const StagePlan &s = p.st[p.d - 1];
    const int row = c / s.K;
    return row * s.KS + (c - row * s.K);
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TT_NTHREADS)
k_ttlinear_fwd(const __grid_constant__ TTLinFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const ChainPlan &p = a.p;
    const int tid = threadIdx.x;
    float *wsm = smem;
    float *xin = wsm + p.w_floats;
    float *P = xin + a.R * p.in_BS;
    float *Q = P + a.R * p.pp_floats[0];
    float *g = Q + a.R * p.pp_floats[1];

    tt_stage_weights(p, a.cores, wsm, tid, TT_NTHREADS);
    const long long ntiles = (a.rows + a.R - 1) / a.R;
    for (long long tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const long long row0 = tile_i * a.R;
        // rows -> X_{d-1} slot (zeros for rows past the end)
        const int per_row = p.n_in;
        if ((per_row & 3) == 0 && (p.st[p.d - 1].K & 3) == 0) {
            for (int e = tid * 4; e < a.R * per_row; e += TT_NTHREADS * 4) {
                const int b = e / per_row, c = e - b * per_row;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const long long row = row0 + b;
                if (row < a.rows) {
                    const long long bb = row / a.rows_per_b, tt = row - bb * a.rows_per_b;
                    v = __ldg(reinterpret_cast<const float4 *>(a.x + bb * a.x_bstride + tt * per_row + c));
                }
                tt_st4(xin + b * p.st[p.d - 1].BS + tt_in_index(p, c), v);
            }
        } else {
            for (int e = tid; e < a.R * per_row; e += TT_NTHREADS) {
                const int b = e / per_row, c = e - b * per_row;
                float v = 0.f;
                const long long row = row0 + b;
                if (row < a.rows) {
                    const long long bb = row / a.rows_per_b, tt = row - bb * a.rows_per_b;
                    v = __ldg(a.x + bb * a.x_bstride + tt * per_row + c);
                }
                xin[b * p.st[p.d - 1].BS + tt_in_index(p, c)] = v;
            }
        }
        __syncthreads();
        tt_chain_fwd_pp(p, a.tile, a.R, xin, P, Q, g, wsm, tid);
        // G -> y (+ biases), coalesced over the output column
        const int nout = p.n_out;
        for (int e = tid; e < a.R * nout; e += TT_NTHREADS) {
            const int b = e / nout, c = e - b * nout;
            const long long row = row0 + b;
            if (row < a.rows) {
                const long long bb = row / a.rows_per_b, tt = row - bb * a.rows_per_b;
                float v = g[b * p.g_BS + tt_g_index(p, c)];
                if (a.bias) v += __ldg(a.bias + c);
                if (a.bias2) v += __ldg(a.bias2 + c);
                a.y[bb * a.y_bstride + tt * nout + c] = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Persistent recurrent forward.  Shared memory:
//   [W_hh][bias_hh G*H][c state R*H][xg tile R*G*H][G buf R*g_BS][h slot R*in_BS][P][Q]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TT_NTHREADS)
k_rnn_fwd(const __grid_constant__ RnnFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const ChainPlan &p = a.p;
    const int tid = threadIdx.x;
    const int H = a.H, GH = a.G * a.H, R = a.R;
    float *wsm = smem;
    float *bsm = wsm + p.w_floats;                                     // hh bias: GRU only
    float *csm = bsm + (a.cell == TTRNN_CELL_LSTM ? 0 : tt_round4(GH));
    float *xgs = csm + tt_round4(R * H);
    float *g = xgs + tt_round4(R * GH);
    float *hin = g + R * p.g_BS;
    float *P = hin + R * p.in_BS;
    float *Q = P + R * p.pp_floats[0];

    tt_stage_weights(p, a.cores, wsm, tid, TT_NTHREADS);
    if (a.cell != TTRNN_CELL_LSTM)
        for (int e = tid; e < GH; e += TT_NTHREADS) bsm[e] = a.bias_hh ? __ldg(a.bias_hh + e) : 0.f;
    const long long ntiles = (a.B + R - 1) / R;
    for (long long tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const long long row0 = tile_i * R;
    __syncthreads();
    for (int e = tid; e < R * H; e += TT_NTHREADS) {
        const int b = e / H, h = e - b * H;
        const long long row = row0 + b;
        float hv = 0.f, cv = 0.f;
        if (row < a.B) {
            if (a.h_in) hv = __ldg(a.h_in + row * H + h);
            if (a.c_in) cv = __ldg(a.c_in + row * H + h);
        }
        hin[b * p.st[p.d - 1].BS + tt_in_index(p, h)] = hv;
        csm[e] = cv;
    }
    __syncthreads();

    const bool vec16 = (GH % 4 == 0);
    for (int t = 0; t < a.steps; ++t) {
        // prefetch this step's ih projection tile while the chain runs
        if (vec16) {
            for (int e = tid * 4; e < R * GH; e += TT_NTHREADS * 4) {
                const int b = e / GH, c = e - b * GH;
                if (row0 + b < a.B) tt_cp_async16(xgs + e, a.xg + (row0 + b) * a.xg_bstride + (long long)t * GH + c);
            }
        } else {
            for (int e = tid; e < R * GH; e += TT_NTHREADS) {
                const int b = e / GH, c = e - b * GH;
                if (row0 + b < a.B) tt_cp_async4(xgs + e, a.xg + (row0 + b) * a.xg_bstride + (long long)t * GH + c);
            }
        }
        for (int k = p.d - 1; k >= 0; --k) {
            const StagePlan &s = p.st[k];
            const float *X = (s.xpp == 0) ? hin : (s.xpp == 1 ? P : Q);
            float *Y = (k == 0) ? g : (p.st[k - 1].xpp == 1 ? P : Q);
            tt_stage_fwd(s, a.tile[k], R, X, wsm + s.w_off, Y, tid, TT_NTHREADS);
            if (k == 0) tt_cp_async_wait_all();
            __syncthreads();
        }
        // gate math + state update; one thread per (row, hidden unit)
        for (int e = tid; e < R * H; e += TT_NTHREADS) {
            const int b = e / H, h = e - b * H;
            const long long row = row0 + b;
            if (row >= a.B) continue;
            const float *gb = g + b * p.g_BS;
            const float *xb = xgs + b * GH;
            float hnew;
            const int hidx = b * p.st[p.d - 1].BS + tt_in_index(p, h);
            if (a.cell == TTRNN_CELL_LSTM) {
                const float pi = gb[tt_g_index(p, h)] + xb[h];
                const float pf = gb[tt_g_index(p, H + h)] + xb[H + h];
                const float pg = gb[tt_g_index(p, 2 * H + h)] + xb[2 * H + h];
                const float po = gb[tt_g_index(p, 3 * H + h)] + xb[3 * H + h];
                const float ig = tt_sigmoid(pi), fg = tt_sigmoid(pf), gg = tanhf(pg), og = tt_sigmoid(po);
                const float cnew = fg * csm[e] + ig * gg;
                hnew = og * tanhf(cnew);
                csm[e] = cnew;
                if (a.c_save) a.c_save[row * a.out_bstride + (long long)t * H + h] = cnew;
            } else {
                const float ur = gb[tt_g_index(p, h)] + bsm[h];
                const float uz = gb[tt_g_index(p, H + h)] + bsm[H + h];
                const float un = gb[tt_g_index(p, 2 * H + h)] + bsm[2 * H + h];
                const float rg = tt_sigmoid(xb[h] + ur);
                const float zg = tt_sigmoid(xb[H + h] + uz);
                const float ng = tanhf(xb[2 * H + h] + rg * un);
                const float hprev = hin[hidx];
                hnew = (1.0f - zg) * ng + zg * hprev;
            }
            hin[hidx] = hnew;
            a.out[row * a.out_bstride + (long long)t * H + h] = hnew;
        }
        __syncthreads();
    }
    for (int e = tid; e < R * H; e += TT_NTHREADS) {
        const int b = e / H, h = e - b * H;
        const long long row = row0 + b;
        if (row < a.B) {
            if (a.h_out) a.h_out[row * H + h] = hin[b * p.st[p.d - 1].BS + tt_in_index(p, h)];
            if (a.c_out && a.cell == TTRNN_CELL_LSTM) a.c_out[row * H + h] = csm[e];
        }
    }
    }
}

// ---------------------------------------------------------------------------
// FP32 FFMA throughput probe (roofline denominator; see include/ttrnn_b200.h)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TT_NTHREADS)
k_ffma_probe(int iters, float *sink) {
    float acc[16];
    const float a = 1.0f + 1e-7f * threadIdx.x, b = 1e-9f * blockIdx.x;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = (float)j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc[j], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 123.456f) sink[0] = s;   // keeps the loop alive; practically never true
}

// ===========================================================================
// Backward kernels
// ===========================================================================
struct TTLinBwdArgs {
    ChainPlan p;
    int tile[TT_MAX_D];    // forward tiles (recompute of the intermediates)
    int tile_bd[TT_MAX_D]; // backward-data tiles
    int mg[TT_MAX_D];      // m-groups of the backward-weight stage
    int R;
    long long rows;
    int rows_per_b;        // row -> (b = row / rows_per_b, t = row % rows_per_b)
    long long x_bstride;   // floats between consecutive b in x   (row stride inside b is n_in)
    long long dy_bstride;  // same for dy                        (row stride n_out)
    long long dx_bstride;  // same for dx                        (row stride n_in)
    const float *x;
    const float *cores;
    const float *dy;
    float *dx;             // or null
    float *partial;        // [gridDim.x][core_floats + n_out]: += at the end of the launch
    float *spill;          // [gridDim.x][R * spill_floats] or null
    int want_dbias;
};

struct RnnBwdArgs {
    ChainPlan p;
    int tile[TT_MAX_D];
    int tile_bd[TT_MAX_D];
    int mg[TT_MAX_D];
    int cell, R, H, G;
    int t0, steps, T;      // this launch covers timesteps [t0, t0+steps) of a length-T sequence
    long long B;
    float *xg;             // chunk buffer (b, steps, G*H): ih projection in, delta_ih out
    long long xg_bstride;
    const float *cores;
    const float *bias_hh;  // GRU
    const float *hs;       // (B,T,H) outputs of this layer
    const float *cs;       // (B,T,H) cell states of this layer (LSTM)
    const float *h0;       // (B,H) or null
    const float *c0;       // (B,H) or null
    const float *dhs;      // (B,T,H) gradient wrt this layer's outputs, or null
    const float *dh_in;    // (B,H) gradient flowing into h_{t0+steps-1} from later steps, or null
    const float *dc_in;    // (B,H) same for c, or null
    float *dh_out;         // (B,H) gradient wrt h_{t0-1}
    float *dc_out;         // (B,H)
    float *partial;        // [gridDim.x][core_floats + G*H]: hh core grads + hh bias grad (GRU)
    float *spill;          // [gridDim.x][R * spill_floats] or null
    // log_grads (reference lstm.py:35-39,66-80): total gradient of h_t / c_t of every step, what the reference's tensor
    // hooks on hy / cy see; (B,T,H) addressed like hs, or null
    float *dh_log;
    float *dc_log;
};

// base of the X_k slot: shared memory or the per-CTA global spill area
TT_DEV float *tt_slot(const ChainPlan &p, int k, int R, float *xa, float *spill) {
    const StagePlan &s = p.st[k];
    return (s.xsp ? spill : xa) + (long long)R * s.xall;
}

TT_DEV int tt_bd_tile_m(int code) { return code == TT_TILE_8x8 || code == TT_TILE_8x4 ? 8 : (code == TT_TILE_4x8 || code == TT_TILE_4x4 ? 4 : (code == TT_TILE_2x4 ? 2 : 1)); }

TT_DEV void tt_stage_bwd_data_dispatch(const StagePlan &s, int code, int R, const float *dY, const float *W,
                                       float *dX, int tid, int nthr) {
    switch (code) {
    case TT_TILE_8x8: tt_stage_bwd_data<8, 8, true>(s, R, dY, W, dX, tid, nthr); break;
    case TT_TILE_4x8: tt_stage_bwd_data<4, 8, true>(s, R, dY, W, dX, tid, nthr); break;
    case TT_TILE_8x4: tt_stage_bwd_data<8, 4, true>(s, R, dY, W, dX, tid, nthr); break;
    case TT_TILE_4x4: tt_stage_bwd_data<4, 4, true>(s, R, dY, W, dX, tid, nthr); break;
    case TT_TILE_2x4: tt_stage_bwd_data<2, 4, true>(s, R, dY, W, dX, tid, nthr); break;
    case TT_TILE_1x4: tt_stage_bwd_data<1, 4, true>(s, R, dY, W, dX, tid, nthr); break;
    default: tt_stage_bwd_data<2, 2, false>(s, R, dY, W, dX, tid, nthr); break;
    }
}

TT_DEV void tt_stage_bwd_weight_dispatch(const StagePlan &s, int R, const float *X, const float *dY, float *dWs,
                                         int mg, int tid, int nthr) {
    const bool vx = (s.K % 4 == 0), vy = (s.r % 4 == 0);
    if (vx && vy) tt_stage_bwd_weight<true, true>(s, R, X, dY, dWs, mg, tid, nthr);
    else if (vx) tt_stage_bwd_weight<true, false>(s, R, X, dY, dWs, mg, tid, nthr);
    else tt_stage_bwd_weight<false, false>(s, R, X, dY, dWs, mg, tid, nthr);
}

// Forward chain keeping every X_k in its own slot (xall offsets); stops before stage `kstop`
// (kstop = 0 runs all stages and writes G; kstop = 1 skips the last GEMM).
TT_DEV void tt_chain_fwd_keep(const ChainPlan &p, const int *tile, int R, float *xa, float *spill, float *g,
                              const float *wsm, int kstop, int tid) {
    for (int k = p.d - 1; k >= kstop; --k) {
        const StagePlan &s = p.st[k];
        const float *X = tt_slot(p, k, R, xa, spill);
        float *Y = (k == 0) ? g : tt_slot(p, k - 1, R, xa, spill);
        tt_stage_fwd(s, tile[k], R, X, wsm + s.w_off, Y, tid, TT_NTHREADS);
        __syncthreads();
    }
}

// Backward chain.  On entry G holds dY_0 and slot k holds X_k; on exit `dxin` (X_{d-1} layout)
// holds dX_{d-1} if want_dx.  Slot k (k < d-1) is overwritten by dX_k.
TT_DEV void tt_chain_bwd(const ChainPlan &p, const int *tile_bd, const int *mg, int R, float *xa, float *spill,
                         float *g, float *dxin, const float *wsm, float *dws, bool want_dx, int tid) {
    for (int k = 0; k < p.d; ++k) {
        const StagePlan &s = p.st[k];
        float *X = tt_slot(p, k, R, xa, spill);
        const float *dY = (k == 0) ? g : tt_slot(p, k - 1, R, xa, spill);
        tt_stage_bwd_weight_dispatch(s, R, X, dY, dws + s.w_off, mg[k], tid, TT_NTHREADS);
        if (k == p.d - 1 && !want_dx) break;
        float *dX = (k == p.d - 1) ? dxin : X;
        if (dX == X) __syncthreads();          // X_k is overwritten in place: all readers must be done
        tt_stage_bwd_data_dispatch(s, tile_bd[k], R, dY, wsm + s.w_off, dX, tid, TT_NTHREADS);
        __syncthreads();
    }
    __syncthreads();
}

// dW (shared, W layout) [+ bias grads] -> += into this CTA's partial slot (blob layout)
TT_DEV void tt_flush_partial(const ChainPlan &p, const float *dws, const float *dbs, int nb, float *slot, int tid) {
    for (int k = 0; k < p.d; ++k) {
        const StagePlan &s = p.st[k];
        const int total = s.r * s.I * s.J * s.rn;
        for (int e = tid; e < total; e += TT_NTHREADS) {
            int ap = e % s.rn;
            int t = e / s.rn;
            int j = t % s.J;
            t /= s.J;
            int i = t % s.I;
            int a = t / s.I;
            slot[s.c_off + e] += dws[s.w_off + (j * s.rn + ap) * s.NS + i * s.r + a];
        }
    }
    if (dbs)
        for (int e = tid; e < nb; e += TT_NTHREADS) slot[p.core_floats + e] += dbs[e];
}

// ---------------------------------------------------------------------------
// Shared memory: [W][dW][db n_out][x slots R*all_floats][G R*g_BS][dx slot R*in_BS]
__global__ void __launch_bounds__(TT_NTHREADS)
k_ttlinear_bwd(const __grid_constant__ TTLinBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const ChainPlan &p = a.p;
    const int tid = threadIdx.x, R = a.R;
    float *wsm = smem;
    float *dws = wsm + p.w_floats;
    float *dbs = dws + p.w_floats;
    float *xa = dbs + tt_round4(p.n_out);
    float *g = xa + R * p.all_floats;
    float *dxin = g + R * p.g_BS;
    float *spill = a.spill ? a.spill + (long long)blockIdx.x * R * p.spill_floats : nullptr;

    tt_stage_weights(p, a.cores, wsm, tid, TT_NTHREADS);
    for (int e = tid; e < p.w_floats; e += TT_NTHREADS) dws[e] = 0.f;
    for (int e = tid; e < p.n_out; e += TT_NTHREADS) dbs[e] = 0.f;
    const StagePlan &sl = p.st[p.d - 1];
    float *xslot = tt_slot(p, p.d - 1, R, xa, spill);
    const int nin = p.n_in, nout = p.n_out;
    const long long ntiles = (a.rows + R - 1) / R;
    __syncthreads();
    for (long long tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const long long row0 = tile_i * R;
        for (int e = tid; e < R * nin; e += TT_NTHREADS) {
            const int b = e / nin, c = e - b * nin;
            const long long row = row0 + b;
            float v = 0.f;
            if (row < a.rows) {
                const long long bb = row / a.rows_per_b, tt = row - bb * a.rows_per_b;
                v = __ldg(a.x + bb * a.x_bstride + tt * nin + c);
            }
            xslot[b * sl.BS + tt_in_index(p, c)] = v;
        }
        // dy rows -> G layout; bias gradient = column sums (one owner thread per column)
        for (int c = tid; c < nout; c += TT_NTHREADS) {
            const int gi = tt_g_index(p, c);
            float sum = 0.f;
            for (int b = 0; b < R; ++b) {
                const long long row = row0 + b;
                float v = 0.f;
                if (row < a.rows) {
                    const long long bb = row / a.rows_per_b, tt = row - bb * a.rows_per_b;
                    v = __ldg(a.dy + bb * a.dy_bstride + tt * nout + c);
                }
                g[b * p.g_BS + gi] = v;
                sum += v;
            }
            dbs[c] += sum;
        }
        __syncthreads();
        tt_chain_fwd_keep(p, a.tile, R, xa, spill, g, wsm, /*kstop=*/1, tid);
        tt_chain_bwd(p, a.tile_bd, a.mg, R, xa, spill, g, dxin, wsm, dws, a.dx != nullptr, tid);
        if (a.dx) {
            for (int e = tid; e < R * nin; e += TT_NTHREADS) {
                const int b = e / nin, c = e - b * nin;
                const long long row = row0 + b;
                if (row < a.rows) {
                    const long long bb = row / a.rows_per_b, tt = row - bb * a.rows_per_b;
                    a.dx[bb * a.dx_bstride + tt * nin + c] = dxin[b * sl.BS + tt_in_index(p, c)];
                }
            }
        }
        __syncthreads();
    }
    tt_flush_partial(p, dws, a.want_dbias ? dbs : nullptr, nout,
                     a.partial + (long long)blockIdx.x * (p.core_floats + nout), tid);
}

// ---------------------------------------------------------------------------
// Reverse-time persistent BPTT.  Shared memory:
//   [W][dW][bias_hh GH][db_hh GH][x slots R*all_floats][G R*g_BS][xg tile R*GH]
//   [dh chain R*in_BS][dh direct R*H][dc R*H][c_prev R*H]
__global__ void __launch_bounds__(TT_NTHREADS)
k_rnn_bwd(const __grid_constant__ RnnBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const ChainPlan &p = a.p;
    const int tid = threadIdx.x, R = a.R, H = a.H, GH = a.G * a.H;
    float *wsm = smem;
    float *dws = wsm + p.w_floats;
    const bool lstm = (a.cell == TTRNN_CELL_LSTM);
    const int nbias = lstm ? 0 : tt_round4(GH);      // hh bias and its gradient: GRU only
    float *bsm = dws + p.w_floats;
    float *dbs = bsm + nbias;
    float *xa = dbs + nbias;
    float *g = xa + R * p.all_floats;
    float *xgs = g + R * p.g_BS;
    float *spill = a.spill ? a.spill + (long long)blockIdx.x * R * p.spill_floats : nullptr;
    float *dhc = xgs + tt_round4(R * GH);
    float *dhd = dhc + R * p.in_BS;
    float *dcs = dhd + tt_round4(R * H);
    float *cps = dcs + tt_round4(R * H);

    const StagePlan &sl = p.st[p.d - 1];
    float *hslot = tt_slot(p, p.d - 1, R, xa, spill);
    const bool vec16 = (GH % 4 == 0);

    tt_stage_weights(p, a.cores, wsm, tid, TT_NTHREADS);
    for (int e = tid; e < p.w_floats; e += TT_NTHREADS) dws[e] = 0.f;
    if (!lstm)
        for (int e = tid; e < GH; e += TT_NTHREADS) {
            bsm[e] = a.bias_hh ? __ldg(a.bias_hh + e) : 0.f;
            dbs[e] = 0.f;
        }
    const long long ntiles = (a.B + R - 1) / R;
    for (long long tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const long long row0 = tile_i * R;
        __syncthreads();
        for (int e = tid; e < R * p.in_BS; e += TT_NTHREADS) dhc[e] = 0.f;
        for (int e = tid; e < R * H; e += TT_NTHREADS) {
            const int b = e / H, h = e - b * H;
            const long long row = row0 + b;
            float dh = 0.f, dc = 0.f;
            if (row < a.B) {
                if (a.dh_in) dh = __ldg(a.dh_in + row * H + h);
                if (a.dc_in) dc = __ldg(a.dc_in + row * H + h);
            }
            dhd[e] = dh;
            dcs[e] = dc;
        }
        __syncthreads();
        for (int t = a.steps - 1; t >= 0; --t) {
            const int tg = a.t0 + t;                    // global timestep
            // ---- load h_{t-1}, c_{t-1}; prefetch the ih projection of step t
            if (vec16) {
                for (int e = tid * 4; e < R * GH; e += TT_NTHREADS * 4) {
                    const int b = e / GH, c = e - b * GH;
                    if (row0 + b < a.B) tt_cp_async16(xgs + e, a.xg + (row0 + b) * a.xg_bstride + (long long)t * GH + c);
                }
            } else {
                for (int e = tid; e < R * GH; e += TT_NTHREADS) {
                    const int b = e / GH, c = e - b * GH;
                    if (row0 + b < a.B) tt_cp_async4(xgs + e, a.xg + (row0 + b) * a.xg_bstride + (long long)t * GH + c);
                }
            }
            for (int e = tid; e < R * H; e += TT_NTHREADS) {
                const int b = e / H, h = e - b * H;
                const long long row = row0 + b;
                float hv = 0.f, cv = 0.f;
                if (row < a.B) {
                    if (tg > 0) {
                        hv = __ldg(a.hs + (row * a.T + (tg - 1)) * H + h);
                        if (lstm) cv = __ldg(a.cs + (row * a.T + (tg - 1)) * H + h);
                    } else {
                        if (a.h0) hv = __ldg(a.h0 + row * H + h);
                        if (lstm && a.c0) cv = __ldg(a.c0 + row * H + h);
                    }
                }
                hslot[b * sl.BS + tt_in_index(p, h)] = hv;
                cps[e] = cv;
            }
            __syncthreads();
            // ---- recompute the hh chain, keeping every intermediate
            for (int k = p.d - 1; k >= 0; --k) {
                const StagePlan &s = p.st[k];
                const float *X = tt_slot(p, k, R, xa, spill);
                float *Y = (k == 0) ? g : tt_slot(p, k - 1, R, xa, spill);
                tt_stage_fwd(s, a.tile[k], R, X, wsm + s.w_off, Y, tid, TT_NTHREADS);
                if (k == 0) tt_cp_async_wait_all();
                __syncthreads();
            }
            // ---- gates, their gradients; G <- delta_hh, xg (global) <- delta_ih
            for (int h = tid; h < H; h += TT_NTHREADS) {
                const int gi0 = tt_g_index(p, h), gi1 = tt_g_index(p, H + h), gi2 = tt_g_index(p, 2 * H + h);
                const int gi3 = lstm ? tt_g_index(p, 3 * H + h) : 0;
                const int hix = tt_in_index(p, h);
                float db0 = 0.f, db1 = 0.f, db2 = 0.f;
                for (int b = 0; b < R; ++b) {
                    const long long row = row0 + b;
                    float *gb = g + b * p.g_BS;
                    if (row >= a.B) {
                        gb[gi0] = 0.f; gb[gi1] = 0.f; gb[gi2] = 0.f;
                        if (lstm) gb[gi3] = 0.f;
                        continue;
                    }
                    const float *xb = xgs + b * GH;
                    float dh = dhd[b * H + h] + dhc[b * sl.BS + hix];
                    if (a.dhs) dh += __ldg(a.dhs + (row * a.T + tg) * H + h);
                    float *dxg = a.xg + row * a.xg_bstride + (long long)t * GH;
                    if (a.dh_log) a.dh_log[(row * a.T + tg) * H + h] = dh;
                    if (lstm) {
                        const float ig = tt_sigmoid(gb[gi0] + xb[h]);
                        const float fg = tt_sigmoid(gb[gi1] + xb[H + h]);
                        const float gg = tanhf(gb[gi2] + xb[2 * H + h]);
                        const float og = tt_sigmoid(gb[gi3] + xb[3 * H + h]);
                        const float cprev = cps[b * H + h];
                        const float cnew = fg * cprev + ig * gg;
                        const float tc = tanhf(cnew);
                        const float dc = dcs[b * H + h] + dh * og * (1.0f - tc * tc);
                        if (a.dc_log) a.dc_log[(row * a.T + tg) * H + h] = dc;
                        const float d_i = dc * gg * ig * (1.0f - ig);
                        const float d_f = dc * cprev * fg * (1.0f - fg);
                        const float d_g = dc * ig * (1.0f - gg * gg);
                        const float d_o = dh * tc * og * (1.0f - og);
                        dcs[b * H + h] = dc * fg;
                        dhd[b * H + h] = 0.f;
                        gb[gi0] = d_i; gb[gi1] = d_f; gb[gi2] = d_g; gb[gi3] = d_o;
                        dxg[h] = d_i; dxg[H + h] = d_f; dxg[2 * H + h] = d_g; dxg[3 * H + h] = d_o;
                    } else {
                        const float ur = gb[gi0] + bsm[h];
                        const float uz = gb[gi1] + bsm[H + h];
                        const float un = gb[gi2] + bsm[2 * H + h];
                        const float rg = tt_sigmoid(xb[h] + ur);
                        const float zg = tt_sigmoid(xb[H + h] + uz);
                        const float ng = tanhf(xb[2 * H + h] + rg * un);
                        const float hprev = hslot[b * sl.BS + hix];
                        const float dn = dh * (1.0f - zg);
                        const float d_n = dn * (1.0f - ng * ng);
                        const float d_z = dh * (hprev - ng) * zg * (1.0f - zg);
                        const float d_r = d_n * un * rg * (1.0f - rg);
                        const float d_nh = d_n * rg;
                        dhd[b * H + h] = dh * zg;
                        gb[gi0] = d_r; gb[gi1] = d_z; gb[gi2] = d_nh;
                        dxg[h] = d_r; dxg[H + h] = d_z; dxg[2 * H + h] = d_n;
                        db0 += d_r; db1 += d_z; db2 += d_nh;
                    }
                }
                if (!lstm) { dbs[h] += db0; dbs[H + h] += db1; dbs[2 * H + h] += db2; }
            }
            __syncthreads();
            // ---- backward chain: core gradients and dh_{t-1}
            tt_chain_bwd(p, a.tile_bd, a.mg, R, xa, spill, g, dhc, wsm, dws, true, tid);
        }
        for (int e = tid; e < R * H; e += TT_NTHREADS) {
            const int b = e / H, h = e - b * H;
            const long long row = row0 + b;
            if (row < a.B) {
                a.dh_out[row * H + h] = dhd[e] + dhc[b * sl.BS + tt_in_index(p, h)];
                if (lstm) a.dc_out[row * H + h] = dcs[e];
            }
        }
    }
    __syncthreads();
    tt_flush_partial(p, dws, lstm ? nullptr : dbs, GH, a.partial + (long long)blockIdx.x * (p.core_floats + GH), tid);
}

// out[e] = sum over slots of partial[slot][e]   (fixed order: deterministic)
__global__ void k_reduce_partials(const float *__restrict__ partial, int nslots, int n, long long slot_stride,
                                  float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
    for (int i = 0; i < nslots; ++i) s += partial[(long long)i * slot_stride + e];
    out[e] = s;
}

// Per-step batch statistics of a (B,T,H) tensor, the two quantities ActivGradLogger keeps (reference rnn_utils.py:217-226):
//   out[t] = mean_b ||v[b,t,:]||^2,   out[T + t] = mean_b log ||v[b,t,:]||^2.   One CTA per timestep, one warp per row.
__global__ void __launch_bounds__(256) k_step_norms(const float *__restrict__ v, long long B, int T, int H,
                                                    float *__restrict__ out) {
    __shared__ float s_sum[8], s_log[8];
    const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f, acc_log = 0.f;
    for (long long b = warp; b < B; b += 8) {
        const float *row = v + (b * T + t) * (long long)H;
        float q = 0.f;
        for (int h = lane; h < H; h += 32) q = fmaf(row[h], row[h], q);
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        acc += q;
        acc_log += logf(q);
    }
    if (lane == 0) { s_sum[warp] = acc; s_log[warp] = acc_log; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, l = 0.f;
        for (int w = 0; w < 8; ++w) { a += s_sum[w]; l += s_log[w]; }
        out[t] = a / (float)B;
        out[T + t] = l / (float)B;
    }
}

__global__ void k_fill(float *dst, float v, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dst[e] = v;
}

// dst[e] (+)= src[e]
__global__ void k_axpy1(const float *__restrict__ src, float *__restrict__ dst, long long n, int accumulate) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dst[e] = accumulate ? dst[e] + src[e] : src[e];
}
