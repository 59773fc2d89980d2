// GE2E head on the device: what the reference computes on the CPU after every RNN step
// (experiments/speaker_verification/encoder/speaker_encoder.py:86-89 ReLU + L2 norm, :93-141 similarity
// matrix with inclusive / exclusive centroids, :143-156 softmax cross-entropy; the reference forces this onto
// the CPU at encoder/main.py:279-280, i.e. a D2H -> CPU loss -> H2D round trip per training step).
//
// Sizes are tiny (64 speakers x 10 utterances x 256): the kernels are organised for determinism and few launches,
// not for a roofline: one pass normalises, one builds the centroids, one computes similarity rows + softmax + the
// per-row loss, and the backward mirrors them.  All reductions run in a fixed order (no atomics).
#pragma once
#include <cuda_runtime.h>

namespace ge2e {

constexpr int NT = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of one value per thread (NT threads); `red` = 8 floats of shared memory; result broadcast
__device__ __forceinline__ float block_sum(float v, float *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    return t;
}

// ---- embeds = relu(x) / ||relu(x)||_2 per row (speaker_encoder.py:86-89); one warp per row ---------------------
__global__ void __launch_bounds__(NT) k_embed_fwd(const float *__restrict__ x, float *__restrict__ y,
                                                  float *__restrict__ inv_norm, int rows, int E) {
    const int row = blockIdx.x * (NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *xr = x + (long long)row * E;
    float ss = 0.f;
    for (int e = lane; e < E; e += 32) {
        const float r = fmaxf(xr[e], 0.f);
        ss += r * r;
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss);
    for (int e = lane; e < E; e += 32) y[(long long)row * E + e] = fmaxf(xr[e], 0.f) * inv;
    if (lane == 0) inv_norm[row] = inv;
}
// dx = (dy - y (y . dy)) * inv_norm * [x > 0]
__global__ void __launch_bounds__(NT) k_embed_bwd(const float *__restrict__ x, const float *__restrict__ y,
                                                  const float *__restrict__ inv_norm, const float *__restrict__ dy,
                                                  float *__restrict__ dx, int rows, int E) {
    const int row = blockIdx.x * (NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long o = (long long)row * E;
    float dot = 0.f;
    for (int e = lane; e < E; e += 32) dot += y[o + e] * dy[o + e];
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    for (int e = lane; e < E; e += 32) dx[o + e] = (x[o + e] > 0.f) ? (dy[o + e] - y[o + e] * dot) * inv : 0.f;
}

// ---- centroids (speaker_encoder.py:109-117): one block per speaker ------------------------------------------------
// c_incl[s] = normalize(mean_u e[s,u]);  c_excl[s,u] = normalize((sum_u e[s,u] - e[s,u]) / (U - 1))
__global__ void __launch_bounds__(NT) k_centroids(const float *__restrict__ emb, float *__restrict__ c_incl,
                                                  float *__restrict__ n_incl, float *__restrict__ c_excl,
                                                  float *__restrict__ n_excl, int U, int E) {
    extern __shared__ float sm[];                 // sum[E] + 8
    float *sum = sm, *red = sm + E;
    const int s = blockIdx.x;
    const float *es = emb + (long long)s * U * E;
    float ss = 0.f;
    for (int e = threadIdx.x; e < E; e += NT) {
        float t = 0.f;
        for (int u = 0; u < U; ++u) t += es[(long long)u * E + e];
        sum[e] = t;
        const float m = t / (float)U;
        ss += m * m;
    }
    ss = block_sum(ss, red);
    const float ni = sqrtf(ss);
    for (int e = threadIdx.x; e < E; e += NT) c_incl[(long long)s * E + e] = (sum[e] / (float)U) / ni;
    if (threadIdx.x == 0) n_incl[s] = ni;
    for (int u = 0; u < U; ++u) {
        float q = 0.f;
        for (int e = threadIdx.x; e < E; e += NT) {
            const float m = (sum[e] - es[(long long)u * E + e]) / (float)(U - 1);
            q += m * m;
        }
        q = block_sum(q, red);
        const float ne = sqrtf(q);
        for (int e = threadIdx.x; e < E; e += NT)
            c_excl[((long long)s * U + u) * E + e] = ((sum[e] - es[(long long)u * E + e]) / (float)(U - 1)) / ne;
        if (threadIdx.x == 0) n_excl[s * U + u] = ne;
    }
}

// ---- similarity rows + softmax cross-entropy (speaker_encoder.py:121-141, 152-156): one block per utterance ---------
// sim[(s,u), j] = e[s,u] . c_incl[j] (j != s) | e[s,u] . c_excl[s,u] (j == s);  logits = w * sim + b
// prob = softmax(logits);  row_loss = -log prob[s]
__global__ void __launch_bounds__(NT) k_sim_loss(const float *__restrict__ emb, const float *__restrict__ c_incl,
                                                 const float *__restrict__ c_excl, const float *__restrict__ wb,
                                                 float *__restrict__ sim, float *__restrict__ prob,
                                                 float *__restrict__ row_loss, float *__restrict__ logits_out, int S, int U,
                                                 int E) {
    extern __shared__ float sm[];                 // e[E] + logits[S]
    float *ev = sm, *lg = sm + E;
    const int row = blockIdx.x, s = row / U;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < E; e += NT) ev[e] = emb[(long long)row * E + e];
    __syncthreads();
    const float w = wb[0], b = wb[1];
    for (int j = warp; j < S; j += NT / 32) {
        const float *c = (j == s) ? c_excl + (long long)row * E : c_incl + (long long)j * E;
        float d = 0.f;
        for (int e = lane; e < E; e += 32) d += ev[e] * c[e];
        d = warp_sum(d);
        if (lane == 0) {
            sim[(long long)row * S + j] = d;
            lg[j] = d * w + b;
            if (logits_out) logits_out[(long long)row * S + j] = d * w + b;
        }
    }
    __syncthreads();
    // softmax over S logits (S is small: every thread scans them)
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) mx = fmaxf(mx, lg[j]);
    float den = 0.f;
    for (int j = 0; j < S; ++j) den += expf(lg[j] - mx);
    for (int j = threadIdx.x; j < S; j += NT) prob[(long long)row * S + j] = expf(lg[j] - mx) / den;
    if (threadIdx.x == 0) row_loss[row] = -(lg[s] - mx - logf(den));
}

// mean of n values in a fixed order (one block)
__global__ void __launch_bounds__(NT) k_mean(const float *__restrict__ v, int n, float *__restrict__ out) {
    __shared__ float red[8];
    float t = 0.f;
    for (int i = threadIdx.x; i < n; i += NT) t += v[i];
    t = block_sum(t, red);
    if (threadIdx.x == 0) out[0] = t / (float)n;
}

// ---- backward, rows: dlogits = (prob - onehot) * g / B;  dsim = dlogits * w;  per-row dw / db partials;
//      de[row] = sum_{j != s} dsim[j] c_incl[j] + dsim[s] c_excl[row]        (direct term)
__global__ void __launch_bounds__(NT) k_bwd_rows(const float *__restrict__ c_incl, const float *__restrict__ c_excl,
                                                 const float *__restrict__ wb, const float *__restrict__ sim,
                                                 const float *__restrict__ prob, const float *__restrict__ gscale,
                                                 float *__restrict__ dsim, float *__restrict__ de,
                                                 float *__restrict__ row_dw, float *__restrict__ row_db, int S, int U, int E) {
    extern __shared__ float sm[];                 // ds[S]
    float *ds = sm;
    const int row = blockIdx.x, s = row / U;
    const float g = gscale[0] / (float)(S * U), w = wb[0];
    for (int j = threadIdx.x; j < S; j += NT) {
        const float dl = (prob[(long long)row * S + j] - (j == s ? 1.f : 0.f)) * g;
        ds[j] = dl * w;
        dsim[(long long)row * S + j] = dl * w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int j = 0; j < S; ++j) {
            const float dl = (prob[(long long)row * S + j] - (j == s ? 1.f : 0.f)) * g;
            a += dl * sim[(long long)row * S + j];
            c += dl;
        }
        row_dw[row] = a;
        row_db[row] = c;
    }
    for (int e = threadIdx.x; e < E; e += NT) {
        float t = 0.f;
        for (int j = 0; j < S; ++j) {
            const float cj = (j == s) ? c_excl[(long long)row * E + e] : c_incl[(long long)j * E + e];
            t += ds[j] * cj;
        }
        de[(long long)row * E + e] = t;
    }
}

// ---- backward, centroids: one block per speaker j.  Adds the centroid paths to de of the rows of speaker j. --------
//   dc_incl[j] = sum_{rows of other speakers} dsim[row, j] e[row];  dm = (dc - c (c . dc)) / n;  de[j,u] += dm / U
//   dc_excl[j,u] = dsim[(j,u), j] e[j,u];  dm_u = (dc - c (c . dc)) / n_excl;  de[j,u'] += sum_{u != u'} dm_u / (U - 1)
__global__ void __launch_bounds__(NT) k_bwd_centroids(const float *__restrict__ emb, const float *__restrict__ c_incl,
                                                      const float *__restrict__ n_incl, const float *__restrict__ c_excl,
                                                      const float *__restrict__ n_excl, const float *__restrict__ dsim,
                                                      float *__restrict__ de, int S, int U, int E) {
    extern __shared__ float sm[];                 // dmsum[E] + 8
    float *dmsum = sm, *red = sm + E;
    const int j = blockIdx.x;
    const int rows = S * U;
    // inclusive centroid
    float dot = 0.f;
    for (int e = threadIdx.x; e < E; e += NT) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) {
            if (r / U == j) continue;
            t += dsim[(long long)r * S + j] * emb[(long long)r * E + e];
        }
        dmsum[e] = t;                             // dc_incl for now
        dot += t * c_incl[(long long)j * E + e];
    }
    dot = block_sum(dot, red);
    const float ni = n_incl[j];
    for (int e = threadIdx.x; e < E; e += NT) {
        const float dm = (dmsum[e] - c_incl[(long long)j * E + e] * dot) / ni / (float)U;
        for (int u = 0; u < U; ++u) de[((long long)j * U + u) * E + e] += dm;
    }
    __syncthreads();
    // exclusive centroids: accumulate sum_u dm_u, then de[j,u'] += (sum - dm_u') / (U - 1)
    for (int e = threadIdx.x; e < E; e += NT) dmsum[e] = 0.f;
    for (int u = 0; u < U; ++u) {
        const long long ro = ((long long)j * U + u) * E;
        const float d = dsim[((long long)j * U + u) * S + j];
        float q = 0.f;
        for (int e = threadIdx.x; e < E; e += NT) q += d * emb[ro + e] * c_excl[ro + e];
        q = block_sum(q, red);
        const float ne = n_excl[j * U + u];
        for (int e = threadIdx.x; e < E; e += NT) dmsum[e] += (d * emb[ro + e] - c_excl[ro + e] * q) / ne;
    }
    for (int u = 0; u < U; ++u) {
        const long long ro = ((long long)j * U + u) * E;
        const float d = dsim[((long long)j * U + u) * S + j];
        float q = 0.f;
        for (int e = threadIdx.x; e < E; e += NT) q += d * emb[ro + e] * c_excl[ro + e];
        q = block_sum(q, red);
        const float ne = n_excl[j * U + u];
        for (int e = threadIdx.x; e < E; e += NT) {
            const float dmu = (d * emb[ro + e] - c_excl[ro + e] * q) / ne;
            de[ro + e] += (dmsum[e] - dmu) / (float)(U - 1);
        }
    }
}

// sum of n values in a fixed order (one block)
__global__ void __launch_bounds__(NT) k_sum(const float *__restrict__ v, int n, float *__restrict__ out) {
    __shared__ float red[8];
    float t = 0.f;
    for (int i = threadIdx.x; i < n; i += NT) t += v[i];
    t = block_sum(t, red);
    if (threadIdx.x == 0) out[0] = t;
}

}  // namespace ge2e
