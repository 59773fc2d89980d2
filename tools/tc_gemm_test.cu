// Stand-alone check + timing of the tcgen05 3xTF32 GEMMs (csrc/tt_tc.cuh) against an FP64 CPU product and the
// FP32 FFMA kernels (csrc/tt_gemm.cuh).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/tc_gemm_test.cu -o tools/tc_gemm_test
// Run on a B200:  tools/tc_gemm_test [quick]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tensorized_rnn_b200/csrc/tt_gemm.cuh"
#include "../tensorized_rnn_b200/csrc/tt_tc.cuh"

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
    } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

struct Err { double rel, maxabs; };
static Err compare(const std::vector<float> &got, const std::vector<double> &ref) {
    double num = 0, den = 0, mx = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        const double d = got[i] - ref[i];
        num += d * d; den += ref[i] * ref[i];
        if (fabs(d) > mx) mx = fabs(d);
    }
    return {sqrt(num / (den > 0 ? den : 1)), mx};
}

// rows = nb * rpb ragged views into a (nb, T, width) tensor
static int test_rows(int nb, int rpb, int T, int K, int N, bool bias, int sms, bool timing) {
    const long long rows = (long long)nb * rpb;
    std::vector<float> A((size_t)nb * T * K), Bt((size_t)N * K), bv(N), C((size_t)nb * T * N, -7.f);
    for (auto &v : A) v = frand();
    for (auto &v : Bt) v = frand() * 0.3f;
    for (auto &v : bv) v = frand();
    float *dA, *dBt, *dBh, *dBl, *dC, *db, *dBkn;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBt, Bt.size() * 4)); CK(cudaMalloc(&dBh, Bt.size() * 4));
    CK(cudaMalloc(&dBl, Bt.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&db, N * 4));
    CK(cudaMalloc(&dBkn, Bt.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBt, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bv.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice));
    ttc::k_split_tf32<<<64, 256>>>(dBt, dBh, dBl, (long long)N * K / 4);
    CK(cudaGetLastError());
    int rc = ttc::launch_tc_rows(rows, rpb, dA, (long long)T * K, K, dBh, dBl, N, bias ? db : nullptr, nullptr, dC, (long long)T * N, sms, 0);
    if (rc) { printf("launch_tc_rows rc=%d\n", rc); return 1; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    Err e{0, 0};
    if (!timing) {
        std::vector<double> ref;
        std::vector<float> got;
        for (int b = 0; b < nb; ++b)
            for (int t = 0; t < T; ++t)
                for (int n = 0; n < N; ++n) {
                    const float g = C[((size_t)b * T + t) * N + n];
                    if (t >= rpb) { if (g != -7.f) ++bad; continue; }      // rows outside the view must be untouched
                    double s = bias ? bv[n] : 0.0;
                    for (int k = 0; k < K; ++k) s += (double)A[((size_t)b * T + t) * K + k] * Bt[(size_t)n * K + k];
                    ref.push_back(s); got.push_back(g);
                }
        e = compare(got, ref);
    }
    float ms = 0, ms_ffma = 0;
    if (timing) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int w = 0; w < 2; ++w) ttc::launch_tc_rows(rows, rpb, dA, (long long)T * K, K, dBh, dBl, N, nullptr, nullptr, dC, (long long)T * N, sms, 0);
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 5; ++w) ttc::launch_tc_rows(rows, rpb, dA, (long long)T * K, K, dBh, dBl, N, nullptr, nullptr, dC, (long long)T * N, sms, 0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
#ifdef TC_PROBE
        {
            long long h[8];
            CK(cudaMemcpy(h, ttc::tc_probe_buf(), sizeof h, cudaMemcpyDeviceToHost));
            printf("  (next line) MMA issuer of CTA 3: total %lld clk, %lld k-blocks (%.0f clk each); waits: tempty %lld full %lld aready %lld sfree %lld; issue %lld\n",
                   h[0], h[6], (double)h[0] / (double)h[6], h[1], h[2], h[3], h[4], h[5]);
        }
#endif
        // FFMA reference kernel: needs B as K x N
        dim3 tg((K + 31) / 32, (N + 31) / 32);
        ttg::k_transpose<<<tg, 256>>>(dBt, dBkn, N, K);
        ttg::GemmRowsArgs g; memset(&g, 0, sizeof g);
        g.rows = (unsigned)rows; g.K = K; g.N = N;
        g.a.p = dA; g.a.bstride = (long long)T * K; g.a.rpb = rpb; g.a.ld = K;
        g.b = dBkn; g.ldb = N; g.c.p = dC; g.c.bstride = (long long)T * N; g.c.rpb = rpb; g.c.ld = N;
        const long long tiles = ((rows + 127) / 128) * (N / 128);
        for (int w = 0; w < 2; ++w) ttg::k_gemm_rows<8, 2><<<(unsigned)tiles, 256>>>(g);
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 3; ++w) ttg::k_gemm_rows<8, 2><<<(unsigned)tiles, 256>>>(g);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_ffma, e0, e1)); ms_ffma /= 3;
    }
    const double fl = 2.0 * rows * K * N;
    if (timing)
        printf("rows  nb=%d rpb=%d T=%d K=%d N=%d : tc %.3f ms = %.1f TFLOP/s | ffma %.3f ms = %.1f TFLOP/s\n", nb, rpb, T, K, N, ms,
               fl / ms / 1e9, ms_ffma, fl / ms_ffma / 1e9);
    else
        printf("rows  nb=%d rpb=%d T=%d K=%d N=%d bias=%d : rel %.3e maxabs %.3e untouched-violations %d %s\n", nb, rpb, T, K, N, (int)bias,
               e.rel, e.maxabs, bad, (e.rel < 6e-6 && bad == 0) ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dBt); cudaFree(dBh); cudaFree(dBl); cudaFree(dC); cudaFree(db); cudaFree(dBkn);
    return (timing || (e.rel < 6e-6 && bad == 0)) ? 0 : 1;
}

static int test_red(int nb, int rpb, int T, int M, int N, int sms, bool timing, int max_split = 40) {
    const long long rows = (long long)nb * rpb;
    std::vector<float> A((size_t)nb * T * M), B((size_t)nb * T * N);
    for (auto &v : A) v = frand();
    for (auto &v : B) v = frand();
    float *dA, *dB, *dP, *dPb, *dOut, *dOb;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dP, (size_t)max_split * M * N * 4)); CK(cudaMalloc(&dPb, (size_t)max_split * N * 4));
    CK(cudaMalloc(&dOut, (size_t)M * N * 4)); CK(cudaMalloc(&dOb, N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    int nsplit = 0;
    int rc = ttc::launch_tc_red(rows, rpb, dA, (long long)T * M, M, dB, (long long)T * N, N, dP, dPb, sms, max_split, &nsplit, 0);
    if (rc) { printf("launch_tc_red rc=%d\n", rc); return 1; }
    const long long wn = (long long)M * N;
    ttg::k_sum_splits<<<(unsigned)((wn / 4 + 255) / 256), 256>>>(dP, nsplit, wn, wn, dOut, 0);
    ttg::k_sum_splits<<<(unsigned)((N / 4 + 255) / 256), 256>>>(dPb, nsplit, N, N, dOb, 0);
    CK(cudaDeviceSynchronize());
    Err e{0, 0}, eb{0, 0};
    if (!timing) {
        std::vector<float> out((size_t)M * N), ob(N);
        CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ob.data(), dOb, N * 4, cudaMemcpyDeviceToHost));
        std::vector<double> ref((size_t)M * N, 0.0), rb(N, 0.0);
        for (int b = 0; b < nb; ++b)
            for (int t = 0; t < rpb; ++t) {
                const float *ar = &A[((size_t)b * T + t) * M], *br = &B[((size_t)b * T + t) * N];
                for (int m = 0; m < M; ++m) {
                    const double am = ar[m];
                    double *rr = &ref[(size_t)m * N];
                    for (int n = 0; n < N; ++n) rr[n] += am * br[n];
                }
                for (int n = 0; n < N; ++n) rb[n] += br[n];
            }
        e = compare(out, ref);
        eb = compare(ob, rb);
    }
    float ms = 0, ms_ffma = 0;
    if (timing) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int w = 0; w < 2; ++w) ttc::launch_tc_red(rows, rpb, dA, (long long)T * M, M, dB, (long long)T * N, N, dP, dPb, sms, max_split, &nsplit, 0);
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 5; ++w) ttc::launch_tc_red(rows, rpb, dA, (long long)T * M, M, dB, (long long)T * N, N, dP, dPb, sms, max_split, &nsplit, 0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
#ifdef TC_PROBE
        {
            long long h[8];
            CK(cudaMemcpy(h, ttc::tc_probe_buf(), sizeof h, cudaMemcpyDeviceToHost));
            printf("  (next line) MMA issuer of CTA 3: total %lld clk, %lld k-blocks (%.0f clk each); waits: aempty %lld ofull %lld; issue %lld\n",
                   h[0], h[6], (double)h[0] / (double)h[6], h[1], h[2], h[5]);
        }
#endif
        ttg::GemmRedArgs g; memset(&g, 0, sizeof g);
        const int tiles = ((M + 127) / 128) * (N / 128);
        long long ns = (4LL * sms) / tiles; if (ns < 1) ns = 1; if (ns > max_split) ns = max_split;
        g.rows = (unsigned)rows; g.M = M; g.N = N; g.nsplit = (int)ns;
        g.a.p = dA; g.a.bstride = (long long)T * M; g.a.rpb = rpb; g.a.ld = M;
        g.b.p = dB; g.b.bstride = (long long)T * N; g.b.rpb = rpb; g.b.ld = N;
        g.part = dP; g.pbias = dPb;
        for (int w = 0; w < 2; ++w) ttg::k_gemm_red<8><<<(unsigned)(tiles * ns), 256>>>(g);
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 3; ++w) ttg::k_gemm_red<8><<<(unsigned)(tiles * ns), 256>>>(g);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_ffma, e0, e1)); ms_ffma /= 3;
    }
    const double fl = 2.0 * rows * M * N;
    if (timing)
        printf("red   nb=%d rpb=%d T=%d M=%d N=%d nsplit=%d : tc %.3f ms = %.1f TFLOP/s | ffma %.3f ms = %.1f TFLOP/s\n", nb, rpb, T, M, N,
               nsplit, ms, fl / ms / 1e9, ms_ffma, fl / ms_ffma / 1e9);
    else
        printf("red   nb=%d rpb=%d T=%d M=%d N=%d nsplit=%d : rel %.3e maxabs %.3e | colsum rel %.3e %s\n", nb, rpb, T, M, N, nsplit, e.rel,
               e.maxabs, eb.rel, (e.rel < 2e-6 && eb.rel < 2e-6) ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dP); cudaFree(dPb); cudaFree(dOut); cudaFree(dOb);
    return (timing || (e.rel < 2e-6 && eb.rel < 2e-6)) ? 0 : 1;
}

int main(int argc, char **argv) {
#ifdef TC_PROBE
    CK(cudaMalloc(&ttc::tc_probe_buf(), 64));
    CK(cudaMemset(ttc::tc_probe_buf(), 0, 64));
#endif
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    if (getenv("TC_RED_VARIANT")) ttc::tc_red_variant() = atoi(getenv("TC_RED_VARIANT"));
    if (getenv("TC_ROWS_VARIANT")) ttc::tc_rows_variant() = atoi(getenv("TC_ROWS_VARIANT"));
    printf("k_tc_red variant: %s\n", ttc::tc_red_variant() ? "A from TMEM (k_tc_red_ts)" : "A from shared memory (k_tc_red)");
    if (argc > 1 && !strcmp(argv[1], "chain")) {
        int dev0 = 0, sms0 = 0;
        CK(cudaGetDevice(&dev0));
        CK(cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0));
        srand(99);
        const int ns[] = {1, 2, 3, 4, 5, 8, 16, 31, 32, 33, 34, 63, 64, 65, 96, 128, 512};
        for (int n : ns) test_red(1, 32 * n, 32 * n, 128, 128, sms0, false, 1);
        return 0;
    }
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    srand(1234);
    int fails = 0;
    // correctness: contiguous, ragged, K remainder, short boxes
    fails += test_rows(1, 128, 128, 32, 128, false, sms, false);
    fails += test_rows(1, 1000, 1000, 40, 256, true, sms, false);
    fails += test_rows(8, 64, 100, 256, 256, true, sms, false);
    fails += test_rows(20, 16, 40, 256, 128, false, sms, false);
    fails += test_rows(5, 160, 200, 72, 384, true, sms, false);
    fails += test_rows(64, 160, 160, 256, 1024, true, sms, false);
    fails += test_red(1, 256, 256, 128, 128, sms, false);
    fails += test_red(1, 1000, 1000, 256, 256, sms, false);
    fails += test_red(8, 63, 100, 256, 128, sms, false);
    fails += test_red(40, 16, 40, 128, 256, sms, false);
    fails += test_red(300, 1, 3, 256, 128, sms, false);
    fails += test_red(16, 160, 160, 72, 384, sms, false);
    // long accumulation chains: one split = 512 / 1024 k-blocks in one CTA (16 / 32 promoted TMEM segments)
    fails += test_red(1, 16384, 16384, 128, 128, sms, false, 1);
    fails += test_red(32, 1000, 1000, 128, 256, sms, false, 1);
    fails += test_rows(4, 200, 200, 1024, 256, false, sms, false);
    fails += test_rows(2, 300, 300, 2048, 128, true, sms, false);
    printf("correctness: %d failures\n", fails);
    if (!quick) {
        test_rows(640, 160, 160, 40, 1024, false, sms, true);      // cfg3 layer 0 ih
        test_rows(640, 160, 160, 256, 1024, false, sms, true);     // cfg3 layers 1-2 ih
        test_rows(640, 160, 160, 1024, 256, false, sms, true);     // cfg3 dX
        test_rows(4096, 64, 2000 > 64 ? 64 : 64, 256, 4096, false, sms, true);    // cfg5 ih chunk (contiguous stand-in)
        test_red(640, 160, 160, 256, 1024, sms, true);             // cfg3 dW (ih layers 1-2, hh)
        test_red(2048, 64, 64, 256, 4096, sms, true);              // cfg5 ih dW (half chunk)
        test_red(1024, 64, 64, 1024, 4096, sms, true);             // cfg5 hh dW (quarter chunk)
    }
    return fails ? 1 : 0;
}
