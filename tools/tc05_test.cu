// Stand-alone correctness + timing test of the tcgen05 convolution against the mma.sync kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/tc05_test tools/tc05_test.cu
//   timeout 120 /tmp/tc05_test [/tmp/nbr.bin] [max_rows]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../pcgcv2_b200/csrc/conv_mma.cuh"
#include "../pcgcv2_b200/csrc/conv_tc05.cuh"

namespace pcgc { void set_error(const char *, ...) {} std::atomic<uint64_t> g_launches{0}; }
using namespace pcgc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int CIN, int NPAD, int COUT>
void test(const float *in, const int32_t *nbr, int64_t n, int64_t pairs, const float *w, const float *bias, float *flush) {
    using C = tc05::Cfg<CIN, NPAD>;
    float *packed_t, *packed_m, *out_t, *out_m;
    CK(cudaMalloc(&packed_t, C::packed_floats() * 4));
    CK(cudaMalloc(&packed_m, MmaCfg<CIN, COUT>::packed_floats() * 4));
    CK(cudaMalloc(&out_t, n * COUT * 4)); CK(cudaMalloc(&out_m, n * COUT * 4));
    CK(cudaMemset(out_t, 0xFF, n * COUT * 4));
    tc05::pack_weights_tc05_kernel<CIN, NPAD><<<64, 256>>>(w, COUT, packed_t);
    pack_weights_mma_kernel<<<64, 256>>>(w, 27, CIN, COUT, packed_m);
    // reference: mma.sync kernel
    {
        using M = MmaCfg<CIN, COUT>;
        auto kern = conv_k3_mma_kernel<CIN, COUT>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M::smem_bytes()));
        kern<<<296, 256, M::smem_bytes()>>>(in, CIN, nbr, n, packed_m, bias, nullptr, 0, out_m, COUT, 1);
        CK(cudaDeviceSynchronize());
    }
    auto kern = tc05::conv_k3_tc05_kernel<CIN, NPAD, COUT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    const int64_t tiles = (n + 127) / 128;
    const int grid = (int)std::min<int64_t>(tiles, 148);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
        CK(cudaMemsetAsync(flush, it, 256u << 20));
        cudaEventRecord(a);
        kern<<<grid, C::THREADS, C::SMEM>>>(in, CIN, nbr, n, packed_t, bias, nullptr, 0, out_t, COUT, 1);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it) best = std::min(best, ms);
    }
    std::vector<float> ht(n * COUT), hm(n * COUT);
    CK(cudaMemcpy(ht.data(), out_t, n * COUT * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hm.data(), out_m, n * COUT * 4, cudaMemcpyDeviceToHost));
    double maxd = 0, maxr = 0; int64_t bad = 0;
    for (int64_t i = 0; i < n * COUT; ++i) {
        const double d = std::fabs((double)ht[i] - hm[i]);
        if (!(d <= 1e30)) { ++bad; continue; }
        maxd = std::max(maxd, d); maxr = std::max(maxr, (double)std::fabs(hm[i]));
    }
    const double alg = 4.0 * n * (CIN + COUT) + 8.0 * pairs + 4.0 * 27 * CIN * COUT;
    printf("tcgen05 %dx%d (N=%d): rows %lld  %.4f ms  %.0f GB/s alg  max|d| %.3e  max|ref| %.3e  rel %.2e  nan/inf %lld  smem %zu\n", CIN,
           COUT, NPAD, (long long)n, best, alg / best / 1e6, maxd, maxr, maxd / maxr, (long long)bad, (size_t)C::SMEM);
    cudaFree(packed_t); cudaFree(packed_m); cudaFree(out_t); cudaFree(out_m);
}

template <int NPAD, int COUT>
void test_t(const float *in, const int32_t *nbr, int64_t n, int64_t pairs, const float *w, const float *bias, float *flush) {
    constexpr int CIN = 16;
    using C = tc05::CfgT<NPAD>;
    float *packed_t, *packed_m, *out_t, *out_m;
    CK(cudaMalloc(&packed_t, 27 * 2 * NPAD * CIN * 4));
    CK(cudaMalloc(&packed_m, MmaCfg<CIN, COUT>::packed_floats() * 4));
    CK(cudaMalloc(&out_t, n * COUT * 4)); CK(cudaMalloc(&out_m, n * COUT * 4));
    CK(cudaMemset(out_t, 0xFF, n * COUT * 4));
    tc05::pack_weights_tc05t_kernel<NPAD><<<64, 256>>>(w, COUT, packed_t);
    pack_weights_mma_kernel<<<64, 256>>>(w, 27, CIN, COUT, packed_m);
    {
        using M = MmaCfg<CIN, COUT>;
        auto kern = conv_k3_mma_kernel<CIN, COUT>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M::smem_bytes()));
        kern<<<296, 256, M::smem_bytes()>>>(in, CIN, nbr, n, packed_m, bias, nullptr, 0, out_m, COUT, 1);
        CK(cudaDeviceSynchronize());
    }
    auto kern = tc05::conv_k3_tc05t_kernel<NPAD, COUT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    const int64_t tiles = (n + 127) / 128;
    const int grid = (int)std::min<int64_t>(tiles, 148);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
        CK(cudaMemsetAsync(flush, it, 256u << 20));
        cudaEventRecord(a);
        kern<<<grid, C::THREADS, C::SMEM>>>(in, CIN, nbr, n, packed_t, bias, nullptr, 0, out_t, COUT, 1);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it) best = std::min(best, ms);
    }
    std::vector<float> ht(n * COUT), hm(n * COUT);
    CK(cudaMemcpy(ht.data(), out_t, n * COUT * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hm.data(), out_m, n * COUT * 4, cudaMemcpyDeviceToHost));
    double maxd = 0, maxr = 0; int64_t bad = 0;
    for (int64_t i = 0; i < n * COUT; ++i) {
        const double d = std::fabs((double)ht[i] - hm[i]);
        if (!(d <= 1e30)) { ++bad; continue; }
        maxd = std::max(maxd, d); maxr = std::max(maxr, (double)std::fabs(hm[i]));
    }
    const double alg = 4.0 * n * (CIN + COUT) + 8.0 * pairs + 4.0 * 27 * CIN * COUT;
    printf("tcgen05-TMEM-A 16x%d: rows %lld  %.4f ms  %.0f GB/s alg  max|d| %.3e  max|ref| %.3e  rel %.2e  nan/inf %lld  threads %d\n", COUT,
           (long long)n, best, alg / best / 1e6, maxd, maxr, maxd / maxr, (long long)bad, C::THREADS);
    cudaFree(packed_t); cudaFree(packed_m); cudaFree(out_t); cudaFree(out_m);
}

int main(int argc, char **argv) {
    int64_t n = 0, pairs = 0;
    std::vector<int32_t> h;
    FILE *f = fopen(argc > 1 ? argv[1] : "/tmp/nbr.bin", "rb");
    if (f) {
        if (fread(&n, 8, 1, f) != 1 || fread(&pairs, 8, 1, f) != 1) return 1;
        h.resize(27 * n);
        if (fread(h.data(), 4, 27 * n, f) != (size_t)(27 * n)) return 1;
        fclose(f);
    } else {                                                  // synthetic map: random neighbours, 70 % present
        n = 100000;
        h.resize(27 * n);
        for (int64_t i = 0; i < 27 * n; ++i) h[i] = (rand() % 10 < 7) ? rand() % n : -1;
        for (int64_t u = 0; u < n; ++u) h[13 * n + u] = (int32_t)u;
        for (auto v : h) pairs += v >= 0;
    }
    if (argc > 2) {                                           // restrict to the first max_rows rows (map entries beyond are dropped)
        const int64_t m = std::min<int64_t>(n, atoll(argv[2]));
        std::vector<int32_t> h2(27 * m);
        pairs = 0;
        for (int k = 0; k < 27; ++k)
            for (int64_t u = 0; u < m; ++u) { int32_t v = h[k * n + u]; if (v >= m) v = -1; h2[k * m + u] = v; pairs += v >= 0; }
        h.swap(h2); n = m;
    }
    printf("rows %lld pairs %lld\n", (long long)n, (long long)pairs);
    int32_t *nbr; CK(cudaMalloc(&nbr, 27 * n * 4)); CK(cudaMemcpy(nbr, h.data(), 27 * n * 4, cudaMemcpyHostToDevice));
    float *flush; CK(cudaMalloc(&flush, 256u << 20));
    constexpr int MAXC = 64;
    float *in, *w, *bias;
    CK(cudaMalloc(&in, n * MAXC * 4)); CK(cudaMalloc(&w, 27 * MAXC * MAXC * 4)); CK(cudaMalloc(&bias, MAXC * 4));
    std::vector<float> hin(n * MAXC);
    for (auto &v : hin) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(in, hin.data(), n * MAXC * 4, cudaMemcpyHostToDevice));
    std::vector<float> hw(27 * MAXC * MAXC);
    for (auto &v : hw) v = ((float)rand() / RAND_MAX - 0.5f) * 0.1f;
    CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> hb(MAXC);
    for (auto &v : hb) v = (float)rand() / RAND_MAX;
    CK(cudaMemcpy(bias, hb.data(), MAXC * 4, cudaMemcpyHostToDevice));
    if (!getenv("TC05_SKIP_V1")) test<16, 16, 16>(in, nbr, n, pairs, w, bias, flush);
    test_t<16, 16>(in, nbr, n, pairs, w, bias, flush);
    return 0;
}
