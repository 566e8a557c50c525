// Micro-benchmark: times template variants of conv_k3_mma_kernel on a kernel map dumped by
// tools/profile_conv.py --dump (real vox10 decoder level).  Build: see tools/run_variants.sh
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pcgcv2_b200/csrc/conv_mma.cuh"

namespace pcgc { void set_error(const char *, ...) {} std::atomic<uint64_t> g_launches{0}; }
using namespace pcgc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static float *g_flush;
template <int CIN, int COUT, int RG, int B, int MINB, int RES>
void run(const char *name, const float *in, const int32_t *nbr, int64_t n, const float *packed, const float *bias, float *out,
         double alg_bytes, double flops) {
    using C = MmaCfg<CIN, COUT, RG, B, RES>;
    auto kern = conv_k3_mma_kernel<CIN, COUT, RG, B, MINB, RES>;
    size_t smem = C::smem_bytes();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int64_t tiles = (n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    int grid = (int)std::min<int64_t>(tiles, (int64_t)148 * nb);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
        CK(cudaMemsetAsync(g_flush, it, 256u << 20));
        cudaEventRecord(a);
        kern<<<grid, 256, smem>>>(in, CIN, nbr, n, packed, bias, nullptr, 0, out, COUT, 1);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0) best = std::min(best, ms);
    }
    printf("%-28s regs %3d  ctas/SM %d  smem %6zu  %.4f ms  %.0f GB/s alg  %.1f TFLOP/s\n", name, fa.numRegs, nb, smem, best,
           alg_bytes / best / 1e6, flops / best / 1e9);
}

int main(int argc, char **argv) {
    FILE *f = fopen(argc > 1 ? argv[1] : "/tmp/nbr.bin", "rb");
    if (!f) { printf("no map dump\n"); return 1; }
    int64_t n, pairs;
    fread(&n, 8, 1, f); fread(&pairs, 8, 1, f);
    std::vector<int32_t> h(27 * n);
    fread(h.data(), 4, 27 * n, f); fclose(f);
    printf("rows %lld pairs %lld\n", (long long)n, (long long)pairs);
    int32_t *nbr; CK(cudaMalloc(&nbr, 27 * n * 4)); CK(cudaMemcpy(nbr, h.data(), 27 * n * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&g_flush, 256u << 20));
    constexpr int MAXC = 64;
    float *in, *out, *w, *packed, *bias;
    CK(cudaMalloc(&in, n * MAXC * 4)); CK(cudaMalloc(&out, n * MAXC * 4));
    CK(cudaMalloc(&w, 27 * MAXC * MAXC * 4)); CK(cudaMalloc(&packed, 27 * MAXC * MAXC * 8 + 1024)); CK(cudaMalloc(&bias, MAXC * 4));
    std::vector<float> hin(n * MAXC);
    for (auto &v : hin) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(in, hin.data(), n * MAXC * 4, cudaMemcpyHostToDevice));
    std::vector<float> hw(27 * MAXC * MAXC);
    for (auto &v : hw) v = ((float)rand() / RAND_MAX - 0.5f) * 0.1f;
    CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(bias, 0, MAXC * 4));
#define RUN(CIN, COUT, RG, B, MINB, RES)                                                                   \
    pack_weights_mma_kernel<<<64, 256>>>(w, 27, CIN, COUT, packed);                                        \
    run<CIN, COUT, RG, B, MINB, RES>(#CIN "x" #COUT " RG" #RG " B" #B " minb" #MINB " res" #RES, in, nbr, n, packed, bias, out, \
                               4.0 * n * (CIN + COUT) + 8.0 * pairs + 4.0 * 27 * CIN * COUT, 2.0 * pairs * CIN * COUT);
#include "/tmp/variants.h"
    return 0;
}
