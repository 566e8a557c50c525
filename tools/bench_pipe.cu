// Micro-benchmark + correctness check of conv_k3_pipe_kernel variants against conv_k3_mma_kernel (the
// round-1 production kernel) on a kernel map dumped by tools/profile_conv.py --dump (real vox10 decoder
// level).  Build and run on the GPU box: tools/run_pipe.sh "RUNP(16,16,false,4,3,2,7) RUNP(16,4,true,2,3,2,7) ..."
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>
#include "../pcgcv2_b200/csrc/conv_pipe.cuh"

namespace pcgc { void set_error(const char *, ...) {} std::atomic<uint64_t> g_launches{0}; }
using namespace pcgc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static float *g_flush;
static int64_t g_n, g_pairs;
static float *g_in, *g_out, *g_ref, *g_w, *g_packed, *g_bias, *g_res;
static const int32_t *g_nbr;
static int g_ref_cin = -1, g_ref_cout = -1;
static std::vector<float> h_ref, h_out;

template <typename K, typename... Args>
static float time_kernel(K kern, int grid, size_t smem, Args... args) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
        CK(cudaMemsetAsync(g_flush, it, 256u << 20));
        cudaEventRecord(a);
        kern<<<grid, 256, smem>>>(args...);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

template <int CIN, int COUT>
static void reference() {                       // production kernel of round 1, output kept for the comparison
    if (g_ref_cin == CIN && g_ref_cout == COUT) return;
    using C = MmaCfg<CIN, COUT>;
    auto kern = conv_k3_mma_kernel<CIN, COUT>;
    size_t smem = C::smem_bytes();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem));
    pack_weights_mma_kernel<<<64, 256>>>(g_w, 27, CIN, COUT, g_packed);
    int64_t tiles = (g_n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    int grid = (int)std::min<int64_t>(tiles, (int64_t)148 * nb);
    float ms = time_kernel(kern, grid, smem, (const float *)g_in, CIN, g_nbr, g_n, (const float *)g_packed, (const float *)g_bias,
                           (const float *)g_res, COUT, g_ref, COUT, 1);
    h_ref.resize((size_t)g_n * COUT);
    CK(cudaMemcpy(h_ref.data(), g_ref, h_ref.size() * 4, cudaMemcpyDeviceToHost));
    g_ref_cin = CIN; g_ref_cout = COUT;
    printf("%-34s                               %.4f ms   (round-1 production kernel)\n",
           (std::to_string(CIN) + "x" + std::to_string(COUT) + " mma ref").c_str(), ms);
}

template <int CIN, int COUT, bool NT, int RG, int D, int MINB, int OPT>
static void run(const char *name) {
    reference<CIN, COUT>();
    using C = PipeCfg<CIN, COUT, NT, RG, D>;
    auto kern = conv_k3_pipe_kernel<CIN, COUT, NT, RG, D, MINB, OPT>;
    size_t smem = C::smem_bytes();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    if (NT) pack_weights_nt_kernel<<<64, 256>>>(g_w, 27, CIN, COUT, g_packed);
    else pack_weights_t_kernel<<<64, 256>>>(g_w, 27, CIN, COUT, g_packed);
    CK(cudaMemset(g_out, 0xff, (size_t)g_n * COUT * 4));
    int64_t tiles = (g_n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    int grid = (int)std::min<int64_t>(tiles, (int64_t)148 * nb);
    float ms = time_kernel(kern, grid, smem, (const float *)g_in, CIN, g_nbr, g_n, (const float *)g_packed, (const float *)g_bias,
                           (const float *)g_res, COUT, g_out, COUT, 1);
    h_out.resize((size_t)g_n * COUT);
    CK(cudaMemcpy(h_out.data(), g_out, h_out.size() * 4, cudaMemcpyDeviceToHost));
    double maxd = 0, maxr = 0;
    for (size_t i = 0; i < h_out.size(); ++i) {
        double d = fabs((double)h_out[i] - (double)h_ref[i]);
        if (!(d <= maxd)) maxd = d;                      // NaN-propagating max
        maxr = std::max(maxr, fabs((double)h_ref[i]));
    }
    const double alg = 4.0 * g_n * (CIN + COUT) + 8.0 * g_pairs + 4.0 * 27 * CIN * COUT, flops = 2.0 * g_pairs * CIN * COUT;
    printf("%-34s regs %3d ctas/SM %d smem %6zu  %.4f ms  %5.0f GB/s alg  %5.1f TFLOP/s  maxdiff/max %.2e %s\n", name, fa.numRegs, nb,
           smem, ms, alg / ms / 1e6, flops / ms / 1e9, maxd / maxr, (maxd / maxr < 1e-6) ? "ok" : "MISMATCH");
}

int main(int argc, char **argv) {
    FILE *f = fopen(argc > 1 ? argv[1] : "/tmp/nbr.bin", "rb");
    if (!f) { printf("no map dump\n"); return 1; }
    int64_t n, pairs;
    if (fread(&n, 8, 1, f) != 1 || fread(&pairs, 8, 1, f) != 1) return 1;
    std::vector<int32_t> h(27 * n);
    if (fread(h.data(), 4, 27 * n, f) != (size_t)(27 * n)) return 1;
    fclose(f);
    printf("rows %lld pairs %lld\n", (long long)n, (long long)pairs);
    g_n = n; g_pairs = pairs;
    int32_t *nbr; CK(cudaMalloc(&nbr, 27 * n * 4)); CK(cudaMemcpy(nbr, h.data(), 27 * n * 4, cudaMemcpyHostToDevice));
    g_nbr = nbr;
    CK(cudaMalloc(&g_flush, 256u << 20));
    constexpr int MAXC = 64;
    CK(cudaMalloc(&g_in, n * MAXC * 4)); CK(cudaMalloc(&g_out, n * MAXC * 4)); CK(cudaMalloc(&g_ref, n * MAXC * 4));
    CK(cudaMalloc(&g_res, n * MAXC * 4));
    CK(cudaMalloc(&g_w, 27 * MAXC * MAXC * 4)); CK(cudaMalloc(&g_packed, 27 * MAXC * MAXC * 8 + 1024)); CK(cudaMalloc(&g_bias, MAXC * 4));
    std::vector<float> hin(n * MAXC);
    for (auto &v : hin) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(g_in, hin.data(), n * MAXC * 4, cudaMemcpyHostToDevice));
    for (auto &v : hin) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(g_res, hin.data(), n * MAXC * 4, cudaMemcpyHostToDevice));
    std::vector<float> hw(27 * MAXC * MAXC);
    for (auto &v : hw) v = ((float)rand() / RAND_MAX - 0.5f) * 0.1f;
    CK(cudaMemcpy(g_w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> hb(MAXC);
    for (auto &v : hb) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(g_bias, hb.data(), MAXC * 4, cudaMemcpyHostToDevice));
#define RUNP(CIN, COUT, NT, RG, D, MINB, OPT) run<CIN, COUT, NT, RG, D, MINB, OPT>(#CIN "x" #COUT " " #NT " RG" #RG " D" #D " minb" #MINB " opt" #OPT);
#include "/tmp/variants_pipe.h"
    return 0;
}
