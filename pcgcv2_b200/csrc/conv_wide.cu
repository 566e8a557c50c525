// conv_wide.cu -- C-ABI entry points of the tcgen05 / TMA k=3 convolution for the wide layers (conv_wide.cuh).
#include "conv_wide.cuh"

namespace pcgc {

template <int CIN, int COUT>
static int launch_wide(const uint32_t *in, int in_ld, const int32_t *nbr, int64_t n, const void *packed, float inv_scale,
                       const float *bias, const float *res, int res_ld, float *out, int out_ld, uint32_t *out_h2, int out_h2_ld,
                       int flags, int *overflow, cudaStream_t s) {
    using C = wide::WCfg<CIN, COUT>;
    static_assert(C::SMEM <= 227 * 1024, "wide conv: shared memory budget");
    auto kern = wide::conv_k3_wide_kernel<CIN, COUT>;
    static int ready = 0;
    if (!ready) {                                       // idempotent: two frame workers may both get here
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { set_error("wide conv %dx%d: %s", CIN, COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        ready = 1;
    }
    const int64_t tiles = (n + C::TM - 1) / C::TM;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);           // persistent: one CTA per SM
    kern<<<grid, C::THREADS, C::SMEM, s>>>(in, in_ld, nbr, n, (const unsigned char *)packed, inv_scale, bias, res, res_ld, out, out_ld,
                                          out_h2, out_h2_ld, flags, overflow);
    return check_launch("conv_k3_wide");
}

template <int CIN, int COUT>
static int pack_wide(const float *w, float scale, void *packed, cudaStream_t s) {
    using C = wide::WCfg<CIN, COUT>;
    wide::pack_weights_wide_kernel<CIN, COUT><<<grid_for(27 * C::B_BYTES / 2, 256, 4), 256, 0, s>>>(w, scale, (__half *)packed);
    return check_launch("pack_weights_wide");
}

#define PCGC_WIDE_SHAPES(X) X(64, 64) X(64, 16) X(64, 1) X(32, 32) X(32, 8) X(32, 1) X(16, 32) X(16, 16) X(16, 4) X(16, 1) X(8, 16) X(8, 8)

}  // namespace pcgc

using namespace pcgc;

extern "C" {

size_t pcgc_conv_k3_wide_packed_bytes(int32_t cin, int32_t cout) {
#define X(CI, CO) if (cin == CI && cout == CO) return wide::WCfg<CI, CO>::packed_bytes();
    PCGC_WIDE_SHAPES(X)
#undef X
    return 0;
}

int pcgc_conv_k3_wide_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, void *packed, void *stream) {
#define X(CI, CO) if (cin == CI && cout == CO) return pack_wide<CI, CO>(weight, scale, packed, (cudaStream_t)stream);
    PCGC_WIDE_SHAPES(X)
#undef X
    set_error("pcgc_conv_k3_wide_pack_weights: no tcgen05 kernel for %d -> %d", cin, cout);
    return PCGC_ERR_INVALID;
}

int pcgc_conv_k3_wide_fwd(const uint32_t *feats_h2, int32_t in_ld, const int32_t *nbr, int64_t n, const void *packed, float inv_scale,
                          const float *bias, int32_t cin, int32_t cout, const float *residual, int32_t res_ld, float *out,
                          int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF, "pcgc_conv_k3_wide_fwd: bad n");
    PCGC_REQUIRE(out || out_h2, "pcgc_conv_k3_wide_fwd: no output requested");
    PCGC_REQUIRE((in_ld & 3) == 0 && ((uintptr_t)feats_h2 & 15) == 0, "pcgc_conv_k3_wide_fwd: h2 rows must be 16-byte aligned");
    PCGC_REQUIRE(((uintptr_t)packed & 15) == 0, "pcgc_conv_k3_wide_fwd: packed weights must be 16-byte aligned");
    PCGC_REQUIRE(!out_h2 || cout % 4 == 0, "pcgc_conv_k3_wide_fwd: h2 output needs cout %% 4 == 0");
    if (n == 0) return PCGC_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define X(CI, CO) if (cin == CI && cout == CO) return launch_wide<CI, CO>(feats_h2, in_ld, nbr, n, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    PCGC_WIDE_SHAPES(X)
#undef X
    set_error("pcgc_conv_k3_wide_fwd: no tcgen05 kernel for %d -> %d", cin, cout);
    return PCGC_ERR_INVALID;
}

}  // extern "C"
