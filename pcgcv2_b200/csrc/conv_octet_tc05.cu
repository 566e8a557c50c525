// conv_octet_tc05.cu -- C-ABI entry points of the full-octet tcgen05 convolution (conv_octet_tc05.cuh).
#include "conv_octet_tc05.cuh"

namespace pcgc {

template <int CIN, int COUT>
static int launch_otc(const uint32_t *in, int in_ld, const int32_t *pnbr, int64_t n_par, const void *packed, float inv_scale,
                      const float *bias, const float *res, int res_ld, float *out, int out_ld, uint32_t *out_h2, int out_h2_ld,
                      int flags, int *overflow, cudaStream_t s) {
    using C = otc::OCfg<CIN, COUT>;
    auto kern = otc::conv_k3_octet_tc05_kernel<CIN, COUT>;
    static int ready = 0;
    if (!ready) {                                       // idempotent: two frame workers may both get here
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { set_error("octet tcgen05 conv %dx%d: %s", CIN, COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        ready = 1;
    }
    const int64_t tiles = (n_par + C::TO - 1) / C::TO;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);           // persistent: one CTA per SM
    kern<<<grid, C::THREADS, C::SMEM, s>>>(in, in_ld, pnbr, n_par, (const unsigned char *)packed, inv_scale, bias, res, res_ld, out,
                                          out_ld, out_h2, out_h2_ld, flags, overflow);
    return check_launch("conv_k3_octet_tc05");
}

template <int CIN, int COUT>
static int pack_otc(const float *w, float scale, void *packed, cudaStream_t s) {
    using C = otc::OCfg<CIN, COUT>;
    otc::pack_weights_otc_kernel<CIN, COUT><<<grid_for(C::W_BYTES / 2, 256, 4), 256, 0, s>>>(w, scale, (__half *)packed);
    return check_launch("pack_weights_otc");
}

#define PCGC_OTC_SHAPES(X) X(16, 16) X(16, 8) X(16, 4) X(16, 1)

}  // namespace pcgc

using namespace pcgc;

extern "C" {

size_t pcgc_conv_k3_octet_tc05_packed_bytes(int32_t cin, int32_t cout) {
#define X(CI, CO) if (cin == CI && cout == CO) return otc::OCfg<CI, CO>::packed_bytes();
    PCGC_OTC_SHAPES(X)
#undef X
    return 0;
}

int pcgc_conv_k3_octet_tc05_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, void *packed, void *stream) {
#define X(CI, CO) if (cin == CI && cout == CO) return pack_otc<CI, CO>(weight, scale, packed, (cudaStream_t)stream);
    PCGC_OTC_SHAPES(X)
#undef X
    set_error("pcgc_conv_k3_octet_tc05_pack_weights: no kernel for %d -> %d", cin, cout);
    return PCGC_ERR_INVALID;
}

int pcgc_conv_k3_octet_tc05_fwd(const uint32_t *feats_h2, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents, const void *packed,
                                float inv_scale, const float *bias, int32_t cin, int32_t cout, const float *residual, int32_t res_ld,
                                float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags, int32_t *overflow,
                                void *stream) {
    PCGC_REQUIRE(n_parents >= 0 && n_parents < (1 << 28), "pcgc_conv_k3_octet_tc05_fwd: bad n_parents");
    PCGC_REQUIRE(out || out_h2, "pcgc_conv_k3_octet_tc05_fwd: no output requested");
    PCGC_REQUIRE((in_ld & 3) == 0 && ((uintptr_t)feats_h2 & 15) == 0, "pcgc_conv_k3_octet_tc05_fwd: h2 rows must be 16-byte aligned");
    PCGC_REQUIRE(((uintptr_t)packed & 15) == 0, "pcgc_conv_k3_octet_tc05_fwd: packed weights must be 16-byte aligned");
    PCGC_REQUIRE(!out_h2 || cout % 4 == 0, "pcgc_conv_k3_octet_tc05_fwd: h2 output needs cout %% 4 == 0");
    if (n_parents == 0) return PCGC_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define X(CI, CO) if (cin == CI && cout == CO) return launch_otc<CI, CO>(feats_h2, in_ld, parent_nbr, n_parents, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    PCGC_OTC_SHAPES(X)
#undef X
    set_error("pcgc_conv_k3_octet_tc05_fwd: no kernel for %d -> %d", cin, cout);
    return PCGC_ERR_INVALID;
}

}  // extern "C"
