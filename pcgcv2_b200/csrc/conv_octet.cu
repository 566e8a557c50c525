// conv_octet.cu -- C-ABI entry points of the full-octet k=3 convolution (conv_octet.cuh): the synthesis
// network's convolutions on 8-child-expanded sets, addressed by the parent set's kernel map.
#include <cstdlib>

#include "conv_octet.cuh"

namespace pcgc {

// per-shape tuning: RG 16-row groups per warp, WARPS per CTA (bounded by 227 KB of shared memory per CTA).
// Variant 0 is the default; PCGC_OCTET_VARIANT=1|2 selects the alternatives (tools/profile_octet.py sweeps).
template <int CIN, int COUT, int V>
struct OctetTune {
    static constexpr int CT = (COUT + 7) / 8;
    // measured on B200 (tools/profile_octet.py, 1.69 M-row decoder level): 16x16 0.463 ms at RG 2 / 8 warps vs
    // 0.412 ms at RG 1 / 16 warps (warps hide the fill, the extra weight reads do not hurt); 16x4 0.283 vs 0.301 ms
    static constexpr bool kWide = CIN == 16 && CT >= 2;
    static constexpr int RG = V == 0 ? (kWide ? 1 : 2) : (V == 1 ? (kWide ? 2 : 1) : 1);
    static constexpr int WARPS = CIN == 8 ? (V == 2 ? 24 : 16)
                                          : (kWide ? (V == 1 ? 8 : 16) : (V == 0 ? 12 : (V == 1 ? 16 : 24)));
    static constexpr int MINB = 1;
};

static int octet_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCGC_OCTET_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 2) v = 0;
    }
    return v;
}

template <int CIN, int COUT, int V>
static int launch_octet_mma_v(const float *in, int in_ld, const int32_t *pnbr, int64_t n_par, const float *packed,
                              const float *bias, const float *res, int res_ld, float *out, int out_ld, int flags,
                              cudaStream_t s) {
    using T = OctetTune<CIN, COUT, V>;
    using C = OctetMmaCfg<CIN, COUT, T::RG, T::WARPS>;
    static_assert(C::smem_bytes() <= 227 * 1024, "octet mma kernel: shared memory budget");
    auto kern = conv_k3_octet_mma_kernel<CIN, COUT, T::RG, T::WARPS, T::MINB>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("octet conv %dx%d: %s", CIN, COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n_par, C::OCTETS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, pnbr, n_par, packed, bias, res,
                                                                                     res_ld, out, out_ld, flags, 0xFFFFE000u);
    return check_launch("conv_k3_octet_mma");
}

template <int CIN, int COUT>
static int launch_octet_mma(const float *in, int in_ld, const int32_t *pnbr, int64_t n_par, const float *packed,
                            const float *bias, const float *res, int res_ld, float *out, int out_ld, int flags,
                            cudaStream_t s) {
    switch (octet_variant()) {
        case 1: return launch_octet_mma_v<CIN, COUT, 1>(in, in_ld, pnbr, n_par, packed, bias, res, res_ld, out, out_ld, flags, s);
        case 2: return launch_octet_mma_v<CIN, COUT, 2>(in, in_ld, pnbr, n_par, packed, bias, res, res_ld, out, out_ld, flags, s);
        default: return launch_octet_mma_v<CIN, COUT, 0>(in, in_ld, pnbr, n_par, packed, bias, res, res_ld, out, out_ld, flags, s);
    }
}

template <int COUT>
static int launch_octet_ffma(const float *in, int in_ld, const int32_t *pnbr, int64_t n_par, const float *weight,
                             const float *bias, const float *res, int res_ld, float *out, int out_ld, int flags,
                             cudaStream_t s) {
    constexpr int RPL = 2, WARPS = 8, MINB = 2;
    using C = OctetFfmaCfg<COUT, RPL, WARPS>;
    auto kern = conv_k3_octet_ffma_kernel<COUT, RPL, WARPS, MINB>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("octet conv 4x%d: %s", COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n_par, C::OCTETS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, pnbr, n_par, weight, bias, res,
                                                                                     res_ld, out, out_ld, flags);
    return check_launch("conv_k3_octet_ffma");
}

static bool octet_mma_shape(int cin, int cout) {
    return (cin == 8 || cin == 16) && (cout == 1 || cout == 4 || cout == 8 || cout == 16);
}
static bool octet_ffma_shape(int cin, int cout) { return cin == 4 && (cout == 4 || cout == 8); }

}  // namespace pcgc

using namespace pcgc;

extern "C" {

size_t pcgc_conv_k3_octet_packed_floats(int32_t cin, int32_t cout) {
    if (octet_mma_shape(cin, cout)) return (size_t)27 * (cin / 8) * ((cout + 7) / 8) * 128;
    if (octet_ffma_shape(cin, cout)) return (size_t)27 * cin * cout;
    return 0;
}

int pcgc_conv_k3_octet_pack_weights(const float *weight, int32_t cin, int32_t cout, float *packed, void *stream) {
    const size_t total = pcgc_conv_k3_octet_packed_floats(cin, cout);
    PCGC_REQUIRE(total > 0 && weight && packed, "pcgc_conv_k3_octet_pack_weights: no full-octet kernel for %dx%d", cin, cout);
    if (octet_ffma_shape(cin, cout)) {
        PCGC_CUDA(cudaMemcpyAsync(packed, weight, total * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return PCGC_OK;
    }
    pack_weights_nt_kernel<<<grid_for((int64_t)total / 2, 256, 4), 256, 0, (cudaStream_t)stream>>>(weight, 27, cin, cout, packed);
    return check_launch("pack_weights_nt");
}

int pcgc_conv_k3_octet_fwd(const float *in, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                           const float *packed, const float *bias, int32_t cin, int32_t cout, const float *residual,
                           int32_t res_ld, float *out, int32_t out_ld, int32_t flags, void *stream) {
    PCGC_REQUIRE(n_parents >= 0 && 8 * n_parents < 0x7FFFFFFF && cin >= 1 && cout >= 1 && in_ld >= cin && out_ld >= cout,
                 "pcgc_conv_k3_octet_fwd: bad shape n_parents=%lld cin=%d cout=%d ld=%d/%d", (long long)n_parents, cin, cout,
                 in_ld, out_ld);
    if (n_parents == 0) return PCGC_OK;
    PCGC_REQUIRE(in && parent_nbr && packed && out, "pcgc_conv_k3_octet_fwd: null pointer");
    PCGC_REQUIRE(pcgc_conv_k3_octet_packed_floats(cin, cout) > 0, "pcgc_conv_k3_octet_fwd: no full-octet kernel for %dx%d", cin, cout);
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_conv_k3_octet_fwd: input rows must be 16-byte aligned (ld %% 4 == 0)");
    cudaStream_t s = (cudaStream_t)stream;
#define MMA(CI, CO) \
    if (cin == CI && cout == CO) return launch_octet_mma<CI, CO>(in, in_ld, parent_nbr, n_parents, packed, bias, residual, res_ld, out, out_ld, flags, s);
    MMA(16, 16) MMA(16, 8) MMA(16, 4) MMA(16, 1) MMA(8, 16) MMA(8, 8) MMA(8, 4) MMA(8, 1)
#undef MMA
    if (cin == 4 && cout == 8) return launch_octet_ffma<8>(in, in_ld, parent_nbr, n_parents, packed, bias, residual, res_ld, out, out_ld, flags, s);
    if (cin == 4 && cout == 4) return launch_octet_ffma<4>(in, in_ld, parent_nbr, n_parents, packed, bias, residual, res_ld, out, out_ld, flags, s);
    set_error("pcgc_conv_k3_octet_fwd: shape %dx%d has no instantiation", cin, cout);
    return PCGC_ERR_INVALID;
}

}  // extern "C"
