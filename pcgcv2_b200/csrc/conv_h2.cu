// conv_h2.cu -- C-ABI entry points of the pre-split half-precision k=3 convolution (conv_h2.cuh) and of the
// fp32 <-> h2 feature format conversions.
#include <cstdlib>

#include "conv_h2.cuh"
#include "conv_octet_h2.cuh"

namespace pcgc {

// per-shape tuning, measured on B200 with tools/bench_h2.cu on the 1.69 M-row decoder level
// (profiles/r01_h2_sweep.txt): RG row groups per warp, D gather stages, WARPS per CTA, MINB CTAs per SM
template <int CIN, int COUT>
struct H2Tune {
    static constexpr bool NT = COUT < 16;
    static constexpr int RG = NT ? (CIN == 32 ? 2 : 1) : ((CIN == 16 && COUT == 16) || CIN == 8 ? 4 : 2);
    static constexpr int D = NT ? (CIN <= 16 ? 3 : 2) : (CIN == 64 ? 1 : 2);
    static constexpr int WARPS = (!NT && CIN * COUT >= 1024) ? 16 : 8;
    static constexpr int MINB = NT ? (CIN <= 16 ? 4 : 2) : (WARPS == 16 ? 1 : 2);
};

template <int CIN, int COUT>
static int launch_h2(const uint32_t *in, int in_ld, const int32_t *nbr, int64_t n, const uint32_t *packed, float inv_scale,
                     const float *bias, const float *res, int res_ld, float *out, int out_ld, uint32_t *out_h2, int out_h2_ld,
                     int flags, int *overflow, cudaStream_t s) {
    using T = H2Tune<CIN, COUT>;
    using C = H2Cfg<CIN, COUT, T::NT, T::RG, T::D, T::WARPS>;
    static_assert(C::smem_bytes() <= 227 * 1024, "h2 kernel: shared memory budget");
    auto kern = conv_k3_h2_kernel<CIN, COUT, T::NT, T::RG, T::D, T::WARPS, T::MINB>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("h2 conv %dx%d: %s", CIN, COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n, C::ROWS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, nbr, n, packed, inv_scale, bias, res, res_ld,
                                                                               out, out_ld, out_h2, out_h2_ld, flags, overflow);
    return check_launch("conv_k3_h2");
}

// full-octet variant: RG row groups per warp, WARPS per CTA (one CTA per SM; each warp owns the halos of its octets).
// PCGC_OCTET_H2_VARIANT=1|2 selects the alternatives (tools/profile_octet_h2.py sweeps them).
template <int CIN, int COUT, int V>
struct OctetH2Tune {
    static constexpr bool NT = COUT < 16;
    // measured on B200 (tools/profile_octet_h2.py, 1.69 M-row decoder level, profiles/r01_h2_sweep.txt):
    // 16x4 0.217 ms at RG 2 / 8 warps vs 0.250 at RG 1 / 16 warps (the loop is shared-memory bound: a bigger row group
    // re-uses each weight LDS); 16x16 0.297 vs 0.307; 16x32 0.477 vs 0.424 (registers)
    // V0 (default): one halo buffer per warp, the biggest row group that fits (the loop is shared-memory bound: a bigger
    // group re-uses each weight LDS).  V1 / V2: double-buffered halos (the next tile is staged during the MMAs) with the
    // smaller row groups that leaves room for -- measured SLOWER on B200 (profiles/r01_h2_sweep.txt: 16x16 0.301 ms vs
    // 0.370 / 0.400; 16x4 0.217 vs 0.311 / 0.238), i.e. the halo latency is already hidden by the other warps and the
    // extra weight reads cost more; kept selectable for the next round's CTA-shared halo work.
    // (12 instead of 8 warps for the narrow-output CIN = 16 kernels: 168 registers, spills, 0.262 vs 0.224 ms -- l1tex bound, not latency bound)
    static constexpr bool DB = V != 0;
    static constexpr int RG = DB ? (NT ? 1 : 2) : (NT ? 2 : (COUT == 16 ? 4 : 2));
    static constexpr int WARPS = DB ? (V == 1 ? 10 : 8) : (NT ? (CIN == 8 ? 16 : 8) : (COUT == 16 ? (CIN == 8 ? 12 : 8) : 16));
};

// PCGC_OCTET_TILE_ORDER=0 keeps the strided tile order; default 1 = one contiguous run of tiles per CTA
static int octet_tile_flag() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCGC_OCTET_TILE_ORDER");
        v = (e && atoi(e) == 0) ? 0 : PCGC_TILES_CHUNKED;
    }
    return v;
}

static int octet_h2_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCGC_OCTET_H2_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 2) v = 0;
    }
    return v;
}

template <int CIN, int COUT, int V>
static int launch_octet_h2_v(const uint32_t *in, int in_ld, const int32_t *pnbr, int64_t n_par, const uint32_t *packed,
                             float inv_scale, const float *bias, const float *res, int res_ld, float *out, int out_ld,
                             uint32_t *out_h2, int out_h2_ld, int flags, int *overflow, cudaStream_t s) {
    using T = OctetH2Tune<CIN, COUT, V>;
    using C = OctetH2Cfg<CIN, COUT, T::NT, T::RG, T::WARPS, T::DB>;
    static_assert(C::smem_bytes() <= 227 * 1024, "octet h2 kernel: shared memory budget");
    auto kern = conv_k3_octet_h2_kernel<CIN, COUT, T::NT, T::RG, T::WARPS, 1, T::DB>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("octet h2 conv %dx%d: %s", CIN, COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n_par, C::OCTETS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, pnbr, n_par, packed, inv_scale, bias,
                                                                                     res, res_ld, out, out_ld, out_h2, out_h2_ld,
                                                                                     flags | octet_tile_flag(), overflow);
    return check_launch("conv_k3_octet_h2");
}

template <int CIN, int COUT>
static int launch_octet_h2(const uint32_t *in, int in_ld, const int32_t *pnbr, int64_t n_par, const uint32_t *packed, float inv_scale,
                           const float *bias, const float *res, int res_ld, float *out, int out_ld, uint32_t *out_h2,
                           int out_h2_ld, int flags, int *overflow, cudaStream_t s) {
    switch (octet_h2_variant()) {
        case 1: return launch_octet_h2_v<CIN, COUT, 1>(in, in_ld, pnbr, n_par, packed, inv_scale, bias, res, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
        case 2: return launch_octet_h2_v<CIN, COUT, 2>(in, in_ld, pnbr, n_par, packed, inv_scale, bias, res, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
        default: return launch_octet_h2_v<CIN, COUT, 0>(in, in_ld, pnbr, n_par, packed, inv_scale, bias, res, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    }
}

template <int COUT, bool K1TAIL = false>
static int launch_octet_h2c4(const uint32_t *in, int in_ld, const int32_t *pnbr, int64_t n_par, const uint32_t *packed, float inv_scale,
                             const float *bias, const float *res, int res_ld, float *out, int out_ld, uint32_t *out_h2,
                             int out_h2_ld, int flags, int *overflow, cudaStream_t s, const float *tail_w = nullptr,
                             const float *tail_b = nullptr) {
    constexpr int RG = 2, WARPS = 8, MINB = 2;
    using C = OctetH2C4Cfg<COUT, RG, WARPS>;
    auto kern = conv_k3_octet_h2c4_kernel<COUT, RG, WARPS, MINB, K1TAIL>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("octet h2 conv 4x%d: %s", COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n_par, C::OCTETS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, pnbr, n_par, packed, inv_scale, bias,
                                                                                     res, res_ld, out, out_ld, out_h2, out_h2_ld,
                                                                                     flags | octet_tile_flag(), overflow, tail_w, tail_b);
    return check_launch("conv_k3_octet_h2c4");
}

template <int RG, int WARPS, int MINB>
static int launch_octet_h2c4_dual(const uint32_t *in, int in_ld, const int32_t *pnbr, int64_t n_par, const uint32_t *packed_a, float inv_a,
                                  const float *bias_a, const uint32_t *packed_b, float inv_b, const float *bias_b, const float *tail_w,
                                  const float *tail_b, const float *res, int res_ld, float *out, int out_ld, uint32_t *out_h2,
                                  int out_h2_ld, int *overflow, cudaStream_t s) {
    using C = OctetH2C4DualCfg<RG, WARPS>;
    static_assert(C::smem_bytes() <= 227 * 1024, "dual octet h2c4 kernel: shared memory budget");
    auto kern = conv_k3_octet_h2c4_dual_kernel<RG, WARPS, MINB>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("dual octet h2c4 conv: %s", cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n_par, C::OCTETS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, pnbr, n_par, packed_a, inv_a, bias_a, packed_b,
                                                                                     inv_b, bias_b, tail_w, tail_b, res, res_ld, out, out_ld,
                                                                                     out_h2, out_h2_ld, octet_tile_flag(), overflow);
    return check_launch("conv_k3_octet_h2c4_dual");
}

static bool octet_h2c4_shape(int cin, int cout) { return cin == 4 && (cout == 4 || cout == 8); }
static bool octet_h2_shape(int cin, int cout) {
    return (cin == 16 && (cout == 1 || cout == 4 || cout == 8 || cout == 16 || cout == 32)) || (cin == 8 && (cout == 8 || cout == 16)) ||
           octet_h2c4_shape(cin, cout);       // (cin = 32 was measured at the 435 k-row level: 0.142 / 0.313 ms for 32x8 / 32x32 against
                                              // 0.119 / 0.217 ms of the gather kernels -- 8 KB halos leave 6-10 warps per SM; not instantiated)
}

// k=2 stride-2 convolution = the gather kernel over the 8 child slots of every parent (KV = 8, T formulation)
template <int CIN, int COUT>
static int launch_down_h2(const uint32_t *in, int in_ld, const int32_t *child_map, int64_t n_par, const uint32_t *packed,
                          float inv_scale, const float *bias, float *out, int out_ld, uint32_t *out_h2, int out_h2_ld, int flags,
                          int *overflow, cudaStream_t s) {
    constexpr int RG = 2, D = CIN == 64 ? 1 : 2, WARPS = 8, MINB = CIN == 64 ? 1 : 2;
    using C = H2Cfg<CIN, COUT, false, RG, D, WARPS, 8>;
    auto kern = conv_k3_h2_kernel<CIN, COUT, false, RG, D, WARPS, MINB, 8>;
    static int ctas = 0;
    if (ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) { set_error("h2 k2s2 conv %dx%d: %s", CIN, COUT, cudaGetErrorString(e)); return PCGC_ERR_CUDA; }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::smem_bytes()) != cudaSuccess || nb < 1) nb = 1;
        ctas = nb;
    }
    kern<<<grid_for(n_par, C::ROWS_PER_CTA, ctas), C::THREADS, C::smem_bytes(), s>>>(in, in_ld, child_map, n_par, packed, inv_scale, bias,
                                                                                   nullptr, 0, out, out_ld, out_h2, out_h2_ld, flags,
                                                                                   overflow);
    return check_launch("conv_k2s2_h2");
}

template <int CIN>
static int launch_dense_h2(const uint32_t *in, int in_ld, int64_t n, const uint32_t *packed, int nc, float inv_scale,
                           const float *bias, float *out, int out_ld, uint32_t *out_h2, int out_h2_ld, int flags, int *overflow,
                           cudaStream_t s) {
    constexpr int RG = 2, WARPS = 8, MINB = 2;
    using C = DenseH2Cfg<CIN, RG, WARPS>;
    auto kern = dense_h2_kernel<CIN, RG, WARPS, MINB>;
    const size_t smem = C::smem_bytes(nc);
    // opt in once to the largest weight block any supported shape needs (nc <= 512); a function-local static is
    // initialised exactly once even when two frame workers arrive here together
    static const cudaError_t opt_in = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes(512));
    if (opt_in != cudaSuccess || smem > C::smem_bytes(512)) {
        set_error("dense h2 %dx%d: %s", CIN, nc, opt_in != cudaSuccess ? cudaGetErrorString(opt_in) : "nc > 512");
        return PCGC_ERR_CUDA;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, smem) != cudaSuccess || nb < 1) nb = 1;
    kern<<<grid_for(n, C::ROWS_PER_CTA, nb), C::THREADS, smem, s>>>(in, in_ld, n, packed, nc, inv_scale, bias, out, out_ld, out_h2,
                                                                   out_h2_ld, flags, overflow);
    return check_launch("dense_h2");
}

static bool down_h2_shape(int cin, int cout) { return (cin == 16 && cout == 32) || (cin == 32 && cout == 64) || (cin == 64 && cout == 32); }
static bool up_h2_shape(int cin, int cout) { return (cin == 16 || cin == 32 || cin == 64) && cout % 2 == 0 && 8 * cout <= 512 && (8 * cout) % 16 == 0; }

static bool h2_shape(int cin, int cout) {
    if (cin == 8) return cout == 8 || cout == 16;
    if (cin == 16) return cout == 1 || cout == 4 || cout == 8 || cout == 16 || cout == 32;
    if (cin == 32) return cout == 1 || cout == 4 || cout == 8 || cout == 32;
    if (cin == 64) return cout == 1 || cout == 8 || cout == 16;
    return false;
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

int pcgc_split_h2(const float *in, int32_t in_ld, int64_t n, int32_t c, uint32_t *out_h2, int32_t out_ld, int32_t *overflow,
                  void *stream) {
    PCGC_REQUIRE(n >= 0 && c >= 4 && c % 4 == 0 && in_ld >= c && out_ld >= c, "pcgc_split_h2: bad shape n=%lld c=%d ld=%d/%d",
                 (long long)n, c, in_ld, out_ld);
    if (n == 0) return PCGC_OK;
    PCGC_REQUIRE(in && out_h2, "pcgc_split_h2: null pointer");
    const int vec_ok = (in_ld % 4 == 0) && (out_ld % 4 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)out_h2 & 15) == 0);
    split_h2_kernel<<<grid_for(n * (c / 4), 256, 8), 256, 0, (cudaStream_t)stream>>>(in, in_ld, n, c, out_h2, out_ld, overflow, vec_ok);
    return check_launch("split_h2");
}

int pcgc_join_h2(const uint32_t *in_h2, int32_t in_ld, int64_t n, int32_t c, float *out, int32_t out_ld, void *stream) {
    PCGC_REQUIRE(n >= 0 && c >= 4 && c % 4 == 0 && in_ld >= c && out_ld >= c, "pcgc_join_h2: bad shape n=%lld c=%d ld=%d/%d",
                 (long long)n, c, in_ld, out_ld);
    if (n == 0) return PCGC_OK;
    PCGC_REQUIRE(in_h2 && out, "pcgc_join_h2: null pointer");
    join_h2_kernel<<<grid_for(n * (c / 4), 256, 8), 256, 0, (cudaStream_t)stream>>>(in_h2, in_ld, n, c, out, out_ld);
    return check_launch("join_h2");
}

size_t pcgc_conv_k3_h2_packed_words(int32_t cin, int32_t cout) {
    if (octet_h2c4_shape(cin, cout)) return (size_t)27 * ((cout + 7) / 8) * 64;      // full-octet kernel only
    if (!h2_shape(cin, cout)) return 0;
    const int ks = cin == 8 ? 1 : cin / 16;                                          // cin 8: [hi | lo] fill one k-step
    return cout < 16 ? (size_t)27 * ks * ((cout + 7) / 8) * 128 : (size_t)27 * ks * (cout / 16) * 256;
}

int pcgc_conv_k3_h2_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, uint32_t *packed, void *stream) {
    const size_t total = pcgc_conv_k3_h2_packed_words(cin, cout);
    PCGC_REQUIRE(total > 0 && weight && packed, "pcgc_conv_k3_h2_pack_weights: no h2 kernel for %dx%d", cin, cout);
    PCGC_REQUIRE(scale > 0.f, "pcgc_conv_k3_h2_pack_weights: scale must be positive");
    if (octet_h2c4_shape(cin, cout)) {
        pack_weights_h2c4_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(weight, cout, scale, packed);
        return check_launch("pack_weights_h2c4");
    }
    if (cin == 8) {
        pack_weights_h2c8_kernel<<<grid_for((int64_t)total / 2, 256, 4), 256, 0, (cudaStream_t)stream>>>(weight, 27, cout, cout < 16 ? 1 : 0, scale,
                                                                                                      packed);
        return check_launch("pack_weights_h2c8");
    }
    pack_weights_h2_kernel<<<grid_for((int64_t)total / 2, 256, 4), 256, 0, (cudaStream_t)stream>>>(weight, 27, cin, cout, cout < 16 ? 1 : 0,
                                                                                                scale, packed);
    return check_launch("pack_weights_h2");
}

int pcgc_conv_k3_h2_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *nbr, int64_t n, const uint32_t *packed,
                        float inv_scale, const float *bias, int32_t cin, int32_t cout, const float *residual, int32_t res_ld,
                        float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags, int32_t *overflow,
                        void *stream) {
    PCGC_REQUIRE(n >= 0 && cin >= 1 && cout >= 1 && in_ld >= cin, "pcgc_conv_k3_h2_fwd: bad shape n=%lld cin=%d cout=%d ld=%d",
                 (long long)n, cin, cout, in_ld);
    if (n == 0) return PCGC_OK;
    PCGC_REQUIRE(in_h2 && nbr && packed && (out || out_h2), "pcgc_conv_k3_h2_fwd: null pointer");
    PCGC_REQUIRE(h2_shape(cin, cout), "pcgc_conv_k3_h2_fwd: no gather h2 kernel for %dx%d", cin, cout);
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in_h2 & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_conv_k3_h2_fwd: input rows must be 16-byte aligned (ld %% 4 == 0)");
    if (cout % 2 == 0) {
        PCGC_REQUIRE(!out || (out_ld >= cout && out_ld % 2 == 0 && ((uintptr_t)out & 7) == 0), "pcgc_conv_k3_h2_fwd: out must be 8-byte aligned");
        PCGC_REQUIRE(!residual || (res_ld % 2 == 0 && ((uintptr_t)residual & 7) == 0), "pcgc_conv_k3_h2_fwd: residual must be 8-byte aligned");
        PCGC_REQUIRE(!out_h2 || (cout % 4 == 0 && out_h2_ld >= cout && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                     "pcgc_conv_k3_h2_fwd: h2 output needs cout %% 4 == 0 and 16-byte aligned rows");
    } else {
        PCGC_REQUIRE(out && !out_h2 && out_ld >= cout, "pcgc_conv_k3_h2_fwd: odd cout writes fp32 only");
    }
    cudaStream_t s = (cudaStream_t)stream;
#define H2(CI, CO) \
    if (cin == CI && cout == CO) return launch_h2<CI, CO>(in_h2, in_ld, nbr, n, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    H2(8, 8) H2(8, 16) H2(16, 1) H2(16, 4) H2(16, 8) H2(16, 16) H2(16, 32) H2(32, 1) H2(32, 4) H2(32, 8) H2(32, 32) H2(64, 1) H2(64, 8) H2(64, 16)
#undef H2
    set_error("pcgc_conv_k3_h2_fwd: shape %dx%d has no instantiation", cin, cout);
    return PCGC_ERR_INVALID;
}

// ---- k=2 stride-2 (a5) and its generative transpose (a6) on the tensor cores over h2 features
size_t pcgc_conv_k2s2_h2_packed_words(int32_t cin, int32_t cout) {
    return down_h2_shape(cin, cout) ? (size_t)8 * (cin / 16) * (cout / 16) * 256 : 0;
}

int pcgc_conv_k2s2_h2_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, uint32_t *packed, void *stream) {
    const size_t total = pcgc_conv_k2s2_h2_packed_words(cin, cout);
    PCGC_REQUIRE(total > 0 && weight && packed && scale > 0.f, "pcgc_conv_k2s2_h2_pack_weights: no h2 kernel for %dx%d", cin, cout);
    pack_weights_h2_kernel<<<grid_for((int64_t)total / 2, 256, 4), 256, 0, (cudaStream_t)stream>>>(weight, 8, cin, cout, 0, scale, packed);
    return check_launch("pack_weights_h2");
}

int pcgc_conv_k2s2_h2_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *child_map, int64_t n_parents,
                          const uint32_t *packed, float inv_scale, const float *bias, int32_t cin, int32_t cout, float *out,
                          int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream) {
    PCGC_REQUIRE(n_parents >= 0 && down_h2_shape(cin, cout) && in_ld >= cin, "pcgc_conv_k2s2_h2_fwd: bad shape n=%lld %dx%d",
                 (long long)n_parents, cin, cout);
    if (n_parents == 0) return PCGC_OK;
    PCGC_REQUIRE(in_h2 && child_map && packed && (out || out_h2), "pcgc_conv_k2s2_h2_fwd: null pointer");
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in_h2 & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_conv_k2s2_h2_fwd: input rows must be 16-byte aligned");
    PCGC_REQUIRE(!out || (out_ld >= cout && out_ld % 2 == 0 && ((uintptr_t)out & 7) == 0), "pcgc_conv_k2s2_h2_fwd: out must be 8-byte aligned");
    PCGC_REQUIRE(!out_h2 || (out_h2_ld >= cout && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                 "pcgc_conv_k2s2_h2_fwd: h2 output rows must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (cin == 16) return launch_down_h2<16, 32>(in_h2, in_ld, child_map, n_parents, packed, inv_scale, bias, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    if (cin == 32) return launch_down_h2<32, 64>(in_h2, in_ld, child_map, n_parents, packed, inv_scale, bias, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    return launch_down_h2<64, 32>(in_h2, in_ld, child_map, n_parents, packed, inv_scale, bias, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
}

size_t pcgc_convT_k2s2_h2_packed_words(int32_t cin, int32_t cout) {
    return up_h2_shape(cin, cout) ? (size_t)(cin / 16) * (8 * cout / 16) * 256 : 0;
}

// weight [8][cin][cout] -> the dense [cin][8 * cout] operand (column k * cout + co), split and packed
static __global__ void gather_up_weights_kernel(const float *__restrict__ w, int cin, int cout, float *__restrict__ dense) {
    const int total = 8 * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout, ci = (i / cout) % cin, k = i / (cout * cin);
        dense[(size_t)ci * 8 * cout + k * cout + co] = w[i];
    }
}

int pcgc_convT_k2s2_h2_pack_weights(const float *weight, int32_t cin, int32_t cout, float scale, float *dense_ws, uint32_t *packed,
                                    void *stream) {
    const size_t total = pcgc_convT_k2s2_h2_packed_words(cin, cout);
    PCGC_REQUIRE(total > 0 && weight && packed && dense_ws && scale > 0.f, "pcgc_convT_k2s2_h2_pack_weights: no h2 kernel for %dx%d", cin, cout);
    gather_up_weights_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(weight, cin, cout, dense_ws);
    pack_weights_h2_kernel<<<grid_for((int64_t)total / 2, 256, 4), 256, 0, (cudaStream_t)stream>>>(dense_ws, 1, cin, 8 * cout, 0, scale, packed);
    return check_launch("pack_weights_h2");
}

int pcgc_convT_k2s2_h2_fwd(const uint32_t *in_h2, int32_t in_ld, int64_t n_in, const uint32_t *packed, float inv_scale,
                           const float *bias8, int32_t cin, int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2,
                           int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream) {
    PCGC_REQUIRE(n_in >= 0 && up_h2_shape(cin, cout) && in_ld >= cin, "pcgc_convT_k2s2_h2_fwd: bad shape n=%lld %dx%d", (long long)n_in, cin,
                 cout);
    if (n_in == 0) return PCGC_OK;
    PCGC_REQUIRE(in_h2 && packed && (out || out_h2), "pcgc_convT_k2s2_h2_fwd: null pointer");
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in_h2 & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_convT_k2s2_h2_fwd: input rows must be 16-byte aligned");
    // the 8 child rows of one input row must be one contiguous row of 8 * cout values
    PCGC_REQUIRE(!out || (out_ld == cout && ((uintptr_t)out & 7) == 0), "pcgc_convT_k2s2_h2_fwd: out must be contiguous [8n, cout]");
    PCGC_REQUIRE(!out_h2 || (out_h2_ld == cout && cout % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                 "pcgc_convT_k2s2_h2_fwd: out_h2 must be contiguous [8n, cout]");
    cudaStream_t s = (cudaStream_t)stream;
    const int nc = 8 * cout;
    if (cin == 16) return launch_dense_h2<16>(in_h2, in_ld, n_in, packed, nc, inv_scale, bias8, out, nc, out_h2, nc, flags, overflow, s);
    if (cin == 32) return launch_dense_h2<32>(in_h2, in_ld, n_in, packed, nc, inv_scale, bias8, out, nc, out_h2, nc, flags, overflow, s);
    return launch_dense_h2<64>(in_h2, in_ld, n_in, packed, nc, inv_scale, bias8, out, nc, out_h2, nc, flags, overflow, s);
}

int pcgc_conv_k3_octet_h2_supported(int32_t cin, int32_t cout) { return octet_h2_shape(cin, cout) ? 1 : 0; }

int pcgc_irn16_second_stage_fwd(const uint32_t *ab_h2, int32_t ab_ld, const int32_t *parent_nbr, int64_t n_parents,
                                const uint32_t *packed01, float inv_scale01, const float *bias01, const uint32_t *packed11,
                                float inv_scale11, const float *bias11, const float *weight12, const float *bias12, const float *x,
                                int32_t x_ld, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t *overflow,
                                void *stream) {
    PCGC_REQUIRE(n_parents >= 0 && 8 * n_parents < 0x7FFFFFFF && ab_ld >= 8, "pcgc_irn16_second_stage_fwd: bad shape");
    if (n_parents == 0) return PCGC_OK;
    PCGC_REQUIRE(ab_h2 && parent_nbr && packed01 && packed11 && weight12 && x && (out || out_h2), "pcgc_irn16_second_stage_fwd: null pointer");
    PCGC_REQUIRE((ab_ld % 4 == 0) && (((uintptr_t)ab_h2 & 15) == 0), "pcgc_irn16_second_stage_fwd: input rows must be 16-byte aligned");
    PCGC_REQUIRE(x_ld >= 16 && x_ld % 2 == 0 && ((uintptr_t)x & 7) == 0, "pcgc_irn16_second_stage_fwd: x must be 8-byte aligned [n][16]");
    PCGC_REQUIRE(!out || (out_ld >= 16 && out_ld % 2 == 0 && ((uintptr_t)out & 7) == 0), "pcgc_irn16_second_stage_fwd: out must be 8-byte aligned");
    PCGC_REQUIRE(!out_h2 || (out_h2_ld >= 16 && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                 "pcgc_irn16_second_stage_fwd: h2 output rows must be 16-byte aligned");
    return launch_octet_h2c4_dual<2, 8, 2>(ab_h2, ab_ld, parent_nbr, n_parents, packed01, inv_scale01, bias01, packed11, inv_scale11, bias11,
                                           weight12, bias12, x, x_ld, out, out_ld, out_h2, out_h2_ld, overflow, (cudaStream_t)stream);
}

int pcgc_conv_k3_octet_h2_k1_supported(int32_t cin, int32_t cmid, int32_t cout) { return cin == 4 && cmid == 4 && cout == 8 ? 1 : 0; }

int pcgc_conv_k3_octet_h2_k1_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                                 const uint32_t *packed, float inv_scale, const float *bias, int32_t cin, int32_t cmid,
                                 const float *tail_weight, const float *tail_bias, int32_t cout, const float *residual, int32_t res_ld,
                                 float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t *overflow, void *stream) {
    PCGC_REQUIRE(n_parents >= 0 && 8 * n_parents < 0x7FFFFFFF && in_ld >= cin, "pcgc_conv_k3_octet_h2_k1_fwd: bad shape");
    PCGC_REQUIRE(pcgc_conv_k3_octet_h2_k1_supported(cin, cmid, cout), "pcgc_conv_k3_octet_h2_k1_fwd: no kernel for %d -> %d -> %d", cin, cmid, cout);
    if (n_parents == 0) return PCGC_OK;
    PCGC_REQUIRE(in_h2 && parent_nbr && packed && tail_weight && (out || out_h2), "pcgc_conv_k3_octet_h2_k1_fwd: null pointer");
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in_h2 & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_conv_k3_octet_h2_k1_fwd: input rows must be 16-byte aligned (ld %% 4 == 0)");
    PCGC_REQUIRE(!out || (out_ld >= cout && out_ld % 2 == 0 && ((uintptr_t)out & 7) == 0), "pcgc_conv_k3_octet_h2_k1_fwd: out must be 8-byte aligned");
    PCGC_REQUIRE(!residual || (res_ld % 2 == 0 && ((uintptr_t)residual & 7) == 0), "pcgc_conv_k3_octet_h2_k1_fwd: residual must be 8-byte aligned");
    PCGC_REQUIRE(!out_h2 || (out_h2_ld >= cout && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                 "pcgc_conv_k3_octet_h2_k1_fwd: h2 output needs 16-byte aligned rows");
    return launch_octet_h2c4<4, true>(in_h2, in_ld, parent_nbr, n_parents, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2,
                                      out_h2_ld, 0, overflow, (cudaStream_t)stream, tail_weight, tail_bias);
}

int pcgc_conv_k3_octet_h2_fwd(const uint32_t *in_h2, int32_t in_ld, const int32_t *parent_nbr, int64_t n_parents,
                              const uint32_t *packed, float inv_scale, const float *bias, int32_t cin, int32_t cout,
                              const float *residual, int32_t res_ld, float *out, int32_t out_ld, uint32_t *out_h2,
                              int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream) {
    PCGC_REQUIRE(n_parents >= 0 && 8 * n_parents < 0x7FFFFFFF && cin >= 1 && cout >= 1 && in_ld >= cin,
                 "pcgc_conv_k3_octet_h2_fwd: bad shape n_parents=%lld cin=%d cout=%d ld=%d", (long long)n_parents, cin, cout, in_ld);
    if (n_parents == 0) return PCGC_OK;
    PCGC_REQUIRE(in_h2 && parent_nbr && packed && (out || out_h2), "pcgc_conv_k3_octet_h2_fwd: null pointer");
    PCGC_REQUIRE(octet_h2_shape(cin, cout), "pcgc_conv_k3_octet_h2_fwd: no full-octet h2 kernel for %dx%d", cin, cout);
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in_h2 & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_conv_k3_octet_h2_fwd: input rows must be 16-byte aligned (ld %% 4 == 0)");
    if (cout % 2 == 0) {
        PCGC_REQUIRE(!out || (out_ld >= cout && out_ld % 2 == 0 && ((uintptr_t)out & 7) == 0), "pcgc_conv_k3_octet_h2_fwd: out must be 8-byte aligned");
        PCGC_REQUIRE(!residual || (res_ld % 2 == 0 && ((uintptr_t)residual & 7) == 0), "pcgc_conv_k3_octet_h2_fwd: residual must be 8-byte aligned");
        PCGC_REQUIRE(!out_h2 || (cout % 4 == 0 && out_h2_ld >= cout && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                     "pcgc_conv_k3_octet_h2_fwd: h2 output needs cout %% 4 == 0 and 16-byte aligned rows");
    } else {
        PCGC_REQUIRE(out && !out_h2 && out_ld >= cout, "pcgc_conv_k3_octet_h2_fwd: odd cout writes fp32 only");
    }
    cudaStream_t s = (cudaStream_t)stream;
#define OH2(CI, CO) \
    if (cin == CI && cout == CO) return launch_octet_h2<CI, CO>(in_h2, in_ld, parent_nbr, n_parents, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    OH2(8, 8) OH2(8, 16) OH2(16, 1) OH2(16, 4) OH2(16, 8) OH2(16, 16) OH2(16, 32)
#undef OH2
    if (cin == 4 && cout == 8) return launch_octet_h2c4<8>(in_h2, in_ld, parent_nbr, n_parents, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    if (cin == 4 && cout == 4) return launch_octet_h2c4<4>(in_h2, in_ld, parent_nbr, n_parents, packed, inv_scale, bias, residual, res_ld, out, out_ld, out_h2, out_h2_ld, flags, overflow, s);
    set_error("pcgc_conv_k3_octet_h2_fwd: shape %dx%d has no instantiation", cin, cout);
    return PCGC_ERR_INVALID;
}

}  // extern "C"
