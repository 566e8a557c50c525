// conv_inst.cu -- template instantiations of the specialised convolution kernels for ONE input
// channel count (compiled once per -DPCGC_CI=<cin>, in parallel; see Makefile).
#ifndef PCGC_CI
#error "compile with -DPCGC_CI=<input channels>"
#endif
#include "conv_rowlane.cuh"
#include "conv_tile.cuh"
#include "conv_mma.cuh"
#include "conv_pipe.cuh"
#include "conv_dispatch.h"

namespace pcgc {

constexpr size_t kRowLaneSmemLimit = 160 * 1024;   // weights resident in smem up to here, else tile kernel

static inline int aligned_bits(const float *in, int in_ld, const float *out, int out_ld, const float *res, int res_ld) {
    const bool a_in = (((uintptr_t)in & 15) == 0) && (in_ld % 4 == 0);
    const bool a_io = (((uintptr_t)out & 15) == 0) && (out_ld % 4 == 0) &&
                      (!res || ((((uintptr_t)res & 15) == 0) && (res_ld % 4 == 0)));
    return (a_in ? 1 : 0) | (a_io ? 2 : 0);
}

template <typename K>
static int prepare_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
            return PCGC_ERR_CUDA;
        }
    }
    return PCGC_OK;
}

static inline int ctas_per_sm_for(size_t smem) {
    if (smem <= 24 * 1024) return 4;          // 256-thread CTAs, <=128 regs/thread: at most 4..8 resident
    if (smem <= 56 * 1024) return 3;
    if (smem <= 100 * 1024) return 2;
    return 1;
}

template <int CIN, int COUT, int KVOL>
static int launch_rowlane(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *w, const float *b,
                          const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s,
                          uint32_t *oh = nullptr, int oh_ld = 0, int *ovf = nullptr) {
    using R = RowLane<CIN, COUT>;
    // the h2 copy needs channel pairs per storing lane, or one channel per lane with lane l = channel l (quad gather, finish_row)
    if (oh && R::FC % 2 != 0 && !(R::FC == 1 && COUT == R::LPR && COUT % 4 == 0)) return kNotHandled;
    const size_t smem = R::weight_smem_bytes(KVOL);
    auto kern = conv_rowlane_kernel<CIN, COUT, KVOL>;
    int rc = prepare_smem(kern, smem);
    if (rc) return rc;
    const int rows_per_block = (kRowLaneThreads / 32) * R::RPW;
    kern<<<grid_for(n, rows_per_block, ctas_per_sm_for(smem)), kRowLaneThreads, smem, s>>>(
        in, in_ld, nbr, n, w, b, res, res_ld, out, out_ld, flags, aligned_bits(in, in_ld, out, out_ld, res, res_ld), oh, oh_ld, ovf);
    return check_launch("conv_rowlane");
}

template <int CIN, int COUT>
static int launch_tile(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *w, const float *b,
                       const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s) {
    using T = TileCfg<CIN, COUT>;
    const size_t smem = T::smem_bytes();
    auto kern = conv_k3_tile_kernel<CIN, COUT>;
    int rc = prepare_smem(kern, smem);
    if (rc) return rc;
    kern<<<grid_for(n, T::TM, ctas_per_sm_for(smem)), T::THREADS, smem, s>>>(in, in_ld, nbr, n, w, b, res, res_ld, out,
                                                                            out_ld, flags);
    return check_launch("conv_k3_tile");
}

template <int CIN, int COUT>
static int launch_down(const float *in, int in_ld, const uint64_t *keys, const int32_t *rows, const int32_t *off,
                       int64_t np, const float *w, const float *b, float *out, int out_ld, int flags, cudaStream_t s,
                       uint32_t *oh, int oh_ld, int *ovf) {
    using R = RowLane<CIN, COUT>;
    if (oh && R::FC % 2 != 0) return kNotHandled;
    const size_t smem = R::weight_smem_bytes(8);
    auto kern = conv_down_rowlane_kernel<CIN, COUT>;
    int rc = prepare_smem(kern, smem);
    if (rc) return rc;
    const int rows_per_block = (kRowLaneThreads / 32) * R::RPW;
    kern<<<grid_for(np, rows_per_block, ctas_per_sm_for(smem)), kRowLaneThreads, smem, s>>>(
        in, in_ld, keys, rows, off, np, w, b, out, out_ld, flags, aligned_bits(in, in_ld, out, out_ld, nullptr, 0), oh, oh_ld, ovf);
    return check_launch("conv_down_rowlane");
}

template <int CIN, int COUT>
static int launch_up(const float *in, int in_ld, int64_t n_in, const float *w, const float *b, float *out, int out_ld,
                     int flags, cudaStream_t s, uint32_t *oh, int oh_ld, int *ovf) {
    using R = RowLane<CIN, COUT>;
    if (oh && R::FC % 2 != 0) return kNotHandled;
    const size_t smem = R::weight_smem_bytes(8);
    auto kern = conv_up_rowlane_kernel<CIN, COUT>;
    int rc = prepare_smem(kern, smem);
    if (rc) return rc;
    const int rows_per_block = (kRowLaneThreads / 32) * R::RPW;
    kern<<<grid_for(n_in, rows_per_block, ctas_per_sm_for(smem)), kRowLaneThreads, smem, s>>>(
        in, in_ld, n_in, w, b, out, out_ld, flags, aligned_bits(in, in_ld, out, out_ld, nullptr, 0), oh, oh_ld, ovf);
    return check_launch("conv_up_rowlane");
}


constexpr int CI = PCGC_CI;

template <int CO>
static int k3_case(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *w, const float *b,
                   const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s) {
    if constexpr (CI <= 64 && CO <= 64 && RowLane<CI, CO>::weight_smem_bytes(27) <= kRowLaneSmemLimit) {
        return launch_rowlane<CI, CO, 27>(in, in_ld, nbr, n, w, b, res, res_ld, out, out_ld, flags, s);
    } else if constexpr (CI % 4 == 0 && CI >= 32 && CO % 32 == 0) {
        const bool tile_ok = (in_ld % 4 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)w & 15) == 0);
        if (tile_ok) return launch_tile<CI, CO>(in, in_ld, nbr, n, w, b, res, res_ld, out, out_ld, flags, s);
    }
    return kNotHandled;
}

template <int CO>
static int k1_case(const float *in, int in_ld, int64_t n, const float *w, const float *b, const float *res, int res_ld,
                   float *out, int out_ld, int flags, cudaStream_t s, uint32_t *oh, int oh_ld, int *ovf) {
    if constexpr (CI >= 4)
        return launch_rowlane<CI, CO, 1>(in, in_ld, nullptr, n, w, b, res, res_ld, out, out_ld, flags, s, oh, oh_ld, ovf);
    return kNotHandled;
}

template <int CO>
static int down_case(const float *in, int in_ld, const uint64_t *keys, const int32_t *rows, const int32_t *off,
                     int64_t np, const float *w, const float *b, float *out, int out_ld, int flags, cudaStream_t s,
                     uint32_t *oh, int oh_ld, int *ovf) {
    if constexpr (CI >= 8 && CO >= 8 && RowLane<CI, CO>::weight_smem_bytes(8) <= kRowLaneSmemLimit)
        return launch_down<CI, CO>(in, in_ld, keys, rows, off, np, w, b, out, out_ld, flags, s, oh, oh_ld, ovf);
    return kNotHandled;
}

template <int CO>
static int up_case(const float *in, int in_ld, int64_t n_in, const float *w, const float *b, float *out, int out_ld,
                   int flags, cudaStream_t s, uint32_t *oh, int oh_ld, int *ovf) {
    if constexpr (CI >= 8 && CO >= 8 && RowLane<CI, CO>::weight_smem_bytes(8) <= kRowLaneSmemLimit)
        return launch_up<CI, CO>(in, in_ld, n_in, w, b, out, out_ld, flags, s, oh, oh_ld, ovf);
    return kNotHandled;
}

template <int CO>
static int mma_case(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *packed, const float *b,
                    const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s) {
    if constexpr ((CI == 8 || CI == 16 || CI == 32 || CI == 64) && CO <= 64 && pipe_route(CI, CO) != kRouteMma) {
        using T = PipeTune<CI, CO>;
        using C = PipeCfg<CI, CO, T::NT, T::RG, T::D>;
        const size_t smem = C::smem_bytes();
        auto kern = conv_k3_pipe_kernel<CI, CO, T::NT, T::RG, T::D, T::MINB, T::OPT>;
        int rc = prepare_smem(kern, smem);
        if (rc) return rc;
        static int ctas = 0;                               // resident CTAs per SM of this instantiation
        if (ctas == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, smem) != cudaSuccess || nb < 1) nb = 1;
            ctas = nb;
        }
        kern<<<grid_for(n, C::ROWS_PER_CTA, ctas), C::THREADS, smem, s>>>(in, in_ld, nbr, n, packed, b, res, res_ld, out,
                                                                         out_ld, flags);
        return check_launch("conv_k3_pipe");
    } else if constexpr ((CI == 8 || CI == 16 || CI == 32 || CI == 64) && CO <= 64) {
        using C = MmaCfg<CI, CO>;
        const size_t smem = C::smem_bytes();
        auto kern = conv_k3_mma_kernel<CI, CO>;
        int rc = prepare_smem(kern, smem);
        if (rc) return rc;
        static int ctas = 0;                               // resident CTAs per SM of this instantiation
        if (ctas == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, smem) != cudaSuccess || nb < 1) nb = 1;
            ctas = nb;
        }
        kern<<<grid_for(n, C::ROWS_PER_CTA, ctas), C::THREADS, smem, s>>>(in, in_ld, nbr, n, packed, b, res, res_ld, out,
                                                                         out_ld, flags);
        return check_launch("conv_k3_mma");
    }
    return kNotHandled;
}

#define PCGC_FOR_CO(X) X(1) X(4) X(8) X(16) X(32) X(64) X(128)
#define PCGC_CAT2(a, b) a##b
#define PCGC_CAT(a, b) PCGC_CAT2(a, b)

int PCGC_CAT(k3_ci, PCGC_CI)(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *w, const float *b,
                             int cout, const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s) {
#define CASE(CO) if (cout == CO) return k3_case<CO>(in, in_ld, nbr, n, w, b, res, res_ld, out, out_ld, flags, s);
    PCGC_FOR_CO(CASE)
#undef CASE
    return kNotHandled;
}

int PCGC_CAT(mma_ci, PCGC_CI)(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *packed,
                              const float *b, int cout, const float *res, int res_ld, float *out, int out_ld,
                              int flags, cudaStream_t s) {
#define CASE(CO) if (cout == CO) return mma_case<CO>(in, in_ld, nbr, n, packed, b, res, res_ld, out, out_ld, flags, s);
    PCGC_FOR_CO(CASE)
#undef CASE
    return kNotHandled;
}

int PCGC_CAT(k1_ci, PCGC_CI)(const float *in, int in_ld, int64_t n, const float *w, const float *b, int cout,
                             const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s,
                             uint32_t *oh, int oh_ld, int *ovf) {
#define CASE(CO) if (cout == CO) return k1_case<CO>(in, in_ld, n, w, b, res, res_ld, out, out_ld, flags, s, oh, oh_ld, ovf);
    PCGC_FOR_CO(CASE)
#undef CASE
    return kNotHandled;
}

int PCGC_CAT(down_ci, PCGC_CI)(const float *in, int in_ld, const uint64_t *keys, const int32_t *rows,
                               const int32_t *off, int64_t np, const float *w, const float *b, int cout, float *out,
                               int out_ld, int flags, cudaStream_t s, uint32_t *oh, int oh_ld, int *ovf) {
#define CASE(CO) if (cout == CO) return down_case<CO>(in, in_ld, keys, rows, off, np, w, b, out, out_ld, flags, s, oh, oh_ld, ovf);
    PCGC_FOR_CO(CASE)
#undef CASE
    return kNotHandled;
}

int PCGC_CAT(up_ci, PCGC_CI)(const float *in, int in_ld, int64_t n_in, const float *w, const float *b, int cout,
                             float *out, int out_ld, int flags, cudaStream_t s, uint32_t *oh, int oh_ld, int *ovf) {
#define CASE(CO) if (cout == CO) return up_case<CO>(in, in_ld, n_in, w, b, out, out_ld, flags, s, oh, oh_ld, ovf);
    PCGC_FOR_CO(CASE)
#undef CASE
    return kNotHandled;
}

}  // namespace pcgc
