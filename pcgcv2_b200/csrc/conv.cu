// conv.cu -- C-ABI entry points + kernel dispatch for the sparse convolutions
// (SURVEY section 8 rows a3, a5, a6, a7).
#include "common.cuh"
#include "conv_dispatch.h"
#include "conv_mma.cuh"
#include "conv_pipe.cuh"

namespace pcgc {

// generic fallback for channel counts the specialised kernels do not cover (correct, slow):
// one thread per (row, output channel), weights from global memory.
__global__ void conv_generic_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr,
                                    int64_t n, int kvol, const float *__restrict__ weight,
                                    const float *__restrict__ bias, int cin, int cout,
                                    const float *__restrict__ residual, int res_ld, float *__restrict__ out,
                                    int out_ld, int flags) {
    const int64_t total = n * cout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / cout;
        const int co = (int)(i % cout);
        float acc = 0.f;
        for (int k = 0; k < kvol; ++k) {
            const int64_t src = nbr ? nbr[(int64_t)k * n + row] : row;
            if (src < 0) continue;
            const float *x = in + src * in_ld;
            const float *w = weight + ((int64_t)k * cin) * cout + co;
            for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], w[(int64_t)ci * cout], acc);
        }
        if (bias) acc += bias[co];
        if (residual) acc += residual[row * res_ld + co];
        if (flags & PCGC_EPI_RELU) acc = fmaxf(acc, 0.f);
        out[row * out_ld + co] = acc;
    }
}

__global__ void conv_down_generic_kernel(const float *__restrict__ in, int in_ld, const uint64_t *__restrict__ in_keys,
                                         const int32_t *__restrict__ child_rows, const int32_t *__restrict__ child_off,
                                         int64_t n_parents, const float *__restrict__ weight,
                                         const float *__restrict__ bias, int cin, int cout, float *__restrict__ out,
                                         int out_ld, int flags) {
    const int64_t total = n_parents * cout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / cout;
        const int co = (int)(i % cout);
        float acc = 0.f;
        for (int j = child_off[row]; j < child_off[row + 1]; ++j) {
            const int64_t src = child_rows[j];
            const int k = (int)(in_keys[src] & 7);
            const float *x = in + src * in_ld;
            const float *w = weight + ((int64_t)k * cin) * cout + co;
            for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], w[(int64_t)ci * cout], acc);
        }
        if (bias) acc += bias[co];
        if (flags & PCGC_EPI_RELU) acc = fmaxf(acc, 0.f);
        out[row * out_ld + co] = acc;
    }
}

__global__ void conv_up_generic_kernel(const float *__restrict__ in, int in_ld, int64_t n_in,
                                       const float *__restrict__ weight, const float *__restrict__ bias, int cin,
                                       int cout, float *__restrict__ out, int out_ld, int flags) {
    const int64_t total = n_in * 8 * cout;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t orow = i / cout;
        const int co = (int)(i % cout);
        const int k = (int)(orow & 7);
        const float *x = in + (orow >> 3) * in_ld;
        const float *w = weight + ((int64_t)k * cin) * cout + co;
        float acc = 0.f;
        for (int ci = 0; ci < cin; ++ci) acc = fmaf(x[ci], w[(int64_t)ci * cout], acc);
        if (bias) acc += bias[co];
        if (flags & PCGC_EPI_RELU) acc = fmaxf(acc, 0.f);
        out[orow * out_ld + co] = acc;
    }
}

}  // namespace pcgc

using namespace pcgc;

static int check_conv_args(const char *who, const void *in, const void *w, const void *out, int64_t n, int cin, int cout,
                           int in_ld, int out_ld) {
    PCGC_REQUIRE(n >= 0 && cin >= 1 && cout >= 1 && in_ld >= cin && out_ld >= cout,
                 "%s: bad shape n=%lld cin=%d cout=%d ld=%d/%d", who, (long long)n, cin, cout, in_ld, out_ld);
    PCGC_REQUIRE(n == 0 || (in && w && out), "%s: null pointer", who);
    return PCGC_OK;
}

extern "C" {

int pcgc_conv_k3_fwd(const float *in, int32_t in_ld, const int32_t *nbr, int64_t n, const float *weight,
                     const float *bias, int32_t cin, int32_t cout, const float *residual, int32_t res_ld, float *out,
                     int32_t out_ld, int32_t flags, void *stream) {
    int rc = check_conv_args("pcgc_conv_k3_fwd", in, weight, out, n, cin, cout, in_ld, out_ld);
    if (rc || n == 0) return rc;
    PCGC_REQUIRE(nbr != nullptr, "pcgc_conv_k3_fwd: null kernel map");
    cudaStream_t s = (cudaStream_t)stream;
    rc = kNotHandled;
#define CASE(CI) if (cin == CI) rc = k3_ci##CI(in, in_ld, nbr, n, weight, bias, cout, residual, res_ld, out, out_ld, flags, s);
    PCGC_FOR_CI(CASE)
#undef CASE
    if (rc != kNotHandled) return rc;
    conv_generic_kernel<<<grid_for(n * cout, 256, 8), 256, 0, s>>>(in, in_ld, nbr, n, 27, weight, bias, cin, cout,
                                                                   residual, res_ld, out, out_ld, flags);
    return check_launch("conv_generic");
}

size_t pcgc_conv_k3_packed_floats(int32_t cin, int32_t cout) {
    const bool ok = (cin == 8 || cin == 16 || cin == 32 || cin == 64) &&
                    (cout == 1 || cout == 4 || cout == 8 || cout == 16 || cout == 32 || cout == 64);
    if (!ok) return 0;
    if (pipe_route(cin, cout) == kRoutePipeNT) return (size_t)27 * (cin / 8) * ((cout + 7) / 8) * 128;
    return (size_t)27 * (cin / 8) * ((cout + 15) / 16) * 256;
}

int pcgc_conv_k3_pack_weights(const float *weight, int32_t cin, int32_t cout, float *packed, void *stream) {
    const size_t total = pcgc_conv_k3_packed_floats(cin, cout);
    PCGC_REQUIRE(total > 0 && weight && packed, "pcgc_conv_k3_pack_weights: no tensor-core kernel for %dx%d", cin, cout);
    // the fragment layout follows the kernel that will consume it (conv_pipe.cuh: pipe_route)
    const int grid = grid_for((int64_t)total / 2, 256, 4);
    switch (pipe_route(cin, cout)) {
        case kRoutePipeNT: pack_weights_nt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(weight, 27, cin, cout, packed); break;
        case kRoutePipeT: pack_weights_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(weight, 27, cin, cout, packed); break;
        default: pack_weights_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(weight, 27, cin, cout, packed); break;
    }
    return check_launch("pack_weights_mma");
}

int pcgc_conv_k3_fwd_packed(const float *in, int32_t in_ld, const int32_t *nbr, int64_t n, const float *packed,
                            const float *bias, int32_t cin, int32_t cout, const float *residual, int32_t res_ld,
                            float *out, int32_t out_ld, int32_t flags, void *stream) {
    int rc = check_conv_args("pcgc_conv_k3_fwd_packed", in, packed, out, n, cin, cout, in_ld, out_ld);
    if (rc || n == 0) return rc;
    PCGC_REQUIRE(nbr != nullptr, "pcgc_conv_k3_fwd_packed: null kernel map");
    PCGC_REQUIRE(pcgc_conv_k3_packed_floats(cin, cout) > 0, "pcgc_conv_k3_fwd_packed: no tensor-core kernel for %dx%d", cin, cout);
    PCGC_REQUIRE((in_ld % 4 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)packed & 15) == 0),
                 "pcgc_conv_k3_fwd_packed: input rows must be 16-byte aligned (ld %% 4 == 0)");
    cudaStream_t s = (cudaStream_t)stream;
    rc = kNotHandled;
#define CASE(CI) if (cin == CI) rc = mma_ci##CI(in, in_ld, nbr, n, packed, bias, cout, residual, res_ld, out, out_ld, flags, s);
    PCGC_FOR_CI(CASE)
#undef CASE
    PCGC_REQUIRE(rc != kNotHandled, "pcgc_conv_k3_fwd_packed: unsupported shape %dx%d", cin, cout);
    return rc;
}

static int conv_k1_impl(const char *who, const float *in, int32_t in_ld, int64_t n, const float *weight, const float *bias,
                        int32_t cin, int32_t cout, const float *residual, int32_t res_ld, float *out, int32_t out_ld,
                        int32_t flags, uint32_t *out_h2, int32_t out_h2_ld, int32_t *overflow, void *stream) {
    int rc = check_conv_args(who, in, weight, out, n, cin, cout, in_ld, out_ld);
    if (rc || n == 0) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    rc = kNotHandled;
#define CASE(CI) if (cin == CI) rc = k1_ci##CI(in, in_ld, n, weight, bias, cout, residual, res_ld, out, out_ld, flags, s, out_h2, out_h2_ld, overflow);
    PCGC_FOR_CI(CASE)
#undef CASE
    if (rc != kNotHandled) return rc;
    PCGC_REQUIRE(!out_h2, "%s: no h2 output for shape %dx%d", who, cin, cout);
    conv_generic_kernel<<<grid_for(n * cout, 256, 8), 256, 0, s>>>(in, in_ld, nullptr, n, 1, weight, bias, cin, cout,
                                                                   residual, res_ld, out, out_ld, flags);
    return check_launch("conv_generic");
}

static int conv_k2s2_impl(const char *who, const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *child_rows,
                          const int32_t *child_off, int64_t n_parents, const float *weight, const float *bias, int32_t cin,
                          int32_t cout, float *out, int32_t out_ld, int32_t flags, uint32_t *out_h2, int32_t out_h2_ld,
                          int32_t *overflow, void *stream) {
    int rc = check_conv_args(who, in, weight, out, n_parents, cin, cout, in_ld, out_ld);
    if (rc || n_parents == 0) return rc;
    PCGC_REQUIRE(in_keys && child_rows && child_off, "%s: null map", who);
    cudaStream_t s = (cudaStream_t)stream;
    rc = kNotHandled;
#define CASE(CI) if (cin == CI) rc = down_ci##CI(in, in_ld, in_keys, child_rows, child_off, n_parents, weight, bias, cout, out, out_ld, flags, s, out_h2, out_h2_ld, overflow);
    PCGC_FOR_CI(CASE)
#undef CASE
    if (rc != kNotHandled) return rc;
    PCGC_REQUIRE(!out_h2, "%s: no h2 output for shape %dx%d", who, cin, cout);
    conv_down_generic_kernel<<<grid_for(n_parents * cout, 256, 8), 256, 0, s>>>(
        in, in_ld, in_keys, child_rows, child_off, n_parents, weight, bias, cin, cout, out, out_ld, flags);
    return check_launch("conv_down_generic");
}

static int convT_k2s2_impl(const char *who, const float *in, int32_t in_ld, int64_t n_in, const float *weight, const float *bias,
                           int32_t cin, int32_t cout, float *out, int32_t out_ld, int32_t flags, uint32_t *out_h2,
                           int32_t out_h2_ld, int32_t *overflow, void *stream) {
    int rc = check_conv_args(who, in, weight, out, n_in, cin, cout, in_ld, out_ld);
    if (rc || n_in == 0) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    rc = kNotHandled;
#define CASE(CI) if (cin == CI) rc = up_ci##CI(in, in_ld, n_in, weight, bias, cout, out, out_ld, flags, s, out_h2, out_h2_ld, overflow);
    PCGC_FOR_CI(CASE)
#undef CASE
    if (rc != kNotHandled) return rc;
    PCGC_REQUIRE(!out_h2, "%s: no h2 output for shape %dx%d", who, cin, cout);
    conv_up_generic_kernel<<<grid_for(n_in * 8 * cout, 256, 8), 256, 0, s>>>(in, in_ld, n_in, weight, bias, cin, cout,
                                                                            out, out_ld, flags);
    return check_launch("conv_up_generic");
}

// shapes whose row-lane kernel leaves every storing lane with an even number of finished channels (conv_rowlane.cuh)
static bool rowlane_h2out_shape(int kind, int cin, int cout) {
    auto listed = [](int c) { return c == 4 || c == 8 || c == 16 || c == 32 || c == 64 || c == 128; };
    if (!listed(cin) || !listed(cout)) return false;
    if (kind != 1 && (cin < 8 || cout < 8 || (size_t)8 * cin * cout * sizeof(float) > 160 * 1024)) return false;
    int lpr = cin / 4, log_lpr = 0, tz = 0;
    while ((1 << (log_lpr + 1)) <= lpr) ++log_lpr;
    while (((cout >> tz) & 1) == 0) ++tz;
    const int hs = log_lpr < tz ? log_lpr : tz;
    if ((cout >> hs) == 1 && cout == lpr && cout % 4 == 0) return true;      // one channel per lane, lane l = channel l: quad gather
    return ((cout >> hs) % 2) == 0;
}

static int check_h2_out(const char *who, const uint32_t *out_h2, int32_t out_h2_ld, int32_t cout) {
    PCGC_REQUIRE(out_h2 && cout % 4 == 0 && out_h2_ld >= cout && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0,
                 "%s: the h2 output needs cout %% 4 == 0 and 16-byte aligned rows", who);
    return PCGC_OK;
}

int pcgc_conv_k1_fwd(const float *in, int32_t in_ld, int64_t n, const float *weight, const float *bias, int32_t cin,
                     int32_t cout, const float *residual, int32_t res_ld, float *out, int32_t out_ld, int32_t flags,
                     void *stream) {
    return conv_k1_impl("pcgc_conv_k1_fwd", in, in_ld, n, weight, bias, cin, cout, residual, res_ld, out, out_ld, flags, nullptr, 0,
                        nullptr, stream);
}

int pcgc_conv_k2s2_fwd(const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *child_rows,
                       const int32_t *child_off, int64_t n_parents, const float *weight, const float *bias,
                       int32_t cin, int32_t cout, float *out, int32_t out_ld, int32_t flags, void *stream) {
    return conv_k2s2_impl("pcgc_conv_k2s2_fwd", in, in_ld, in_keys, child_rows, child_off, n_parents, weight, bias, cin, cout, out,
                          out_ld, flags, nullptr, 0, nullptr, stream);
}

int pcgc_convT_k2s2_fwd(const float *in, int32_t in_ld, int64_t n_in, const float *weight, const float *bias,
                        int32_t cin, int32_t cout, float *out, int32_t out_ld, int32_t flags, void *stream) {
    return convT_k2s2_impl("pcgc_convT_k2s2_fwd", in, in_ld, n_in, weight, bias, cin, cout, out, out_ld, flags, nullptr, 0, nullptr,
                           stream);
}

int pcgc_conv_h2out_supported(int32_t kind, int32_t cin, int32_t cout) { return rowlane_h2out_shape(kind, cin, cout) ? 1 : 0; }

// the same three layers writing, next to the fp32 output, its pre-split half-precision copy for a following h2 k=3 layer
int pcgc_conv_k1_fwd_h2out(const float *in, int32_t in_ld, int64_t n, const float *weight, const float *bias, int32_t cin,
                           int32_t cout, const float *residual, int32_t res_ld, float *out, int32_t out_ld, uint32_t *out_h2,
                           int32_t out_h2_ld, int32_t flags, int32_t *overflow, void *stream) {
    int rc = check_h2_out("pcgc_conv_k1_fwd_h2out", out_h2, out_h2_ld, cout);
    if (rc) return rc;
    return conv_k1_impl("pcgc_conv_k1_fwd_h2out", in, in_ld, n, weight, bias, cin, cout, residual, res_ld, out, out_ld, flags, out_h2,
                        out_h2_ld, overflow, stream);
}

int pcgc_conv_k2s2_fwd_h2out(const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *child_rows,
                             const int32_t *child_off, int64_t n_parents, const float *weight, const float *bias, int32_t cin,
                             int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags,
                             int32_t *overflow, void *stream) {
    int rc = check_h2_out("pcgc_conv_k2s2_fwd_h2out", out_h2, out_h2_ld, cout);
    if (rc) return rc;
    return conv_k2s2_impl("pcgc_conv_k2s2_fwd_h2out", in, in_ld, in_keys, child_rows, child_off, n_parents, weight, bias, cin, cout,
                          out, out_ld, flags, out_h2, out_h2_ld, overflow, stream);
}

int pcgc_convT_k2s2_fwd_h2out(const float *in, int32_t in_ld, int64_t n_in, const float *weight, const float *bias, int32_t cin,
                              int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags,
                              int32_t *overflow, void *stream) {
    int rc = check_h2_out("pcgc_convT_k2s2_fwd_h2out", out_h2, out_h2_ld, cout);
    if (rc) return rc;
    return convT_k2s2_impl("pcgc_convT_k2s2_fwd_h2out", in, in_ld, n_in, weight, bias, cin, cout, out, out_ld, flags, out_h2, out_h2_ld,
                           overflow, stream);
}

}  // extern "C"
