// conv_dispatch.h -- per-input-channel-count dispatch functions (one object file each).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcgc {

constexpr int kNotHandled = 1;
#define PCGC_FOR_CI(X) X(1) X(4) X(8) X(16) X(32) X(64) X(128)

#define PCGC_DECL(CI)                                                                                              \
    int k3_ci##CI(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *w, const float *b,      \
                  int cout, const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s);      \
    int mma_ci##CI(const float *in, int in_ld, const int32_t *nbr, int64_t n, const float *packed, const float *b, \
                   int cout, const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s);    \
    int k1_ci##CI(const float *in, int in_ld, int64_t n, const float *w, const float *b, int cout,                \
                  const float *res, int res_ld, float *out, int out_ld, int flags, cudaStream_t s,                 \
                  uint32_t *out_h2, int out_h2_ld, int *overflow);                                                 \
    int down_ci##CI(const float *in, int in_ld, const uint64_t *keys, const int32_t *rows, const int32_t *off,    \
                    int64_t np, const float *w, const float *b, int cout, float *out, int out_ld, int flags,       \
                    cudaStream_t s, uint32_t *out_h2, int out_h2_ld, int *overflow);                               \
    int up_ci##CI(const float *in, int in_ld, int64_t n_in, const float *w, const float *b, int cout, float *out, \
                  int out_ld, int flags, cudaStream_t s, uint32_t *out_h2, int out_h2_ld, int *overflow);
PCGC_FOR_CI(PCGC_DECL)
#undef PCGC_DECL

}  // namespace pcgc
