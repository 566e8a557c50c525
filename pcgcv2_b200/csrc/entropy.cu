// entropy.cu -- factorised-prior EntropyBottleneck evaluation on the GPU
// (reference entropy_model.py:82-130,151-196).
#include "common.cuh"

namespace pcgc {

constexpr int PPC = PCGC_EB_PARAMS_PER_CHANNEL;
// offsets inside one channel's parameter block
constexpr int M0 = 0, M1 = 3, M2 = 12, M3 = 21, B0 = 24, B1 = 27, B2 = 30, B3 = 33, F0 = 34, F1 = 37, F2 = 40, F3 = 43;

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch threshold 20
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// raw parameters -> softplus(matrices), biases, tanh(factors)  (entropy_model.py:95-99)
__device__ __forceinline__ void transform_params(const float *__restrict__ raw, float *__restrict__ tp, int channels) {
    for (int i = threadIdx.x; i < channels * PPC; i += blockDim.x) {
        const int o = i % PPC;
        const float v = raw[i];
        tp[i] = o < B0 ? softplus_f(v) : (o < F0 ? v : (o < 44 ? tanhf(v) : 0.f));
    }
}

// logits of the cumulative density of one channel at x  (entropy_model.py:82-101)
__device__ __forceinline__ float logits_cumulative(const float *__restrict__ p, float x) {
    float h[3], g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = fmaf(p[M0 + j], x, p[B0 + j]);
        h[j] = v + p[F0 + j] * tanhf(v);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = p[M1 + 3 * j] * h[0];
        v = fmaf(p[M1 + 3 * j + 1], h[1], v);
        v = fmaf(p[M1 + 3 * j + 2], h[2], v);
        v += p[B1 + j];
        g[j] = v + p[F1 + j] * tanhf(v);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = p[M2 + 3 * j] * g[0];
        v = fmaf(p[M2 + 3 * j + 1], g[1], v);
        v = fmaf(p[M2 + 3 * j + 2], g[2], v);
        v += p[B2 + j];
        h[j] = v + p[F2 + j] * tanhf(v);
    }
    float v = p[M3] * h[0];
    v = fmaf(p[M3 + 1], h[1], v);
    v = fmaf(p[M3 + 2], h[2], v);
    v += p[B3];
    return v + p[F3] * tanhf(v);
}

// |sigmoid(s*upper) - sigmoid(s*lower)|, s = -sign(lower + upper)  (entropy_model.py:121-125)
__device__ __forceinline__ float likelihood_at(const float *__restrict__ p, float x) {
    const float lower = logits_cumulative(p, x - 0.5f);
    const float upper = logits_cumulative(p, x + 0.5f);
    const float sum = lower + upper;
    const float s = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
    return fabsf(sigmoid_f(s * upper) - sigmoid_f(s * lower));
}

__global__ void eb_likelihood_kernel(const float *__restrict__ values, int64_t total, int channels,
                                     const float *__restrict__ raw, float *__restrict__ lik) {
    extern __shared__ float tp[];
    transform_params(raw, tp, channels);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        lik[i] = likelihood_at(tp + (int)(i % channels) * PPC, __ldg(values + i));
}

// one block per channel: pmf over the symbol grid, sequential cumsum (double accumulator, as the
// reference's CPU torch.cumsum), clamp, and the torchac integer table.
__global__ void eb_cdf_table_kernel(const float *__restrict__ raw, int channels, int min_v, int L,
                                    float *__restrict__ cdf_float, uint16_t *__restrict__ cdf_u16) {
    extern __shared__ float sm[];
    float *tp = sm;                 // [PPC]
    float *pmf = sm + PPC;          // [L]
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < PPC; i += blockDim.x) {
        const float v = raw[c * PPC + i];
        tp[i] = i < B0 ? softplus_f(v) : (i < F0 ? v : (i < 44 ? tanhf(v) : 0.f));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) pmf[j] = fmaxf(likelihood_at(tp, (float)(min_v + j)), 1e-9f);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int Lp = L + 1;
        const float scale = (float)(65536 - (Lp - 1));
        double run = 0.0;
        for (int j = 0; j <= L; ++j) {
            if (j > 0) run += (double)pmf[j - 1];
            const float cdf = fminf((float)run, 1.f);
            cdf_float[c * Lp + j] = cdf;
            if (cdf_u16) cdf_u16[c * Lp + j] = (uint16_t)((int)rintf(cdf * scale) + j);   // Appendix B.1 (wraps at 65536)
        }
    }
}

__global__ void eb_round_minmax_kernel(const float *__restrict__ x, int64_t count, int32_t *__restrict__ minmax) {
    int lo = 0x7FFFFFFF, hi = (int)0x80000000;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)rintf(__ldg(x + i));          // torch.round: half to even
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax, lo);
        atomicMax(minmax + 1, hi);
    }
}

__global__ void eb_symbols_kernel(const float *__restrict__ x, int64_t count, const int32_t *__restrict__ minmax,
                                  int16_t *__restrict__ sym) {
    const int lo = minmax[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        sym[i] = (int16_t)((int)rintf(__ldg(x + i)) - lo);
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

int pcgc_eb_likelihood_fwd(const float *values, int64_t n, int32_t channels, const float *params, float *likelihood,
                           void *stream) {
    PCGC_REQUIRE(n >= 0 && channels >= 1 && channels <= 256, "pcgc_eb_likelihood_fwd: bad shape");
    if (n == 0) return PCGC_OK;
    const int64_t total = n * channels;
    eb_likelihood_kernel<<<grid_for(total, 256, 4), 256, sizeof(float) * channels * PPC, (cudaStream_t)stream>>>(
        values, total, channels, params, likelihood);
    return check_launch("eb_likelihood");
}

int pcgc_eb_cdf_table(const float *params, int32_t channels, int32_t min_v, int32_t max_v, float *cdf_float,
                      uint16_t *cdf_u16, void *stream) {
    const int64_t L = (int64_t)max_v - min_v + 1;
    PCGC_REQUIRE(channels >= 1 && L >= 1 && L <= 8192, "pcgc_eb_cdf_table: bad symbol range [%d, %d]", min_v, max_v);
    eb_cdf_table_kernel<<<channels, 128, sizeof(float) * (PPC + L), (cudaStream_t)stream>>>(params, channels, min_v, (int)L,
                                                                                          cdf_float, cdf_u16);
    return check_launch("eb_cdf_table");
}

int pcgc_eb_round_minmax(const float *feats, int64_t count, int32_t *minmax, void *stream) {
    PCGC_REQUIRE(count >= 0, "pcgc_eb_round_minmax: bad count");
    if (count == 0) return PCGC_OK;
    eb_round_minmax_kernel<<<grid_for(count, 256, 4), 256, 0, (cudaStream_t)stream>>>(feats, count, minmax);
    return check_launch("eb_round_minmax");
}

int pcgc_eb_symbols(const float *feats, int64_t count, const int32_t *minmax, int16_t *sym, void *stream) {
    PCGC_REQUIRE(count >= 0, "pcgc_eb_symbols: bad count");
    if (count == 0) return PCGC_OK;
    eb_symbols_kernel<<<grid_for(count, 256, 4), 256, 0, (cudaStream_t)stream>>>(feats, count, minmax, sym);
    return check_launch("eb_symbols");
}

}  // extern "C"
