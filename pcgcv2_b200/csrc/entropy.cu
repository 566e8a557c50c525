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

// ---- backward of the likelihood (training path: loss.py:17-20 bits = -sum log2 p, trainer.py:131-136) -----------
// forward of one cumulative-logit path with every intermediate kept, then its reverse sweep.
struct EbPath {
    float z, v0[3], h0[3], v1[3], h1[3], v2[3], h2[3], v3;
};

__device__ __forceinline__ float eb_path_fwd(const float *__restrict__ p, float z, EbPath &s) {
    s.z = z;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        s.v0[j] = fmaf(p[M0 + j], z, p[B0 + j]);
        s.h0[j] = s.v0[j] + p[F0 + j] * tanhf(s.v0[j]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = p[M1 + 3 * j] * s.h0[0];
        v = fmaf(p[M1 + 3 * j + 1], s.h0[1], v);
        v = fmaf(p[M1 + 3 * j + 2], s.h0[2], v);
        s.v1[j] = v + p[B1 + j];
        s.h1[j] = s.v1[j] + p[F1 + j] * tanhf(s.v1[j]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = p[M2 + 3 * j] * s.h1[0];
        v = fmaf(p[M2 + 3 * j + 1], s.h1[1], v);
        v = fmaf(p[M2 + 3 * j + 2], s.h1[2], v);
        s.v2[j] = v + p[B2 + j];
        s.h2[j] = s.v2[j] + p[F2 + j] * tanhf(s.v2[j]);
    }
    float v = p[M3] * s.h2[0];
    v = fmaf(p[M3 + 1], s.h2[1], v);
    v = fmaf(p[M3 + 2], s.h2[2], v);
    s.v3 = v + p[B3];
    return s.v3 + p[F3] * tanhf(s.v3);
}

// g = dLoss/d(out of this path); accumulates d/d(transformed params) into gp[44], returns dLoss/dz
__device__ __forceinline__ float eb_path_bwd(const float *__restrict__ p, const EbPath &s, float g, float *__restrict__ gp) {
    float t = tanhf(s.v3);
    gp[F3] += g * t;
    const float dv3 = g * (1.f + p[F3] * (1.f - t * t));
    gp[B3] += dv3;
    float dh2[3], dh1[3] = {0.f, 0.f, 0.f}, dh0[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) { gp[M3 + i] += dv3 * s.h2[i]; dh2[i] = dv3 * p[M3 + i]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        t = tanhf(s.v2[j]);
        gp[F2 + j] += dh2[j] * t;
        const float dv = dh2[j] * (1.f + p[F2 + j] * (1.f - t * t));
        gp[B2 + j] += dv;
#pragma unroll
        for (int i = 0; i < 3; ++i) { gp[M2 + 3 * j + i] += dv * s.h1[i]; dh1[i] = fmaf(dv, p[M2 + 3 * j + i], dh1[i]); }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        t = tanhf(s.v1[j]);
        gp[F1 + j] += dh1[j] * t;
        const float dv = dh1[j] * (1.f + p[F1 + j] * (1.f - t * t));
        gp[B1 + j] += dv;
#pragma unroll
        for (int i = 0; i < 3; ++i) { gp[M1 + 3 * j + i] += dv * s.h0[i]; dh0[i] = fmaf(dv, p[M1 + 3 * j + i], dh0[i]); }
    }
    float dz = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        t = tanhf(s.v0[j]);
        gp[F0 + j] += dh0[j] * t;
        const float dv = dh0[j] * (1.f + p[F0 + j] * (1.f - t * t));
        gp[B0 + j] += dv;
        gp[M0 + j] += dv * s.z;
        dz = fmaf(dv, p[M0 + j], dz);
    }
    return dz;
}

// every thread keeps to ONE channel (blockDim and the grid stride are multiples of `channels`), accumulates the 44
// parameter gradients of that channel in registers, then block-reduces through shared memory.
__global__ void __launch_bounds__(256)
eb_likelihood_bwd_kernel(const float *__restrict__ values, int64_t total, int channels, const float *__restrict__ raw,
                         const float *__restrict__ grad_lik, float *__restrict__ grad_values,
                         float *__restrict__ grad_params) {
    extern __shared__ float sm[];
    float *tp = sm;                               // transformed params [channels][PPC]
    float *acc = sm + channels * PPC;             // block accumulators [channels][PPC]
    transform_params(raw, tp, channels);
    for (int i = threadIdx.x; i < channels * PPC; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int c = threadIdx.x % channels;
    const float *p = tp + c * PPC;
    float gp[44];
#pragma unroll
    for (int i = 0; i < 44; ++i) gp[i] = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(values + i), g = __ldg(grad_lik + i);
        EbPath lo, up;
        const float lower = eb_path_fwd(p, x - 0.5f, lo), upper = eb_path_fwd(p, x + 0.5f, up);
        const float sum = lower + upper;
        const float sg = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
        const float su = sigmoid_f(sg * upper), sl = sigmoid_f(sg * lower);
        const float d = su - sl, sd = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);      // d|d|/dd
        const float gu = g * sd * su * (1.f - su) * sg, gl = -g * sd * sl * (1.f - sl) * sg;
        const float dx = eb_path_bwd(p, up, gu, gp) + eb_path_bwd(p, lo, gl, gp);
        if (grad_values) grad_values[i] = dx;
    }
#pragma unroll
    for (int i = 0; i < 44; ++i)
        if (gp[i] != 0.f) atomicAdd(acc + c * PPC + i, gp[i]);
    __syncthreads();
    // chain rule to the raw parameters: softplus' = sigmoid, tanh' = 1 - tanh^2, biases pass through
    for (int i = threadIdx.x; i < channels * PPC; i += blockDim.x) {
        const int o = i % PPC;
        if (o >= 44 || acc[i] == 0.f) continue;
        const float r = raw[i];
        const float scale = o < B0 ? (r > 20.f ? 1.f : sigmoid_f(r)) : (o < F0 ? 1.f : 1.f - tp[i] * tp[i]);
        atomicAdd(grad_params + i, acc[i] * scale);
    }
}

// ---- the range coder's table (entropy_model.py:151-171 + torchac's float -> 16-bit conversion, Appendix B.1) ----------
// An arithmetic-coded stream desynchronises when ONE table entry differs by one count between encoder and decoder, so
// the table is not evaluated with the fast float32 intrinsics of the likelihood kernel above: every transcendental
// (softplus, tanh, sigmoid) is evaluated in float64 and the likelihood is rounded to float32 ONCE, then clamped, summed
// in a float64 accumulator (the reference's CPU torch.cumsum accumulates float32 rows in double), rounded to float32,
// clamped to 1 and converted exactly like torchac does (float32 multiply, round-half-even, + arange).  This reproduces
// the reference-generated golden tables exactly (tests/test_ops_gpu.py); the reference's own float32 CPU evaluation
// still differs from the correctly rounded value in ~0.1 % of the entries over all seven shipped checkpoints, which is
// why Codec builds the table it codes with on the host with the reference's own operator sequence (codec.py).
__device__ __forceinline__ double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

__device__ double logits_cumulative_d(const double *__restrict__ p, double x) {
    double h[3], g[3];
    for (int j = 0; j < 3; ++j) {
        const double v = p[M0 + j] * x + p[B0 + j];
        h[j] = v + p[F0 + j] * tanh(v);
    }
    for (int j = 0; j < 3; ++j) {
        const double v = p[M1 + 3 * j] * h[0] + p[M1 + 3 * j + 1] * h[1] + p[M1 + 3 * j + 2] * h[2] + p[B1 + j];
        g[j] = v + p[F1 + j] * tanh(v);
    }
    for (int j = 0; j < 3; ++j) {
        const double v = p[M2 + 3 * j] * g[0] + p[M2 + 3 * j + 1] * g[1] + p[M2 + 3 * j + 2] * g[2] + p[B2 + j];
        h[j] = v + p[F2 + j] * tanh(v);
    }
    const double v = p[M3] * h[0] + p[M3 + 1] * h[1] + p[M3 + 2] * h[2] + p[B3];
    return v + p[F3] * tanh(v);
}

// one block per channel
__global__ void eb_cdf_table_kernel(const float *__restrict__ raw, int channels, int min_v, int L,
                                    float *__restrict__ cdf_float, uint16_t *__restrict__ cdf_u16) {
    extern __shared__ double smd[];
    double *tp = smd;                                   // [PPC] transformed parameters, float64
    float *pmf = reinterpret_cast<float *>(smd + PPC);  // [L]
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < PPC; i += blockDim.x) {
        const double v = (double)raw[c * PPC + i];
        tp[i] = i < B0 ? softplus_d(v) : (i < F0 ? v : (i < 44 ? tanh(v) : 0.0));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const double x = (double)(min_v + j);
        const double lower = logits_cumulative_d(tp, x - 0.5), upper = logits_cumulative_d(tp, x + 0.5);
        const double sum = lower + upper;
        const double s = sum > 0.0 ? -1.0 : (sum < 0.0 ? 1.0 : 0.0);
        pmf[j] = fmaxf((float)fabs(sigmoid_d(s * upper) - sigmoid_d(s * lower)), 1e-9f);
    }
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

// per-symbol coding interval of the range coder (Appendix B.2): symbol i uses table row i % n_tables;
// ranges[i] = c_low | (c_high - 1) << 16 with c_high = 0x10000 for the last symbol of the alphabet
__global__ void symbol_ranges_kernel(const int16_t *__restrict__ sym, int64_t n, const uint16_t *__restrict__ cdf, int n_tables,
                                     int lp, uint32_t *__restrict__ ranges, int32_t *__restrict__ bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = sym[i];
        if (s < 0 || s > lp - 2) { *bad = 1; ranges[i] = 0; continue; }
        const uint16_t *row = cdf + (i % n_tables) * lp;
        const uint32_t lo = row[s], hi = s == lp - 2 ? 0x10000u : row[s + 1];
        ranges[i] = lo | ((hi - 1u) << 16);
    }
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

int pcgc_eb_likelihood_bwd(const float *values, int64_t n, int32_t channels, const float *params, const float *grad_likelihood,
                           float *grad_values, float *grad_params, void *stream) {
    PCGC_REQUIRE(n >= 0 && channels >= 1 && channels <= 64 && 256 % channels == 0,
                 "pcgc_eb_likelihood_bwd: channels must divide 256 (got %d)", channels);
    cudaStream_t s = (cudaStream_t)stream;
    PCGC_CUDA(cudaMemsetAsync(grad_params, 0, sizeof(float) * channels * PPC, s));
    if (n == 0) return PCGC_OK;
    const int64_t total = n * channels;
    eb_likelihood_bwd_kernel<<<grid_for(total, 256, 2), 256, sizeof(float) * 2 * channels * PPC, s>>>(
        values, total, channels, params, grad_likelihood, grad_values, grad_params);
    return check_launch("eb_likelihood_bwd");
}

int pcgc_eb_cdf_table(const float *params, int32_t channels, int32_t min_v, int32_t max_v, float *cdf_float,
                      uint16_t *cdf_u16, void *stream) {
    const int64_t L = (int64_t)max_v - min_v + 1;
    PCGC_REQUIRE(channels >= 1 && L >= 1 && L <= 8192, "pcgc_eb_cdf_table: bad symbol range [%d, %d]", min_v, max_v);
    eb_cdf_table_kernel<<<channels, 128, sizeof(double) * PPC + sizeof(float) * L, (cudaStream_t)stream>>>(params, channels, min_v, (int)L,
                                                                                          cdf_float, cdf_u16);
    return check_launch("eb_cdf_table");
}

int pcgc_eb_round_minmax(const float *feats, int64_t count, int32_t *minmax, void *stream) {
    PCGC_REQUIRE(count >= 0, "pcgc_eb_round_minmax: bad count");
    if (count == 0) return PCGC_OK;
    eb_round_minmax_kernel<<<grid_for(count, 256, 4), 256, 0, (cudaStream_t)stream>>>(feats, count, minmax);
    return check_launch("eb_round_minmax");
}

int pcgc_symbol_ranges(const int16_t *sym, int64_t n_sym, const uint16_t *cdf_u16, int32_t n_tables, int32_t lp,
                       uint32_t *ranges, int32_t *bad, void *stream) {
    PCGC_REQUIRE(n_sym >= 0 && n_tables >= 1 && lp >= 2 && cdf_u16 && bad, "pcgc_symbol_ranges: bad arguments");
    if (n_sym == 0) return PCGC_OK;
    symbol_ranges_kernel<<<grid_for(n_sym, 256, 4), 256, 0, (cudaStream_t)stream>>>(sym, n_sym, cdf_u16, n_tables, lp, ranges, bad);
    return check_launch("symbol_ranges");
}

int pcgc_eb_symbols(const float *feats, int64_t count, const int32_t *minmax, int16_t *sym, void *stream) {
    PCGC_REQUIRE(count >= 0, "pcgc_eb_symbols: bad count");
    if (count == 0) return PCGC_OK;
    eb_symbols_kernel<<<grid_for(count, 256, 4), 256, 0, (cudaStream_t)stream>>>(feats, count, minmax, sym);
    return check_launch("eb_symbols");
}

}  // extern "C"
