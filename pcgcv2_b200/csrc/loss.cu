// loss.cu -- fused occupancy loss of the training path.
//
// Reference: loss.py:7-15 `get_bce(data, ground_truth)`: mask = isin(data.C, ground_truth.C) on the HOST (D2H of both
// coordinate sets + np.isin), then torch BCEWithLogitsLoss over the mask, divided by ln 2, times the row count.  Here one
// pass over the candidate rows probes the ground-truth hash table (the same table the kernel maps use), evaluates the
// numerically stable binary cross entropy with logits, writes d loss / d logit for the backward pass and reduces the sum
// DETERMINISTICALLY: fixed grid, per-block partial sums in double, a second single-block pass adds them in block order.
#include "common.cuh"

namespace pcgc {

constexpr int kBceBlocks = 2 * kNumSMs, kBceThreads = 256;

__global__ void __launch_bounds__(kBceThreads)
bce_isin_kernel(const float *__restrict__ logits, int ld, const uint64_t *__restrict__ keys, int64_t n,
                const uint64_t *__restrict__ tkeys, uint64_t mask, float *__restrict__ grad_unit, uint8_t *__restrict__ target,
                double *__restrict__ partial) {
    __shared__ double red[kBceThreads / 32];
    const double inv_ln2 = 1.4426950408889634;
    double sum = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        uint64_t s = hash_slot(key, mask);
        bool hit = false;
        while (true) {
            const uint64_t k = __ldg(tkeys + s);
            if (k == key) { hit = true; break; }
            if (k == PCGC_EMPTY_KEY) break;
            s = (s + 1) & mask;
        }
        const float x = logits[i * ld], t = hit ? 1.f : 0.f;
        // max(x, 0) - x t + log1p(exp(-|x|))   (torch.nn.functional.binary_cross_entropy_with_logits)
        const float l = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        sum += (double)l;
        if (grad_unit) grad_unit[i] = (1.f / (1.f + expf(-x)) - t) * (float)inv_ln2;
        if (target) target[i] = hit ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < kBceThreads / 32; ++w) b += red[w];
        partial[blockIdx.x] = b * inv_ln2;
    }
}

__global__ void bce_finish_kernel(const double *__restrict__ partial, int blocks, float *__restrict__ loss_sum) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < blocks; ++b) s += partial[b];           // fixed order: bit-reproducible
        *loss_sum = (float)s;
    }
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

size_t pcgc_bce_isin_ws_bytes(void) { return sizeof(double) * kBceBlocks; }

int pcgc_bce_isin(const float *logits, int32_t ld, const uint64_t *cand_keys, int64_t n, const uint64_t *gt_table_keys, int64_t cap,
                  float *loss_sum_bits, float *grad_unit, uint8_t *target, void *ws, size_t ws_bytes, void *stream) {
    PCGC_REQUIRE(n >= 0 && ld >= 1 && cap >= 64 && (cap & (cap - 1)) == 0, "pcgc_bce_isin: bad arguments");
    PCGC_REQUIRE(ws && ws_bytes >= pcgc_bce_isin_ws_bytes(), "pcgc_bce_isin: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = n == 0 ? 1 : grid_for(n, kBceThreads, 2);
    PCGC_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * kBceBlocks, s));
    if (n > 0) {
        bce_isin_kernel<<<blocks, kBceThreads, 0, s>>>(logits, ld, cand_keys, n, gt_table_keys, (uint64_t)cap - 1, grad_unit, target,
                                                     (double *)ws);
        int rc = check_launch("bce_isin");
        if (rc) return rc;
    }
    bce_finish_kernel<<<1, 32, 0, s>>>((const double *)ws, blocks, loss_sum_bits);
    return check_launch("bce_finish");
}

}  // extern "C"
