// conv_rowlane.cuh -- "row-lane" sparse-convolution kernels for NARROW layers (all kernel
// weights resident in shared memory).  One output row is owned by LPR = CIN/4 adjacent
// lanes; lane l holds input channels [4l, 4l+4) of every gathered neighbour row (one 16-byte
// load per lane => a row is one contiguous 4*CIN-byte request) and accumulates a partial
// sum for ALL output channels in registers; a reduce-scatter over the LPR lanes (warp
// shuffles, no atomics, no read-modify-write of the output) leaves each lane with
// COUT/LPR finished channels that are written once, fused with bias / residual / ReLU.
//
// These layers are HBM/L2-bound (rows of 4..64 floats, SURVEY section 8 d): the design goal is
// "read each map entry once, gather rows through L1/L2, write each output once".
#pragma once
#include "common.cuh"
#include "conv_h2.cuh"   // split_pair_h2: optional pre-split half-precision copy of the output

namespace pcgc {

template <int CIN, int COUT>
struct RowLane {
    static_assert(CIN == 1 || CIN == 2 || CIN % 4 == 0, "CIN must be 1, 2 or a multiple of 4");
    static_assert(COUT == 1 || COUT == 2 || COUT % 4 == 0, "COUT must be 1, 2 or a multiple of 4");
    static constexpr int VEC = CIN >= 4 ? 4 : CIN;      // input channels per lane
    static constexpr int LPR = CIN / VEC;               // lanes per row
    static_assert(LPR >= 1 && LPR <= 32 && (LPR & (LPR - 1)) == 0, "CIN/4 must be a power of two <= 32");
    static constexpr int RPW = 32 / LPR;                // rows per warp
    static constexpr int CV = COUT >= 4 ? 4 : COUT;     // output vector width
    static constexpr int Q = COUT / CV;                 // output vectors per row
    static constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v / 2); }
    static constexpr int LOG_LPR = log2c(LPR);
    static constexpr int HS = LOG_LPR < log2c(COUT & -COUT) ? LOG_LPR : log2c(COUT & -COUT);  // halving steps
    static constexpr int AS = LOG_LPR - HS;             // all-reduce steps (lowest lane bits)
    static constexpr int FC = COUT >> HS;               // finished channels per storing lane
    static constexpr size_t weight_smem_bytes(int kvol) { return (size_t)kvol * CIN * COUT * sizeof(float); }
};

// Weights [kvol][CIN][COUT] -> shared memory, permuted so that for a fixed (k, c, q) the LPR
// lanes of a row read LPR consecutive CV-vectors (conflict-free; other rows broadcast):
//   smem index of (k, ci = VEC*l + c, co = CV*q + e) = (((k*VEC + c)*Q + q)*LPR + l)*CV + e
template <int CIN, int COUT>
__device__ __forceinline__ void load_weights_rowlane(const float *__restrict__ w, int kvol, float *__restrict__ ws) {
    using R = RowLane<CIN, COUT>;
    const int total = kvol * CIN * COUT;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int co = i % COUT, ci = (i / COUT) % CIN, k = i / (COUT * CIN);
        const int l = ci / R::VEC, c = ci % R::VEC, q = co / R::CV, e = co % R::CV;
        ws[(((k * R::VEC + c) * R::Q + q) * R::LPR + l) * R::CV + e] = __ldg(w + i);
    }
}

template <int VEC>
struct InVec { float v[VEC]; };

template <int VEC>
__device__ __forceinline__ InVec<VEC> load_in(const float *__restrict__ p, bool pred, bool aligned) {
    InVec<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = 0.f;
    if (pred) {
        if constexpr (VEC == 4) {
            if (aligned) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
                r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
                return r;
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) r.v[i] = __ldg(p + i);
    }
    return r;
}

// acc[COUT] += v (this lane's VEC input channels) x W[k] slice
template <int CIN, int COUT>
__device__ __forceinline__ void accum_rowlane(float (&acc)[COUT], const InVec<RowLane<CIN, COUT>::VEC> &v,
                                              const float *__restrict__ ws, int k, int l) {
    using R = RowLane<CIN, COUT>;
#pragma unroll
    for (int c = 0; c < R::VEC; ++c) {
        const float a = v.v[c];
#pragma unroll
        for (int q = 0; q < R::Q; ++q) {
            const float *p = ws + (((k * R::VEC + c) * R::Q + q) * R::LPR + l) * R::CV;
            if constexpr (R::CV == 4) {
                const float4 w4 = *reinterpret_cast<const float4 *>(p);
                acc[q * 4 + 0] = fmaf(a, w4.x, acc[q * 4 + 0]);
                acc[q * 4 + 1] = fmaf(a, w4.y, acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(a, w4.z, acc[q * 4 + 2]);
                acc[q * 4 + 3] = fmaf(a, w4.w, acc[q * 4 + 3]);
            } else {
#pragma unroll
                for (int e = 0; e < R::CV; ++e) acc[q * R::CV + e] = fmaf(a, p[e], acc[q * R::CV + e]);
            }
        }
    }
}

// reduce-scatter of acc[] over the LPR lanes of a row (xor shuffles stay inside the group)
template <int D, int CNT>
struct LaneReduce {
    template <int COUT>
    static __device__ __forceinline__ void run(float (&acc)[COUT], int l, int &base) {
        if constexpr (D >= 1) {
            const bool hi = (l & D) != 0;
            if constexpr (CNT >= 2 && CNT % 2 == 0) {
                constexpr int H = CNT / 2;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float mine = hi ? acc[H + i] : acc[i];
                    const float theirs = hi ? acc[i] : acc[H + i];
                    acc[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, D);
                }
                base += hi ? H : 0;
                LaneReduce<D / 2, H>::run(acc, l, base);
            } else {
#pragma unroll
                for (int i = 0; i < CNT; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], D);
                LaneReduce<D / 2, CNT>::run(acc, l, base);
            }
        }
    }
};

// epilogue for one output row: reduce over lanes, + bias (+ residual), ReLU, one store.
template <int CIN, int COUT>
__device__ __forceinline__ void finish_row(float (&acc)[COUT], int l, bool valid, int64_t row,
                                           const float *__restrict__ bias, const float *__restrict__ residual,
                                           int res_ld, float *__restrict__ out, int out_ld, int flags,
                                           bool io_aligned, uint32_t *__restrict__ out_h2 = nullptr, int out_h2_ld = 0,
                                           bool *over = nullptr) {
    using R = RowLane<CIN, COUT>;
    int base = 0;
    LaneReduce<R::LPR / 2, COUT>::run(acc, l, base);
    const bool store = valid && (l & ((1 << R::AS) - 1)) == 0;
    float *o = out + row * out_ld + base;
    const float *r = (residual && store) ? residual + row * res_ld + base : nullptr;
#pragma unroll
    for (int i = 0; i < R::FC; ++i) {
        float v = acc[i];
        if (bias) v += __ldg(bias + base + i);
        if (r) v += __ldg(r + i);
        if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
        acc[i] = v;
    }
    if constexpr (R::FC == 1 && COUT == R::LPR && COUT % 4 == 0) {
        // one finished channel per lane and lane l holds channel l (COUT == CIN / 4: conv1_0 of the InceptionResNet blocks):
        // lane pairs form the (hi, lo) pairs, the lane two further on holds the other pair of the 16-byte group
        if (out_h2) {                      // (uniform over the warp)
            const float other = __shfl_xor_sync(0xffffffffu, acc[0], 1);
            uint32_t hi, lo;
            split_pair_h2(acc[0], other, hi, lo);                                  // meaningful on even lanes
            const uint32_t phi = __shfl_xor_sync(0xffffffffu, hi, 2), plo = __shfl_xor_sync(0xffffffffu, lo, 2);
            if (store) {
                *over |= !(fabsf(acc[0]) <= kH2Limit);
                if ((base & 3) == 0) *reinterpret_cast<uint4 *>(out_h2 + row * out_h2_ld + base) = make_uint4(hi, phi, lo, plo);
            }
        }
    } else if constexpr (R::FC % 2 == 0) { // h2 copy for the next k=3 layer: 16-byte groups {hi01 hi23 lo01 lo23} (conv_h2.cuh)
        if (out_h2) {                      // (uniform over the warp)
            uint32_t *oh = out_h2 + row * out_h2_ld;
            if constexpr (R::FC % 4 == 0) {
                if (store) {
#pragma unroll
                    for (int i = 0; i < R::FC; i += 4) {
                        uint4 q;
                        split_pair_h2(acc[i], acc[i + 1], q.x, q.z);
                        split_pair_h2(acc[i + 2], acc[i + 3], q.y, q.w);
                        *over |= !(fabsf(acc[i]) <= kH2Limit) || !(fabsf(acc[i + 1]) <= kH2Limit) || !(fabsf(acc[i + 2]) <= kH2Limit) ||
                                 !(fabsf(acc[i + 3]) <= kH2Limit);
                        *reinterpret_cast<uint4 *>(oh + base + i) = q;
                    }
                }
            } else {                       // FC == 2: the lane one halving step away holds the other pair of the group
                uint32_t hi, lo;
                split_pair_h2(acc[0], acc[1], hi, lo);
                const uint32_t phi = __shfl_xor_sync(0xffffffffu, hi, 1 << R::AS), plo = __shfl_xor_sync(0xffffffffu, lo, 1 << R::AS);
                if (store) {
                    *over |= !(fabsf(acc[0]) <= kH2Limit) || !(fabsf(acc[1]) <= kH2Limit);
                    if ((base & 2) == 0) *reinterpret_cast<uint4 *>(oh + base) = make_uint4(hi, phi, lo, plo);
                }
            }
        }
    }
    if (!store) return;
    if constexpr (R::FC % 4 == 0) {
        if (io_aligned) {
#pragma unroll
            for (int i = 0; i < R::FC; i += 4)
                *reinterpret_cast<float4 *>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < R::FC; ++i) o[i] = acc[i];
}

constexpr int kRowLaneThreads = 256;

// ---- k=3 stride-1 (KVOL = 27) and k=1 (KVOL = 1, identity map) --------------------------
template <int CIN, int COUT, int KVOL>
__global__ void __launch_bounds__(kRowLaneThreads)
conv_rowlane_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                    const float *__restrict__ weight, const float *__restrict__ bias,
                    const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags,
                    int aligned_bits, uint32_t *__restrict__ out_h2, int out_h2_ld, int *__restrict__ overflow) {
    using R = RowLane<CIN, COUT>;
    extern __shared__ __align__(16) float ws[];
    bool over = false;
    load_weights_rowlane<CIN, COUT>(weight, KVOL, ws);
    __syncthreads();
    const bool in_aligned = aligned_bits & 1, io_aligned = aligned_bits & 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % R::LPR, g = lane / R::LPR;
    constexpr int ROWS_PER_BLOCK = (kRowLaneThreads / 32) * R::RPW;
    const int64_t n_tiles = (n + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS_PER_BLOCK + warp * R::RPW + g;
        const bool valid = row < n;
        float acc[COUT];
#pragma unroll
        for (int i = 0; i < COUT; ++i) acc[i] = 0.f;
        if constexpr (KVOL == 1) {
            const InVec<R::VEC> v = load_in<R::VEC>(in + row * in_ld + R::VEC * l, valid, in_aligned);
            accum_rowlane<CIN, COUT>(acc, v, ws, 0, l);
        } else {
            constexpr int B = 9;                               // offsets per load batch (memory-level parallelism)
#pragma unroll 1
            for (int k0 = 0; k0 < KVOL; k0 += B) {
                int32_t idx[B];
#pragma unroll
                for (int j = 0; j < B; ++j) idx[j] = valid ? __ldg(nbr + (int64_t)(k0 + j) * n + row) : -1;
                InVec<R::VEC> v[B];
#pragma unroll
                for (int j = 0; j < B; ++j)
                    v[j] = load_in<R::VEC>(in + (int64_t)idx[j] * in_ld + R::VEC * l, idx[j] >= 0, in_aligned);
#pragma unroll
                for (int j = 0; j < B; ++j)
                    if (__any_sync(0xffffffffu, idx[j] >= 0)) accum_rowlane<CIN, COUT>(acc, v[j], ws, k0 + j, l);
            }
        }
        finish_row<CIN, COUT>(acc, l, valid, row, bias, residual, res_ld, out, out_ld, flags, io_aligned, out_h2, out_h2_ld, &over);
    }
    if (over && overflow) *overflow = 1;
}

// ---- k=2 stride-2 down: one output row per parent, <= 8 children from the CSR map ---------
template <int CIN, int COUT>
__global__ void __launch_bounds__(kRowLaneThreads)
conv_down_rowlane_kernel(const float *__restrict__ in, int in_ld, const uint64_t *__restrict__ in_keys,
                         const int32_t *__restrict__ child_rows, const int32_t *__restrict__ child_off,
                         int64_t n_parents, const float *__restrict__ weight, const float *__restrict__ bias,
                         float *__restrict__ out, int out_ld, int flags, int aligned_bits,
                         uint32_t *__restrict__ out_h2, int out_h2_ld, int *__restrict__ overflow) {
    using R = RowLane<CIN, COUT>;
    extern __shared__ __align__(16) float ws[];
    bool over = false;
    load_weights_rowlane<CIN, COUT>(weight, 8, ws);
    __syncthreads();
    const bool in_aligned = aligned_bits & 1, io_aligned = aligned_bits & 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % R::LPR, g = lane / R::LPR;
    constexpr int ROWS_PER_BLOCK = (kRowLaneThreads / 32) * R::RPW;
    const int64_t n_tiles = (n_parents + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS_PER_BLOCK + warp * R::RPW + g;
        const bool valid = row < n_parents;
        const int beg = valid ? __ldg(child_off + row) : 0;
        const int cnt = valid ? __ldg(child_off + row + 1) - beg : 0;
        float acc[COUT];
#pragma unroll
        for (int i = 0; i < COUT; ++i) acc[i] = 0.f;
        int32_t src[8];
        int kk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            src[j] = j < cnt ? __ldg(child_rows + beg + j) : -1;
            kk[j] = src[j] >= 0 ? (int)(__ldg(in_keys + src[j]) & 7) : 0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (__any_sync(0xffffffffu, src[j] >= 0)) {
                const InVec<R::VEC> v = load_in<R::VEC>(in + (int64_t)src[j] * in_ld + R::VEC * l, src[j] >= 0, in_aligned);
                accum_rowlane<CIN, COUT>(acc, v, ws, kk[j], l);
            }
        }
        finish_row<CIN, COUT>(acc, l, valid, row, bias, nullptr, 0, out, out_ld, flags, io_aligned, out_h2, out_h2_ld, &over);
    }
    if (over && overflow) *overflow = 1;
}

// ---- generative k=2 stride-2 up: input row i -> output rows 8i .. 8i+7 ----------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(kRowLaneThreads)
conv_up_rowlane_kernel(const float *__restrict__ in, int in_ld, int64_t n_in, const float *__restrict__ weight,
                       const float *__restrict__ bias, float *__restrict__ out, int out_ld, int flags,
                       int aligned_bits, uint32_t *__restrict__ out_h2, int out_h2_ld, int *__restrict__ overflow) {
    using R = RowLane<CIN, COUT>;
    extern __shared__ __align__(16) float ws[];
    bool over = false;
    load_weights_rowlane<CIN, COUT>(weight, 8, ws);
    __syncthreads();
    const bool in_aligned = aligned_bits & 1, io_aligned = aligned_bits & 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % R::LPR, g = lane / R::LPR;
    constexpr int ROWS_PER_BLOCK = (kRowLaneThreads / 32) * R::RPW;
    const int64_t n_tiles = (n_in + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS_PER_BLOCK + warp * R::RPW + g;
        const bool valid = row < n_in;
        const InVec<R::VEC> v = load_in<R::VEC>(in + row * in_ld + R::VEC * l, valid, in_aligned);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float acc[COUT];
#pragma unroll
            for (int i = 0; i < COUT; ++i) acc[i] = 0.f;
            accum_rowlane<CIN, COUT>(acc, v, ws, k, l);
            finish_row<CIN, COUT>(acc, l, valid, row * 8 + k, bias, nullptr, 0, out, out_ld, flags, io_aligned, out_h2, out_h2_ld, &over);
        }
    }
    if (over && overflow) *overflow = 1;
}

}  // namespace pcgc
