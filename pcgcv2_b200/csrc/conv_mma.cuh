// conv_mma.cuh -- k=3 sparse convolution on the tensor cores (mma.sync m16n8k8 TF32, 3xTF32).
//
// The product is computed TRANSPOSED: D^T[cout, row] += W_k^T[cout, cin] * X_k^T[cin, row].  The
// weights are the 16 x 8 "A" operand (pre-split into hi/lo TF32 parts and pre-packed in fragment
// order in shared memory: one LDS.128 per part), and the gathered neighbour rows are the 8 x 8 "B"
// operand: lane (g = lane/4, t = lane%4) needs, for row g of an 8-row group, the channel pair
// (k = t, k = t+4).  The contraction index is permuted so that this pair is two ADJACENT floats of the
// row (k-step ks of a 16-channel chunk uses physical channels 4t+2(ks&1) and 4t+2(ks&1)+1): a lane's
// natural 16-byte load of the neighbour row IS its B fragments for two k-steps -- the irregular
// operand goes from global memory straight into MMA registers, with no shared-memory staging, no
// register shuffling and fully used 64-byte row segments.
//
// FP32 accuracy (the codec needs 1e-4 on activations and has a round() at the bottleneck): every
// product is 3xTF32, x = hi + lo with hi = top 19 bits (the tensor core ignores the low 13 bits of a
// TF32 operand, so hi costs nothing) and lo = x - hi.  The two small terms (W_hi*X_lo, W_lo*X_hi) are
// chained in a tensor-core accumulator; the main term W_hi*X_hi of one kernel offset is computed with
// a zero accumulator and joined to the running sum by a round-to-nearest FADD, because the tensor
// core adds with truncation and chaining all 27*CIN/8 MMAs biased the result by ~2e-5 (measured).
//
// A warp owns RG groups of 8 consecutive output rows and all output channels; a CTA 8 warps; the
// grid is persistent over row tiles.  Offsets are processed in batches (all gathers of a batch are
// issued before its math); big layers stream their weights per batch through a cp.async double
// buffer.  Output-stationary: every output row is written once, fused with bias/residual/ReLU.
#pragma once
#include "common.cuh"
#include "conv_tile.cuh"   // cp_async helpers

namespace pcgc {

// tuning defaults: RG = 8-row groups per warp, BATCH = offsets gathered ahead of the math
template <int CIN, int COUT>
struct MmaTune {
    static constexpr int MT = (COUT + 15) / 16;
    // measured on B200 with tools/run_variants.sh on the 1.69 M-row decoder level (profiles/r01_conv_variants.txt)
    static constexpr int RG = (MT >= 2 || CIN >= 64) ? 2 : 4;
    static constexpr int BATCH = (CIN <= 16) ? 3 : 1;
    static constexpr int MINB = CIN <= 32 ? 2 : 1;                    // min CTAs/SM handed to the register allocator
    // weights resident in smem when all 27 offsets (hi+lo) fit with 2 CTAs/SM; else streamed per batch
    static constexpr int RES = ((size_t)27 * (CIN / 8) * MT * 256 * 4 <= 112 * 1024) && !(CIN == 16 && MT == 2) ? 1 : 0;
};

template <int CIN, int COUT, int RG_ = MmaTune<CIN, COUT>::RG, int BATCH_ = MmaTune<CIN, COUT>::BATCH,
          int RES_ = MmaTune<CIN, COUT>::RES>
struct MmaCfg {
    static_assert(CIN == 8 || CIN % 16 == 0, "mma kernel: CIN must be 8 or a multiple of 16");
    static_assert(27 % BATCH_ == 0, "BATCH must divide 27");
    static constexpr int KS = CIN / 8;                        // k-steps (8 channels) per offset
    static constexpr int MT = (COUT + 15) / 16;               // m-tiles (16 output channels)
    static constexpr int CHUNKS = CIN >= 16 ? CIN / 16 : 1;   // loads per gathered row per lane
    static constexpr int AV = CIN >= 16 ? 4 : 2;              // floats per load
    static constexpr int RG = RG_;
    static constexpr int BATCH = BATCH_;
    static constexpr int W_OFF = KS * MT * 256;               // packed floats per offset (hi + lo quads)
    static constexpr int THREADS = 256;
    static constexpr int ROWS_PER_WARP = 8 * RG;
    static constexpr int ROWS_PER_CTA = (THREADS / 32) * ROWS_PER_WARP;
    static constexpr bool RESIDENT = RES_ != 0;
    // tile's kernel-map slice staged in smem unless big resident weights leave no room for 2 CTAs/SM
    static constexpr bool IDX_SMEM = !(RESIDENT && (size_t)27 * W_OFF * 4 > 64 * 1024);
    static constexpr size_t weight_smem_floats() { return RESIDENT ? (size_t)27 * W_OFF : (size_t)2 * BATCH * W_OFF; }
    // + the tile's slice of the kernel map, [27][ROWS_PER_WARP] int32 per warp (loaded coalesced once per tile)
    static constexpr size_t smem_bytes() {
        return weight_smem_floats() * 4 + (IDX_SMEM ? (size_t)(THREADS / 32) * 27 * ROWS_PER_WARP * 4 : 0);
    }
    static constexpr size_t packed_floats() { return (size_t)27 * W_OFF; }
};

// physical input channel in fragment column (t or t+4) of k-step ks
__host__ __device__ __forceinline__ int mma_phys_channel(int cin, int ks, int t, int hi_col) {
    return cin >= 16 ? 16 * (ks >> 1) + 4 * t + 2 * (ks & 1) + hi_col : 2 * t + hi_col;
}

__host__ __device__ __forceinline__ void split_tf32_f(float x, float &hi, float &lo) {
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u &= 0xFFFFE000u;
    hi = v.f;
    lo = x - hi;
}

// W [27][cin][cout] -> packed [27][KS][MT][32 lanes][8] = {a0..a3 hi, a0..a3 lo} with
// a0 = W[phys(ks,t,0)][16m+g], a1 = W[phys(ks,t,0)][16m+g+8], a2 = W[phys(ks,t,1)][16m+g], a3 = W[phys(ks,t,1)][16m+g+8]
static __global__ void pack_weights_mma_kernel(const float *__restrict__ w, int kvol, int cin, int cout, float *__restrict__ packed) {
    const int KS = cin / 8, MT = (cout + 15) / 16;
    const int64_t total = (int64_t)kvol * KS * MT * 32 * 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 3), lane = (int)((i >> 2) & 31);
        int64_t r = i >> 7;
        const int m = (int)(r % MT); r /= MT;
        const int ks = (int)(r % KS);
        const int k = (int)(r / KS);
        const int g = lane >> 2, t = lane & 3;
        const int ci = mma_phys_channel(cin, ks, t, e >> 1), co = 16 * m + g + 8 * (e & 1);
        const float x = co < cout ? w[((int64_t)k * cin + ci) * cout + co] : 0.f;
        float hi, lo;
        split_tf32_f(x, hi, lo);
        float *dst = packed + ((((int64_t)k * KS + ks) * MT + m) * 32 + lane) * 8;
        dst[e] = hi;
        dst[4 + e] = lo;
    }
}

// D = A*B + C, A = 16x8 (weights), B = 8x8 (features), all fragments in registers
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float4 &a, float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const float4 &a, float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};\n"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int CIN, int COUT, int RG_ = MmaTune<CIN, COUT>::RG, int BATCH_ = MmaTune<CIN, COUT>::BATCH,
          int MINB = MmaTune<CIN, COUT>::MINB, int RES_ = MmaTune<CIN, COUT>::RES>
__global__ void __launch_bounds__(256, MINB)
conv_k3_mma_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                   const float *__restrict__ packed, const float *__restrict__ bias,
                   const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using C = MmaCfg<CIN, COUT, RG_, BATCH_, RES_>;
    constexpr int KS = C::KS, MT = C::MT, CHUNKS = C::CHUNKS, AV = C::AV, RG = C::RG, W_OFF = C::W_OFF, B = C::BATCH;
    extern __shared__ __align__(16) float wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    int32_t *idx_s = reinterpret_cast<int32_t *>(wsm + C::weight_smem_floats()) + warp * 27 * C::ROWS_PER_WARP;

    if constexpr (C::RESIDENT) {                        // all 27 offsets stay in shared memory
        for (int i = threadIdx.x; i < 27 * W_OFF / 4; i += C::THREADS) cp_async16(wsm + 4 * i, packed + 4 * i, true);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    auto stage_weights = [&](int batch, int buf) {      // STAGED: one batch of offsets -> smem buffer
        const float *src = packed + (size_t)batch * B * W_OFF;
        float *dst = wsm + (size_t)buf * B * W_OFF;
        for (int i = threadIdx.x; i < B * W_OFF / 4; i += C::THREADS) cp_async16(dst + 4 * i, src + 4 * i, true);
        cp_async_commit();
    };

    const int64_t n_tiles = (n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * C::ROWS_PER_CTA + warp * C::ROWS_PER_WARP;     // first row of this warp
        float acc[MT][RG][4], small[MT][RG][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][r][e] = small[m][r][e] = 0.f;
        if constexpr (!C::RESIDENT) {
            __syncthreads();                             // previous tile's readers are done with both buffers
            stage_weights(0, 0);
        }
        // this warp's slice of the kernel map: 27 coalesced loads, no dependent index load per offset later
        if constexpr (C::IDX_SMEM) {
            __syncwarp();
            for (int i = lane; i < 27 * C::ROWS_PER_WARP; i += 32) {
                const int k = i / C::ROWS_PER_WARP, rr = i % C::ROWS_PER_WARP;
                idx_s[i] = row0 + rr < n ? __ldg(nbr + (int64_t)k * n + row0 + rr) : -1;
            }
            __syncwarp();
        }
        const char *in_lane = reinterpret_cast<const char *>(in + (AV == 4 ? 4 * t : 2 * t));
        const int64_t ld_bytes = (int64_t)in_ld * 4;
#pragma unroll 1
        for (int batch = 0; batch < 27 / B; ++batch) {
            // ---- gather: lane loads its piece of the neighbour row of output row 8*r + g, for every group r
            int32_t idx[B][RG];
#pragma unroll
            for (int o = 0; o < B; ++o)
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    if constexpr (C::IDX_SMEM) {
                        idx[o][r] = idx_s[(batch * B + o) * C::ROWS_PER_WARP + 8 * r + g];
                    } else {
                        const int64_t row = row0 + 8 * r + g;
                        idx[o][r] = row < n ? __ldg(nbr + (int64_t)(batch * B + o) * n + row) : -1;
                    }
                }
            float x[B][RG][CHUNKS][AV];
#pragma unroll
            for (int o = 0; o < B; ++o)
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int q = 0; q < CHUNKS; ++q) {
                        const char *src = in_lane + idx[o][r] * ld_bytes + 64 * q;
                        const bool ok = idx[o][r] >= 0;
                        if constexpr (AV == 4) {
                            const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            x[o][r][q][0] = v.x; x[o][r][q][1] = v.y; x[o][r][q][2] = v.z; x[o][r][q][3] = v.w;
                        } else {
                            const float2 v = ok ? __ldg(reinterpret_cast<const float2 *>(src)) : make_float2(0.f, 0.f);
                            x[o][r][q][0] = v.x; x[o][r][q][1] = v.y;
                        }
                    }
            const float *wb = wsm + (size_t)batch * B * W_OFF;
            if constexpr (!C::RESIDENT) {
                cp_async_wait<0>();                      // this batch's weights have landed ...
                __syncthreads();                         // ... for every thread; everyone left the other buffer
                if (batch + 1 < 27 / B) stage_weights(batch + 1, (batch + 1) & 1);
                wb = wsm + (size_t)(batch & 1) * B * W_OFF;
            }
            // ---- math
#pragma unroll
            for (int o = 0; o < B; ++o) {
                bool any_here = false;
#pragma unroll
                for (int r = 0; r < RG; ++r) any_here |= idx[o][r] >= 0;
                if (!__any_sync(0xffffffffu, any_here)) continue;            // no row of the warp has this neighbour
                float part[MT][RG][4];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int q = CIN >= 16 ? ks >> 1 : 0, s = CIN >= 16 ? 2 * (ks & 1) : 0;
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const float4 *wp = reinterpret_cast<const float4 *>(wb + (((o * KS + ks) * MT + m) * 32 + lane) * 8);
                        const float4 wh = wp[0], wl = wp[1];
#pragma unroll
                        for (int r = 0; r < RG; ++r) {
                            const float b0 = x[o][r][q][s], b1 = x[o][r][q][s + 1];
                            const float h0 = tf32_hi(b0), h1 = tf32_hi(b1);
                            mma_tf32(small[m][r], wh, b0 - h0, b1 - h1);     // W_hi * X_lo
                            mma_tf32(small[m][r], wl, h0, h1);               // W_lo * X_hi
                            if (ks == 0) mma_tf32_zero(part[m][r], wh, h0, h1);   // W_hi * X_hi
                            else mma_tf32(part[m][r], wh, h0, h1);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int r = 0; r < RG; ++r)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[m][r][e] += part[m][r][e];
            }
        }
        // ---- epilogue: fragment (m, r): e=0 -> (cout 16m+g, row 8r+2t), e=1 -> (16m+g, 8r+2t+1),
        //                               e=2 -> (16m+g+8, 8r+2t), e=3 -> (16m+g+8, 8r+2t+1)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = 16 * m + g + 8 * h;
                if (co >= COUT) continue;
                const float bv = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int64_t row = row0 + 8 * r + 2 * t + e;
                        if (row >= n) continue;
                        float v = acc[m][r][2 * h + e] + small[m][r][2 * h + e] + bv;
                        if (residual) v += __ldg(residual + row * res_ld + co);
                        if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
                        out[row * out_ld + co] = v;
                    }
            }
        }
    }
}

}  // namespace pcgc
