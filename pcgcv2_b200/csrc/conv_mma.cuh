// conv_mma.cuh -- k=3 sparse convolution on the tensor cores: gathered input rows go straight
// from global memory into mma.sync A fragments (no shared-memory staging of the irregular
// operand), weights sit in shared memory pre-packed in B-fragment order, and every product is
// computed as 3xTF32 (hi*hi + hi*lo + lo*hi with hi = top 19 bits, lo = remainder) so the
// result keeps FP32 accuracy (the codec's 1e-4 activation tolerance and the round() at the
// bottleneck rule out single-pass TF32/BF16).
//
// Work split: a warp owns 16 consecutive output rows (one m16 tile) and all output channels;
// a CTA (8 warps) owns 128 rows; the grid is persistent over row tiles.  Per kernel offset a
// lane loads, for its two rows g and g+8, one 16-byte piece of each 16-channel chunk of the
// neighbour row: the contraction index is permuted (lane t holds physical channels 4t..4t+3 of
// a chunk, k-step s uses channels 4t+2s and 4t+2s+1) so that these natural float4 loads ARE the
// A fragments; the packed weights use the same permutation.  Offsets are processed in batches:
// all gathers of a batch are issued before its math (memory-level parallelism), and in STAGED
// mode the batch's weights arrive through a cp.async double buffer.  Output-stationary: each
// output row is written once, fused with bias / residual / ReLU.  No atomics.
#pragma once
#include "common.cuh"
#include "conv_tile.cuh"   // cp_async helpers

namespace pcgc {

template <int CIN, int COUT>
struct MmaCfg {
    static_assert(CIN == 8 || CIN % 16 == 0, "mma kernel: CIN must be 8 or a multiple of 16");
    static constexpr int KS = CIN / 8;                       // k-steps (of 8 channels) per offset
    static constexpr int NT = (COUT + 7) / 8;                // n-tiles (of 8 output channels)
    static constexpr int CHUNKS = CIN >= 16 ? CIN / 16 : 1;  // 16-byte loads per row per lane
    static constexpr int AV = CIN >= 16 ? 4 : 2;             // floats per load
    static constexpr int W_OFF = KS * NT * 64;               // packed floats per kernel offset
    static constexpr int THREADS = 256;
    static constexpr int ROWS_PER_CTA = (THREADS / 32) * 16;
    static constexpr bool RESIDENT = (size_t)27 * W_OFF * 4 <= 64 * 1024;
    static constexpr int BATCH = CIN <= 16 ? 9 : 3;          // offsets per gather batch
    static constexpr size_t smem_bytes() {
        return RESIDENT ? (size_t)27 * W_OFF * 4 : (size_t)2 * BATCH * W_OFF * 4;
    }
    static constexpr size_t packed_floats() { return (size_t)27 * W_OFF; }
};

// physical input channel held in A-fragment column (t or t+4) of k-step ks
__host__ __device__ __forceinline__ int mma_phys_channel(int cin, int ks, int t, int hi_col) {
    return cin >= 16 ? 16 * (ks >> 1) + 4 * t + 2 * (ks & 1) + hi_col : 2 * t + hi_col;
}

// W [kvol][cin][cout] -> packed [kvol][KS][NT][32 lanes][2]: lane (g = lane/4, t = lane%4) of n-tile j
// holds b0 = W[phys(ks,t,0)][8j+g], b1 = W[phys(ks,t,1)][8j+g] (zero beyond cout).
static __global__ void pack_weights_mma_kernel(const float *__restrict__ w, int kvol, int cin, int cout, float *__restrict__ packed) {
    const int KS = cin / 8, NT = (cout + 7) / 8;
    const int64_t total = (int64_t)kvol * KS * NT * 64;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 1), lane = (int)((i >> 1) & 31);
        int64_t r = i >> 6;
        const int j = (int)(r % NT); r /= NT;
        const int ks = (int)(r % KS);
        const int k = (int)(r / KS);
        const int g = lane >> 2, t = lane & 3;
        const int ci = mma_phys_channel(cin, ks, t, e), co = 8 * j + g;
        packed[i] = co < cout ? w[((int64_t)k * cin + ci) * cout + co] : 0.f;
    }
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// hi = top 19 bits of x (exactly representable in TF32), lo = x - hi (exact in fp32; the tensor
// core drops its low 13 bits: relative error of the 3-term product sum ~2^-21)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv_k3_mma_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                   const float *__restrict__ packed, const float *__restrict__ bias,
                   const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using C = MmaCfg<CIN, COUT>;
    constexpr int KS = C::KS, NT = C::NT, CHUNKS = C::CHUNKS, AV = C::AV, W_OFF = C::W_OFF, B = C::BATCH;
    extern __shared__ __align__(16) float wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

    if constexpr (C::RESIDENT) {                       // all 27 offsets stay in shared memory
        for (int i = threadIdx.x; i < 27 * W_OFF / 4; i += C::THREADS)
            cp_async16(wsm + 4 * i, packed + 4 * i, true);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    auto stage_weights = [&](int batch, int buf) {     // STAGED: one batch of offsets -> smem buffer
        const float *src = packed + (size_t)batch * B * W_OFF;
        float *dst = wsm + (size_t)buf * B * W_OFF;
        for (int i = threadIdx.x; i < B * W_OFF / 4; i += C::THREADS) cp_async16(dst + 4 * i, src + 4 * i, true);
        cp_async_commit();
    };

    const int64_t n_tiles = (n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row_a = tile * C::ROWS_PER_CTA + warp * 16 + g, row_b = row_a + 8;
        const bool va = row_a < n, vb = row_b < n;
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
        if constexpr (!C::RESIDENT) {
            __syncthreads();                            // previous tile's readers are done with both buffers
            stage_weights(0, 0);
        }
#pragma unroll 1
        for (int batch = 0; batch < 27 / B; ++batch) {
            // ---- gather the A fragments of this batch (issued before any math of the batch)
            int32_t ia[B], ib[B];
#pragma unroll
            for (int o = 0; o < B; ++o) {
                const int64_t k = batch * B + o;
                ia[o] = va ? __ldg(nbr + k * n + row_a) : -1;
                ib[o] = vb ? __ldg(nbr + k * n + row_b) : -1;
            }
            float fa[B][CHUNKS][AV], fb[B][CHUNKS][AV];
#pragma unroll
            for (int o = 0; o < B; ++o) {
#pragma unroll
                for (int q = 0; q < CHUNKS; ++q) {
                    if constexpr (AV == 4) {
                        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 x = ia[o] >= 0 ? __ldg(reinterpret_cast<const float4 *>(in + (int64_t)ia[o] * in_ld + 16 * q + 4 * t)) : z;
                        const float4 y = ib[o] >= 0 ? __ldg(reinterpret_cast<const float4 *>(in + (int64_t)ib[o] * in_ld + 16 * q + 4 * t)) : z;
                        fa[o][q][0] = x.x; fa[o][q][1] = x.y; fa[o][q][2] = x.z; fa[o][q][3] = x.w;
                        fb[o][q][0] = y.x; fb[o][q][1] = y.y; fb[o][q][2] = y.z; fb[o][q][3] = y.w;
                    } else {
                        const float2 z = make_float2(0.f, 0.f);
                        const float2 x = ia[o] >= 0 ? __ldg(reinterpret_cast<const float2 *>(in + (int64_t)ia[o] * in_ld + 2 * t)) : z;
                        const float2 y = ib[o] >= 0 ? __ldg(reinterpret_cast<const float2 *>(in + (int64_t)ib[o] * in_ld + 2 * t)) : z;
                        fa[o][q][0] = x.x; fa[o][q][1] = x.y;
                        fb[o][q][0] = y.x; fb[o][q][1] = y.y;
                    }
                }
            }
            const float *wb = wsm + (size_t)batch * B * W_OFF;
            if constexpr (!C::RESIDENT) {
                cp_async_wait<0>();                     // this batch's weights have landed
                __syncthreads();                        // ... for every thread; everyone left the other buffer
                if (batch + 1 < 27 / B) stage_weights(batch + 1, (batch + 1) & 1);
                wb = wsm + (size_t)(batch & 1) * B * W_OFF;
            }
            // ---- math
#pragma unroll
            for (int o = 0; o < B; ++o) {
                if (!__any_sync(0xffffffffu, (ia[o] >= 0) | (ib[o] >= 0))) continue;   // no row of the tile has this neighbour
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int q = CIN >= 16 ? ks >> 1 : 0, s = CIN >= 16 ? ks & 1 : 0;
                    uint32_t ah[4], al[4];
                    split_tf32(fa[o][q][2 * s], ah[0], al[0]);          // (row g,   col t)
                    split_tf32(fb[o][q][2 * s], ah[1], al[1]);          // (row g+8, col t)
                    split_tf32(fa[o][q][2 * s + 1], ah[2], al[2]);      // (row g,   col t+4)
                    split_tf32(fb[o][q][2 * s + 1], ah[3], al[3]);      // (row g+8, col t+4)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const float2 w = *reinterpret_cast<const float2 *>(wb + ((o * KS + ks) * NT + j) * 64 + 2 * lane);
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(w.x, bh0, bl0);
                        split_tf32(w.y, bh1, bl1);
                        // The tensor core adds into its accumulator with truncation (a biased ~2^-23 relative
                        // error per MMA).  Chaining all 27*KS*3 MMAs of a row into one accumulator let that bias
                        // grow to ~2e-5 per layer (measured), so each 3-MMA group sums into a zeroed temporary
                        // and joins the running sum through a round-to-nearest FADD.
                        float part[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_tf32(part, al, bh0, bh1);
                        mma_tf32(part, ah, bl0, bl1);
                        mma_tf32(part, ah, bh0, bh1);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[j][e] += part[e];
                    }
                }
            }
        }
        // ---- epilogue: (row g: acc[j][0..1]), (row g+8: acc[j][2..3]) at columns 8j + 2t, 8j + 2t + 1
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c0 = 8 * j + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t row = h ? row_b : row_a;
                if (!(h ? vb : va)) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = c0 + e;
                    if (c >= COUT) continue;
                    float v = acc[j][2 * h + e];
                    if (bias) v += __ldg(bias + c);
                    if (residual) v += __ldg(residual + row * res_ld + c);
                    if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
                    acc[j][2 * h + e] = v;
                }
                float *o = out + row * out_ld + c0;
                if (COUT % 2 == 0 && c0 + 1 < COUT && (out_ld & 1) == 0 && ((uintptr_t)out & 7) == 0)
                    *reinterpret_cast<float2 *>(o) = make_float2(acc[j][2 * h], acc[j][2 * h + 1]);
                else {
                    if (c0 < COUT) o[0] = acc[j][2 * h];
                    if (c0 + 1 < COUT) o[1] = acc[j][2 * h + 1];
                }
            }
        }
    }
}

}  // namespace pcgc
