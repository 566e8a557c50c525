// conv_h2.cuh -- k=3 sparse convolution on the tensor cores over PRE-SPLIT half-precision features
// (round-1 v7: mma.sync.m16n8k16 f16 with f32 accumulation, three products per term pair).
//
// Why: the 3xTF32 kernels (conv_mma.cuh / conv_pipe.cuh / conv_octet.cuh) split every gathered value into
// hi + lo inside the hot loop -- each row is split ~19 times, once per neighbour that gathers it -- and
// issue 6 HMMA.1688.TF32 per 16 input channels.  Measured on B200 (tools/mma_rate.cu): HMMA.16816.F16
// issues at the same 8 cycles per sub-partition as HMMA.1688.TF32, i.e. twice the work per issue.  So:
//
//  * features travel between layers in the "h2" format: x = hi + lo with hi = f16(x), lo = f16(x - hi),
//    22 significand bits, 4 bytes per value like the fp32 they replace.  The PRODUCER splits each value
//    once in its epilogue (or split_h2_kernel does, for producers that only write fp32).  Four channels
//    form one 16-byte group {hi0 hi1 | hi2 hi3 | lo0 lo1 | lo2 lo3}: a lane's natural 16-byte load of
//    a gathered row IS its MMA fragments (the contraction index is permuted, as in conv_mma.cuh), so the
//    hot loop is LDG.128 -> 3 x HMMA.16816 with no ALU work at all;
//  * weights are pre-split the same way after scaling by a power of two (so that the lo parts stay
//    normal f16 numbers); the epilogue multiplies by the inverse scale (exact);
//  * W_hi*X_lo and W_lo*X_hi are chained in one tensor-core accumulator, W_hi*X_hi of one kernel offset
//    is accumulated from zero and joined to the running sum by a round-to-nearest FADD (the tensor core
//    adds with truncation), exactly like the 3xTF32 kernels.
//
// Accuracy (CPU emulation over the whole network, tools/h2_emulation.py, vox8 known-answer cloud): max|delta|/max|ref|
// <= 1.7e-6 on every layer, bit-identical bitstream and decoded occupancy.  Range: |x| must stay below
// 65504 (activations of the trained networks peak at ~30); the epilogue raises *overflow when a value
// leaves the range so that the caller can re-run on the 3xTF32 path instead of returning a wrong result.
//
// Two formulations, as in conv_pipe.cuh: T (COUT % 16 == 0): D^T[cout,row] += W^T X^T, weights are the
// 16x16 A operand, 8 gathered rows the B operand (register pairs straight out of the LDG.128); NT (COUT <= 8):
// D[row,cout] += X W, 16 gathered rows the A operand, weights the 16x8 B operand.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_tile.cuh"   // cp_async helpers

namespace pcgc {

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }

// (a, b) -> packed hi pair, packed lo pair (a in the low half)
__device__ __forceinline__ void split_pair_h2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(a - hf.x, b - hf.y));
}
__device__ __forceinline__ float2 join_pair_h2(uint32_t hi, uint32_t lo) {
    const float2 a = __half22float2(*reinterpret_cast<__half2 *>(&hi)), b = __half22float2(*reinterpret_cast<__half2 *>(&lo));
    return make_float2(a.x + b.x, a.y + b.y);
}

constexpr float kH2Limit = 60000.f;     // |x| above this raises the overflow flag (f16 max = 65504)

// fp32 [n, c] (row stride in_ld) -> h2 [n, c] (row stride out_ld, in 4-byte units); c % 4 == 0
static __global__ void split_h2_kernel(const float *__restrict__ in, int in_ld, int64_t n, int c, uint32_t *__restrict__ out,
                                       int out_ld, int *__restrict__ overflow, int vec_ok) {
    const int groups = c >> 2;
    const int64_t total = n * groups;
    bool over = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / groups;
        const int u = (int)(i - row * groups);
        const float *src = in + row * in_ld + 4 * u;
        float4 v;
        if (vec_ok) v = __ldg(reinterpret_cast<const float4 *>(src));
        else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        over |= !(fabsf(v.x) <= kH2Limit) || !(fabsf(v.y) <= kH2Limit) || !(fabsf(v.z) <= kH2Limit) || !(fabsf(v.w) <= kH2Limit);
        uint4 o;
        split_pair_h2(v.x, v.y, o.x, o.z);
        split_pair_h2(v.z, v.w, o.y, o.w);
        uint32_t *dst = out + row * out_ld + 4 * u;
        if (vec_ok) *reinterpret_cast<uint4 *>(dst) = o;
        else { dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w; }
    }
    if (over && overflow) *overflow = 1;
}

// h2 [n, c] -> fp32 [n, c] (tests, and consumers that have no h2 path)
static __global__ void join_h2_kernel(const uint32_t *__restrict__ in, int in_ld, int64_t n, int c, float *__restrict__ out, int out_ld) {
    const int groups = c >> 2;
    const int64_t total = n * groups;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / groups;
        const int u = (int)(i - row * groups);
        const uint32_t *src = in + row * in_ld + 4 * u;
        const float2 a = join_pair_h2(src[0], src[2]), b = join_pair_h2(src[1], src[3]);
        float *dst = out + row * out_ld + 4 * u;
        dst[0] = a.x; dst[1] = a.y; dst[2] = b.x; dst[3] = b.y;
    }
}

template <int CIN, int COUT, bool NT_, int RG_, int D_, int WARPS_, int KV_ = 27>
struct H2Cfg {
    static constexpr int KV = KV_;                                    // kernel volume: 27 (k=3) or 8 (k=2 stride 2: child slots)
    static_assert(CIN % 16 == 0 || CIN == 8, "h2 kernel: CIN must be 8 or a multiple of 16");
    static_assert(NT_ || COUT % 16 == 0, "h2 kernel, T formulation: COUT must be a multiple of 16");
    static constexpr bool NT = NT_;
    static constexpr bool C8 = CIN == 8;                              // 8 channels: [hi | lo] of a row fill ONE k-step (see below)
    static constexpr int KS = C8 ? 1 : CIN / 16;                      // k-steps (16 contraction slots) per offset
    static constexpr int CT = NT ? (COUT + 7) / 8 : COUT / 16;        // output-channel tiles (N=8 / M=16)
    static constexpr int RG = RG_, D = D_, WARPS = WARPS_;
    static constexpr int GROUP_ROWS = NT ? 16 : 8;
    static constexpr int NR = NT ? 2 * RG : RG;                       // gathered rows per lane per offset
    static constexpr int RPW = GROUP_ROWS * RG;                       // output rows per warp
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int ROWS_PER_CTA = WARPS * RPW;
    static constexpr int W_OFF = KS * CT * (NT ? 128 : 256);          // packed 32-bit words per offset (hi + lo)
    static constexpr size_t packed_words() { return (size_t)KV * W_OFF; }
    static constexpr size_t smem_bytes() { return packed_words() * 4 + (size_t)WARPS * KV * RPW * 4; }
};

// physical input channel of logical pair slot (q, t, half h in 0..1, element e in 0..1): 16q + 4t + 2h + e
// NT packing: W [27][cin][cout] -> [27][KS][CT][32 lanes] x uint4 {b0 hi, b1 hi, b0 lo, b1 lo},
//   b0 = (W[16q+4t][8c+g], W[16q+4t+1][8c+g]), b1 = (W[16q+4t+2][8c+g], W[16q+4t+3][8c+g])
// T packing:  -> [27][KS][CT][2 (hi, lo)][32 lanes] x uint4 {a0, a1, a2, a3},
//   a0 = (W[16q+4t][16c+g], W[16q+4t+1][16c+g]), a1 = same channels, cout 16c+g+8, a2/a3 = channels 16q+4t+2, +3
static __global__ void pack_weights_h2_kernel(const float *__restrict__ w, int kvol, int cin, int cout, int nt, float scale,
                                              uint32_t *__restrict__ packed) {
    const int KS = cin / 16, CT = nt ? (cout + 7) / 8 : cout / 16;
    const int per_lane = nt ? 2 : 4;                                  // channel pairs per lane per (k, q, c)
    const int64_t total = (int64_t)kvol * KS * CT * 32 * per_lane;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % per_lane), lane = (int)((i / per_lane) & 31);
        int64_t r = i / (per_lane * 32);
        const int c = (int)(r % CT); r /= CT;
        const int q = (int)(r % KS);
        const int k = (int)(r / KS);
        const int g = lane >> 2, t = lane & 3;
        int ci, co;
        if (nt) { ci = 16 * q + 4 * t + 2 * e; co = 8 * c + g; }
        else { ci = 16 * q + 4 * t + 2 * (e >> 1); co = 16 * c + g + 8 * (e & 1); }
        const float x0 = co < cout ? w[((int64_t)k * cin + ci) * cout + co] * scale : 0.f;
        const float x1 = co < cout ? w[((int64_t)k * cin + ci + 1) * cout + co] * scale : 0.f;
        uint32_t hi, lo;
        split_pair_h2(x0, x1, hi, lo);
        if (nt) {
            uint32_t *dst = packed + ((((int64_t)k * KS + q) * CT + c) * 32 + lane) * 4;
            dst[e] = hi;
            dst[2 + e] = lo;
        } else {
            uint32_t *dst = packed + (((int64_t)k * KS + q) * CT + c) * 256 + lane * 4 + e;
            dst[0] = hi;
            dst[128] = lo;
        }
    }
}

// ---- CIN = 8 ------------------------------------------------------------------------------------------------------
// An h2 row of eight channels is 32 bytes = the words {hi01 hi23 lo01 lo23 | hi45 hi67 lo45 lo67}.  Lane (g, t) loads the
// 8 bytes at word 2t: t = 0: (hi01, hi23), t = 1: (lo01, lo23), t = 2: (hi45, hi67), t = 3: (lo45, lo67) -- its two
// m16n8k16 fragment registers (contraction slots 2t, 2t+1 and 2t+8, 2t+9).  The sixteen slots of ONE k-step so carry the hi
// AND the lo halves of all eight channels, and the split product needs two MMAs per offset instead of three:
//     MMA A: weights W_hi against every slot          -> W_hi x_hi + W_hi x_lo
//     MMA B: weights W_lo against the hi slots, 0 against the lo slots -> W_lo x_hi
// both chained from zero per kernel offset and joined to the running sum by a round-to-nearest FADD.
// slot pair (2t, 2t+1) holds channels P(t) = {0, 0, 4, 4}[t], pair (2t+8, 2t+9) channels Q(t) = {2, 2, 6, 6}[t]; odd t = lo halves.
// T packing:  [kvol][CT][2 (A, B)][32 lanes] x uint4 {a0, a1, a2, a3}: a0 = (W[P], W[P+1])[16c+g], a1 = same, cout 16c+g+8,
//             a2 / a3 = channels Q, Q+1
// NT packing: [kvol][CT][32 lanes] x uint4 {b0 A, b1 A, b0 B, b1 B}: b0 = (W[P], W[P+1])[8c+g], b1 = (W[Q], W[Q+1])[8c+g]
static __global__ void pack_weights_h2c8_kernel(const float *__restrict__ w, int kvol, int cout, int nt, float scale,
                                                uint32_t *__restrict__ packed) {
    const int CT = nt ? (cout + 7) / 8 : cout / 16;
    const int per_lane = nt ? 2 : 4;
    const int64_t total = (int64_t)kvol * CT * 32 * per_lane;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % per_lane), lane = (int)((i / per_lane) & 31);
        int64_t r = i / (per_lane * 32);
        const int c = (int)(r % CT);
        const int k = (int)(r / CT);
        const int g = lane >> 2, t = lane & 3;
        const int P = (t >> 1) * 4, Q = P + 2;
        int ci, co;
        if (nt) { ci = e ? Q : P; co = 8 * c + g; }
        else { ci = (e >> 1) ? Q : P; co = 16 * c + g + 8 * (e & 1); }
        const float x0 = co < cout ? w[((int64_t)k * 8 + ci) * cout + co] * scale : 0.f;
        const float x1 = co < cout ? w[((int64_t)k * 8 + ci + 1) * cout + co] * scale : 0.f;
        uint32_t hi, lo;
        split_pair_h2(x0, x1, hi, lo);
        if (t & 1) lo = 0u;                                           // MMA B: W_lo only against the hi slots
        if (nt) {
            uint32_t *dst = packed + (((int64_t)k * CT + c) * 32 + lane) * 4;
            dst[e] = hi;
            dst[2 + e] = lo;
        } else {
            uint32_t *dst = packed + ((int64_t)k * CT + c) * 256 + lane * 4 + e;
            dst[0] = hi;
            dst[128] = lo;
        }
    }
}

// D(16x8, f32) = A(16x16, f16) * B(16x8, f16) + C
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};\n"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// one output pair (row, channels co, co+1): scale, bias, residual, ReLU; fp32 and/or h2 store
struct H2Epilogue {
    const float *bias, *residual;
    float *out;
    uint32_t *out_h2;
    int res_ld, out_ld, out_h2_ld, flags;
    float inv_scale;
    bool over = false;
    __device__ __forceinline__ void store_pair(int64_t row, int co, float v0, float v1) {
        v0 = v0 * inv_scale + (bias ? __ldg(bias + co) : 0.f);
        v1 = v1 * inv_scale + (bias ? __ldg(bias + co + 1) : 0.f);
        if (residual) {
            const float2 rv = __ldg(reinterpret_cast<const float2 *>(residual + row * res_ld + co));
            v0 += rv.x; v1 += rv.y;
        }
        if (flags & PCGC_EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        if (out) *reinterpret_cast<float2 *>(out + row * out_ld + co) = make_float2(v0, v1);
        if (out_h2) {
            over |= !(fabsf(v0) <= kH2Limit) || !(fabsf(v1) <= kH2Limit);
            uint32_t hi, lo;
            split_pair_h2(v0, v1, hi, lo);
            uint32_t *dst = out_h2 + row * out_h2_ld + (co & ~3) + ((co >> 1) & 1);
            dst[0] = hi;
            dst[2] = lo;
        }
    }
    __device__ __forceinline__ void store_one(int64_t row, int co, float v) {     // odd COUT (classifier): fp32 only
        v = v * inv_scale + (bias ? __ldg(bias + co) : 0.f);
        if (residual) v += __ldg(residual + row * res_ld + co);
        if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
        out[row * out_ld + co] = v;
    }
};

template <int CIN, int COUT, bool NT, int RG, int D, int WARPS, int MINB, int KV = 27>
__global__ void __launch_bounds__(32 * WARPS, MINB)
conv_k3_h2_kernel(const uint32_t *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                  const uint32_t *__restrict__ packed, float inv_scale, const float *__restrict__ bias,
                  const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld,
                  uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    using C = H2Cfg<CIN, COUT, NT, RG, D, WARPS, KV>;
    constexpr int KS = C::KS, CT = C::CT, NR = C::NR, RPW = C::RPW, W_OFF = C::W_OFF;
    constexpr bool C8 = C::C8;
    constexpr bool IDXV = NR == 2 || NR == 4;                         // lane's NR kernel-map entries adjacent in smem
    extern __shared__ __align__(16) uint32_t wsm_h2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    int32_t *idx_s = reinterpret_cast<int32_t *>(wsm_h2 + C::packed_words()) + warp * KV * RPW;

    for (int i = threadIdx.x; i < KV * W_OFF / 4; i += C::THREADS) cp_async16(wsm_h2 + 4 * i, packed + 4 * i, true);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // gathered-row addressing: kernel-map entries are staged as row offsets in 16-byte units (in_ld % 4 == 0)
    const char *in_lane = reinterpret_cast<const char *>(in + (C8 ? 2 : 4) * t);
    const int32_t ld16 = in_ld >> 2;
    const int64_t n_tiles = (n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    H2Epilogue epi{bias, residual, out, out_h2, res_ld, out_ld, out_h2_ld, flags, inv_scale};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * C::ROWS_PER_CTA + warp * RPW;     // first row of this warp
        {   // L2 prefetch of this CTA's next tile: its own rows are the first-touch (DRAM latency) part of the gathers
            const int64_t nt = tile + gridDim.x;
            if (nt < n_tiles) {
                const int64_t r0 = nt * C::ROWS_PER_CTA;
                const int64_t rows = n - r0 < C::ROWS_PER_CTA ? n - r0 : C::ROWS_PER_CTA;
                const char *p = reinterpret_cast<const char *>(in + r0 * in_ld);
                const int64_t bytes = rows * in_ld * 4;
                for (int64_t off = (int64_t)threadIdx.x * 128; off < bytes; off += C::THREADS * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
                constexpr int LPS = (C::ROWS_PER_CTA * 4 + 127) / 128;
                for (int i = threadIdx.x; i < KV * LPS; i += C::THREADS) {
                    const int k = i / LPS, l = i % LPS;
                    if ((int64_t)l * 32 < rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(nbr + (int64_t)k * n + r0 + l * 32));
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < KV * RPW; i += 32) {                   // this warp's slice of the kernel map, coalesced
            const int k = i / RPW, rr = i % RPW;                      // local row rr = 8j + g (j-th gathered row of lane group g)
            const int32_t v = row0 + rr < n ? __ldg(nbr + (int64_t)k * n + row0 + rr) : -1;
            idx_s[IDXV ? k * RPW + (rr & 7) * NR + (rr >> 3) : i] = v >= 0 ? v * ld16 : -1;
        }
        __syncwarp();

        float acc[CT][RG][4], small[CT][RG][4];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[c][r][e] = small[c][r][e] = 0.f;

        uint4 x[D][NR][KS];
        int32_t idn[NR];                                              // kernel-map entries of the NEXT offset to gather
        auto load_idx = [&](int o) {
            if constexpr (NR == 4) {
                const int4 v = *reinterpret_cast<const int4 *>(idx_s + o * RPW + g * 4);
                idn[0] = v.x; idn[1] = v.y; idn[2] = v.z; idn[3] = v.w;
            } else if constexpr (NR == 2) {
                const int2 v = *reinterpret_cast<const int2 *>(idx_s + o * RPW + g * 2);
                idn[0] = v.x; idn[1] = v.y;
            } else {
#pragma unroll
                for (int j = 0; j < NR; ++j) idn[j] = idx_s[o * RPW + 8 * j + g];
            }
        };
        auto gather = [&](int o, int st) {
            int32_t id[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) id[j] = idn[j];
            if (o + 1 < KV) load_idx(o + 1);                          // one offset ahead: its LDS latency hides behind this offset's math
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const bool ok = id[j] >= 0;
#pragma unroll
                for (int q = 0; q < KS; ++q) {
                    const char *src = in_lane + (uint64_t)(uint32_t)id[j] * 16u + 64 * q;
                    if constexpr (C8) {                               // 8 bytes per lane: the row's two fragment registers
                        const uint2 v = ok ? __ldg(reinterpret_cast<const uint2 *>(src)) : make_uint2(0u, 0u);
                        x[st][j][q] = make_uint4(v.x, v.y, 0u, 0u);
                    } else {
                        x[st][j][q] = ok ? __ldg(reinterpret_cast<const uint4 *>(src)) : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
        };
        auto math = [&](int o, int st) {
            const uint32_t *wb = wsm_h2 + (size_t)o * W_OFF;
            float part[CT][RG][4];
#pragma unroll
            for (int q = 0; q < KS; ++q) {
                if constexpr (C8 && NT) {
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(wb + (c * 32 + lane) * 4);
#pragma unroll
                        for (int r = 0; r < RG; ++r)
                            mma_f16_zero(part[c][r], x[st][2 * r][0].x, x[st][2 * r + 1][0].x, x[st][2 * r][0].y, x[st][2 * r + 1][0].y, w.x, w.y);
#pragma unroll
                        for (int r = 0; r < RG; ++r)
                            mma_f16(part[c][r], x[st][2 * r][0].x, x[st][2 * r + 1][0].x, x[st][2 * r][0].y, x[st][2 * r + 1][0].y, w.z, w.w);
                    }
                } else if constexpr (C8) {
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const uint4 *wp = reinterpret_cast<const uint4 *>(wb + c * 256) + lane;
                        const uint4 wa = wp[0], wl = wp[32];
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16_zero(part[c][r], wa.x, wa.y, wa.z, wa.w, x[st][r][0].x, x[st][r][0].y);
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16(part[c][r], wl.x, wl.y, wl.z, wl.w, x[st][r][0].x, x[st][r][0].y);
                    }
                } else if constexpr (NT) {
                    uint4 w[CT];
#pragma unroll
                    for (int c = 0; c < CT; ++c) w[c] = *reinterpret_cast<const uint4 *>(wb + ((q * CT + c) * 32 + lane) * 4);
#pragma unroll
                    for (int c = 0; c < CT; ++c)
#pragma unroll
                        for (int r = 0; r < RG; ++r)                                                   // X_lo * W_hi
                            mma_f16(small[c][r], x[st][2 * r][q].z, x[st][2 * r + 1][q].z, x[st][2 * r][q].w, x[st][2 * r + 1][q].w, w[c].x, w[c].y);
#pragma unroll
                    for (int c = 0; c < CT; ++c)
#pragma unroll
                        for (int r = 0; r < RG; ++r) {                                                 // X_hi * W_hi
                            if (q == 0) mma_f16_zero(part[c][r], x[st][2 * r][q].x, x[st][2 * r + 1][q].x, x[st][2 * r][q].y, x[st][2 * r + 1][q].y, w[c].x, w[c].y);
                            else mma_f16(part[c][r], x[st][2 * r][q].x, x[st][2 * r + 1][q].x, x[st][2 * r][q].y, x[st][2 * r + 1][q].y, w[c].x, w[c].y);
                        }
#pragma unroll
                    for (int c = 0; c < CT; ++c)
#pragma unroll
                        for (int r = 0; r < RG; ++r)                                                   // X_hi * W_lo
                            mma_f16(small[c][r], x[st][2 * r][q].x, x[st][2 * r + 1][q].x, x[st][2 * r][q].y, x[st][2 * r + 1][q].y, w[c].z, w[c].w);
                } else {
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const uint4 *wp = reinterpret_cast<const uint4 *>(wb + (q * CT + c) * 256) + lane;
                        const uint4 wh = wp[0], wl = wp[32];
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16(small[c][r], wh.x, wh.y, wh.z, wh.w, x[st][r][q].z, x[st][r][q].w);   // W_hi * X_lo
#pragma unroll
                        for (int r = 0; r < RG; ++r) {                                                                           // W_hi * X_hi
                            if (q == 0) mma_f16_zero(part[c][r], wh.x, wh.y, wh.z, wh.w, x[st][r][q].x, x[st][r][q].y);
                            else mma_f16(part[c][r], wh.x, wh.y, wh.z, wh.w, x[st][r][q].x, x[st][r][q].y);
                        }
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16(small[c][r], wl.x, wl.y, wl.z, wl.w, x[st][r][q].x, x[st][r][q].y);   // W_lo * X_hi
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[c][r][e] += part[c][r][e];
        };

        load_idx(0);
#pragma unroll
        for (int o = 0; o < D - 1; ++o) gather(o, o);
#pragma unroll 1
        for (int ob = 0; ob < KV; ob += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int o = ob + d;
                if (o < KV) {
                    if (o + D - 1 < KV) gather(o + D - 1, (d + D - 1) % D);
                    math(o, d);
                }
            }
        }

        // ---- epilogue
        if constexpr (NT) {
            // fragment (c, r): e=0 -> (row 16r+g, cout 8c+2t), e=1 -> (.., 8c+2t+1), e=2/3 -> row 16r+g+8
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const int co = 8 * c + 2 * t;
                if (co >= COUT) continue;
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int64_t row = row0 + 16 * r + 8 * h + g;
                        if (row >= n) continue;
                        const float v0 = acc[c][r][2 * h] + small[c][r][2 * h], v1 = acc[c][r][2 * h + 1] + small[c][r][2 * h + 1];
                        if constexpr (COUT % 2 == 0) epi.store_pair(row, co, v0, v1);
                        else {
                            epi.store_one(row, co, v0);
                            if (co + 1 < COUT) epi.store_one(row, co + 1, v1);
                        }
                    }
            }
        } else {
            // fragment (c, r): e=0 -> (cout 16c+g, row 8r+2t), e=1 -> (16c+g, 8r+2t+1), e=2/3 -> cout 16c+g+8.
            // Lanes g and g^1 swap one value so that each holds two ADJACENT output channels of ONE row
            // (even g: row 8r+2t, odd g: row 8r+2t+1): float2 / packed-half2 stores instead of scalar ones.
            const bool odd = g & 1;
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = acc[c][r][e] + small[c][r][e];
                    const float r0 = __shfl_xor_sync(0xffffffffu, odd ? v[0] : v[1], 4);
                    const float r1 = __shfl_xor_sync(0xffffffffu, odd ? v[2] : v[3], 4);
                    const int64_t row = row0 + 8 * r + 2 * t + (odd ? 1 : 0);
                    if (row >= n) continue;
                    const int co = 16 * c + (g & ~1);
                    epi.store_pair(row, co, odd ? r0 : v[0], odd ? v[1] : r0);
                    epi.store_pair(row, co + 8, odd ? r1 : v[2], odd ? v[3] : r1);
                }
        }
    }
    if (epi.over && overflow) *overflow = 1;
}

// ---- dense product over h2 rows: out[n, NC] = X[n, CIN] * W[CIN, NC]  (no gather) -----------------------------------
// The generative transposed convolution (ME.MinkowskiGenerativeConvolutionTranspose k=2 s=2, autoencoder.py:155,182,
// 209) IS this product: input row i makes the 8 child rows 8i..8i+7, i.e. one row of NC = 8 * COUT values with
// W[ci][k * COUT + co] = kernel[k][ci][co]; the [8n, COUT] child tensor is the same memory as [n, 8 * COUT].
// T formulation of conv_k3_h2_kernel with the weights resident in shared memory: a warp loads the h2 fragments of
// its rows once and walks the NC / 16 output tiles, finishing (bias, ReLU, fp32 + h2 stores) each tile at once.
template <int CIN, int RG_, int WARPS_>
struct DenseH2Cfg {
    static_assert(CIN % 16 == 0, "dense h2 kernel: CIN must be a multiple of 16");
    static constexpr int KS = CIN / 16, RG = RG_, WARPS = WARPS_, THREADS = 32 * WARPS_;
    static constexpr int ROWS_PER_CTA = WARPS * 8 * RG;
    static constexpr size_t packed_words(int nc) { return (size_t)KS * (nc / 16) * 256; }
    static constexpr size_t smem_bytes(int nc) { return packed_words(nc) * 4; }
};

template <int CIN, int RG, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
dense_h2_kernel(const uint32_t *__restrict__ in, int in_ld, int64_t n, const uint32_t *__restrict__ packed, int nc,
                float inv_scale, const float *__restrict__ bias, float *__restrict__ out, int out_ld,
                uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    using C = DenseH2Cfg<CIN, RG, WARPS>;
    constexpr int KS = C::KS;
    extern __shared__ __align__(16) uint32_t wsm_dense[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int CT = nc >> 4;
    for (int i = threadIdx.x; i < KS * CT * 64; i += C::THREADS) cp_async16(wsm_dense + 4 * i, packed + 4 * i, true);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    H2Epilogue epi{bias, nullptr, out, out_h2, 0, out_ld, out_h2_ld, flags, inv_scale};
    const bool odd = g & 1;
    const int64_t n_tiles = (n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * C::ROWS_PER_CTA + warp * 8 * RG;
        uint4 x[RG][KS];
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const int64_t row = row0 + 8 * r + g;
#pragma unroll
            for (int q = 0; q < KS; ++q)
                x[r][q] = row < n ? __ldg(reinterpret_cast<const uint4 *>(in + row * in_ld + 16 * q + 4 * t)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll 2
        for (int c = 0; c < CT; ++c) {
            float part[RG][4], small[RG][4];
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) small[r][e] = 0.f;
#pragma unroll
            for (int q = 0; q < KS; ++q) {
                const uint4 *wp = reinterpret_cast<const uint4 *>(wsm_dense + (q * CT + c) * 256) + lane;
                const uint4 wh = wp[0], wl = wp[32];
#pragma unroll
                for (int r = 0; r < RG; ++r) mma_f16(small[r], wh.x, wh.y, wh.z, wh.w, x[r][q].z, x[r][q].w);
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    if (q == 0) mma_f16_zero(part[r], wh.x, wh.y, wh.z, wh.w, x[r][q].x, x[r][q].y);
                    else mma_f16(part[r], wh.x, wh.y, wh.z, wh.w, x[r][q].x, x[r][q].y);
                }
#pragma unroll
                for (int r = 0; r < RG; ++r) mma_f16(small[r], wl.x, wl.y, wl.z, wl.w, x[r][q].x, x[r][q].y);
            }
#pragma unroll
            for (int r = 0; r < RG; ++r) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = part[r][e] + small[r][e];
                const float r0 = __shfl_xor_sync(0xffffffffu, odd ? v[0] : v[1], 4);
                const float r1 = __shfl_xor_sync(0xffffffffu, odd ? v[2] : v[3], 4);
                const int64_t row = row0 + 8 * r + 2 * t + (odd ? 1 : 0);
                if (row >= n) continue;
                const int co = 16 * c + (g & ~1);
                epi.store_pair(row, co, odd ? r0 : v[0], odd ? v[1] : r0);
                epi.store_pair(row, co + 8, odd ? r1 : v[2], odd ? v[3] : r1);
            }
        }
    }
    if (epi.over && overflow) *overflow = 1;
}

}  // namespace pcgc
