// conv_pipe.cuh -- k=3 sparse convolution on the tensor cores, software-pipelined (round-1 v5).
//
// Same arithmetic as conv_mma.cuh (3xTF32 mma.sync.m16n8k8, main term joined by an RN FADD per
// kernel offset, gathered rows go from global memory straight into MMA fragments through the
// permuted contraction index), with two changes that the ncu profile of conv_k3_mma_kernel asked
// for (profiles/r01_conv_mma16x16.txt: tensor pipe 52 % busy, the rest long-scoreboard stalls at
// the batch boundaries; and Cout <= 8 layers paying for a 16-wide M tile):
//
//  * the gather of offset o+D-1 is issued before the math of offset o (D register stages per
//    lane, rotating; the 27 offsets are walked in a loop unrolled by D so the stage index is a
//    compile-time constant) -- a warp always has D-1 offsets of loads in flight while it multiplies;
//  * two formulations, picked per layer shape:
//      T  (Cout >= 16): D^T[cout,row] += W_k^T * X_k^T -- weights are the 16x8 A operand, 8 gathered
//                       rows the B operand (as conv_mma.cuh);
//      NT (Cout <= 8):  D[row,cout] += X_k * W_k -- 16 gathered rows are the A operand, the weights the
//                       8x8 B operand: an 8-wide N tile is enough, half the MMAs of T for these shapes.
//
// Weights (hi/lo TF32 parts, fragment order) stay resident in shared memory: only shapes whose 27
// packed offsets fit are instantiated (see PipeCfg::fits); wider layers keep conv_mma.cuh's streamed path.
#pragma once
#include "common.cuh"
#include "conv_mma.cuh"

namespace pcgc {

template <int CIN, int COUT, bool NT_, int RG_, int D_>
struct PipeCfg {
    static_assert(CIN == 8 || CIN % 16 == 0, "pipe kernel: CIN must be 8 or a multiple of 16");
    static constexpr bool NT = NT_;
    static constexpr int KS = CIN / 8;                               // k-steps (8 channels) per offset
    static constexpr int CT = NT ? (COUT + 7) / 8 : (COUT + 15) / 16;  // output-channel tiles (N=8 / M=16)
    static constexpr int CHUNKS = CIN >= 16 ? CIN / 16 : 1;         // loads per gathered row per lane
    static constexpr int AV = CIN >= 16 ? 4 : 2;                    // floats per load
    static constexpr int RG = RG_, D = D_;
    static constexpr int GROUP_ROWS = NT ? 16 : 8;
    static constexpr int NR = NT ? 2 * RG : RG;                     // gathered rows per lane per offset
    static constexpr int RPW = GROUP_ROWS * RG;                     // output rows per warp
    static constexpr int THREADS = 256;
    static constexpr int ROWS_PER_CTA = (THREADS / 32) * RPW;
    static constexpr int W_OFF = KS * CT * (NT ? 128 : 256);        // packed floats per offset (hi + lo)
    static constexpr size_t packed_floats() { return (size_t)27 * W_OFF; }
    static constexpr size_t smem_bytes() { return packed_floats() * 4 + (size_t)(THREADS / 32) * 27 * RPW * 4; }
    static constexpr bool fits = smem_bytes() <= 200 * 1024;
};

// which kernel serves a (cin, cout) k=3 layer through the packed-weights entry point
enum { kRouteMma = 0, kRoutePipeT = 1, kRoutePipeNT = 2 };
constexpr int pipe_route(int cin, int cout) {
    if (!(cin == 8 || cin == 16 || cin == 32)) return kRouteMma;     // cin 64: registers / smem favour the streamed kernel
    if (cout <= 8) return kRoutePipeNT;
    if (cout == 16 && cin <= 16) return kRoutePipeT;
    return kRouteMma;
}

// tuning defaults from the B200 variant sweep (tools/bench_pipe.cu, profiles/r01_pipe_sweep.txt)
template <int CIN, int COUT>
struct PipeTune {
    static constexpr bool NT = pipe_route(CIN, COUT) == kRoutePipeNT;
    static constexpr int RG = NT ? 2 : 4;
    static constexpr int D = 2;
    static constexpr int MINB = CIN == 8 ? (NT ? 4 : 3) : 2;
    static constexpr int OPT = 39;
};

// NT weight packing: W [27][cin][cout] -> [27][KS][CT][32 lanes][4] = {b0 hi, b1 hi, b0 lo, b1 lo},
// b0 = W[phys(ks,t,0)][8j+g], b1 = W[phys(ks,t,1)][8j+g]   (B fragment of m16n8k8: (k=t, n=g), (k=t+4, n=g))
static __global__ void pack_weights_nt_kernel(const float *__restrict__ w, int kvol, int cin, int cout, float *__restrict__ packed) {
    const int KS = cin / 8, CT = (cout + 7) / 8;
    const int64_t total = (int64_t)kvol * KS * CT * 32 * 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 1), lane = (int)((i >> 1) & 31);
        int64_t r = i >> 6;
        const int j = (int)(r % CT); r /= CT;
        const int ks = (int)(r % KS);
        const int k = (int)(r / KS);
        const int g = lane >> 2, t = lane & 3;
        const int ci = mma_phys_channel(cin, ks, t, e), co = 8 * j + g;
        const float x = co < cout ? w[((int64_t)k * cin + ci) * cout + co] : 0.f;
        float hi, lo;
        split_tf32_f(x, hi, lo);
        float *dst = packed + ((((int64_t)k * KS + ks) * CT + j) * 32 + lane) * 4;
        dst[e] = hi;
        dst[2 + e] = lo;
    }
}

// D = A*B + C with every fragment as scalars (A 16x8: a0..a3, B 8x8: b0,b1)
__device__ __forceinline__ void mma_tf32_s(float (&d)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void mma_tf32_s_zero(float (&d)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};\n"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}

// T weight packing, bank-conflict-free: W [27][cin][cout] -> [27][KS][CT][2 (hi, lo)][32 lanes][4] = {a0..a3}
// (conv_mma.cuh keeps hi and lo of a lane adjacent: 32-byte lane stride, a 2-way conflict on every LDS.128 --
// 23 M conflict cycles per launch in profiles/r01_conv_mma16x16.txt)
static __global__ void pack_weights_t_kernel(const float *__restrict__ w, int kvol, int cin, int cout, float *__restrict__ packed) {
    const int KS = cin / 8, CT = (cout + 15) / 16;
    const int64_t total = (int64_t)kvol * KS * CT * 32 * 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 3), lane = (int)((i >> 2) & 31);
        int64_t r = i >> 7;
        const int m = (int)(r % CT); r /= CT;
        const int ks = (int)(r % KS);
        const int k = (int)(r / KS);
        const int g = lane >> 2, t = lane & 3;
        const int ci = mma_phys_channel(cin, ks, t, e >> 1), co = 16 * m + g + 8 * (e & 1);
        const float x = co < cout ? w[((int64_t)k * cin + ci) * cout + co] : 0.f;
        float hi, lo;
        split_tf32_f(x, hi, lo);
        float *dst = packed + (((int64_t)k * KS + ks) * CT + m) * 256 + lane * 4 + e;
        dst[0] = hi;
        dst[128] = lo;
    }
}

// OPT bits (kept switchable for the variant sweep, tools/bench_pipe.cu): 1 = dependent MMAs spaced RG apart,
// 2 = no per-offset "any neighbour?" vote/branch (lets ptxas interleave the next gather with the MMAs),
// 4 = the lane's NR kernel-map entries of one offset are adjacent in smem (one vector LDS instead of NR)
template <int CIN, int COUT, bool NT, int RG, int D, int MINB, int OPT = 7>
__global__ void __launch_bounds__(256, MINB)
conv_k3_pipe_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                    const float *__restrict__ packed, const float *__restrict__ bias,
                    const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using C = PipeCfg<CIN, COUT, NT, RG, D>;
    constexpr int KS = C::KS, CT = C::CT, CHUNKS = C::CHUNKS, AV = C::AV, NR = C::NR, RPW = C::RPW, W_OFF = C::W_OFF;
    constexpr bool SPACED = (OPT & 1) != 0, NOSKIP = (OPT & 2) != 0, IDXV = (OPT & 4) != 0 && (NR == 2 || NR == 4);
    constexpr bool PREFETCH = (OPT & 32) != 0;   // L2 prefetch of this CTA's next tile (input rows + kernel-map slice)
    constexpr bool DIAG_NOMMA = (OPT & 8) != 0, DIAG_NOGATHER = (OPT & 16) != 0;   // diagnostics only (wrong results): isolate the gather / the math
    extern __shared__ __align__(16) float wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    int32_t *idx_s = reinterpret_cast<int32_t *>(wsm + C::packed_floats()) + warp * 27 * RPW;

    for (int i = threadIdx.x; i < 27 * W_OFF / 4; i += C::THREADS) cp_async16(wsm + 4 * i, packed + 4 * i, true);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // gathered-row addressing: the kernel-map slice is staged as row offsets in 16-byte units (in_ld % 4 == 0),
    // so a lane's address is one IMAD.WIDE.U32: base(lane) + off16 * 16
    const char *in_lane = reinterpret_cast<const char *>(in + (AV == 4 ? 4 * t : 2 * t));
    const int32_t ld16 = in_ld >> 2;
    const int64_t n_tiles = (n + C::ROWS_PER_CTA - 1) / C::ROWS_PER_CTA;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * C::ROWS_PER_CTA + warp * RPW;     // first row of this warp
        if constexpr (PREFETCH) {
            const int64_t nt = tile + gridDim.x;                      // the rows the next tile of this CTA owns are the
            if (nt < n_tiles) {                                       // first-touch (DRAM-latency) part of its gathers
                const int64_t r0 = nt * C::ROWS_PER_CTA;
                const int64_t rows = n - r0 < C::ROWS_PER_CTA ? n - r0 : C::ROWS_PER_CTA;
                const char *p = reinterpret_cast<const char *>(in + r0 * in_ld);
                const int64_t bytes = rows * in_ld * 4;
                for (int64_t off = (int64_t)threadIdx.x * 128; off < bytes; off += C::THREADS * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
                constexpr int LPS = (C::ROWS_PER_CTA * 4 + 127) / 128;          // 128-byte lines per kernel-map segment
                for (int i = threadIdx.x; i < 27 * LPS; i += C::THREADS) {
                    const int k = i / LPS, l = i % LPS;
                    if ((int64_t)l * 32 < rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(nbr + (int64_t)k * n + r0 + l * 32));
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < 27 * RPW; i += 32) {                   // this warp's slice of the kernel map, coalesced
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

        float x[D][NR][CHUNKS][AV];
        bool have[D];
        // lane's gathered rows of one group: T: row 8r+g; NT: rows 16r+g and 16r+8+g  (local row 8j+g)
        int32_t idn[NR];                                              // kernel-map entries of the NEXT offset to gather
        auto load_idx = [&](int o) {
            if constexpr (IDXV && NR == 4) {
                const int4 v = *reinterpret_cast<const int4 *>(idx_s + o * RPW + g * 4);
                idn[0] = v.x; idn[1] = v.y; idn[2] = v.z; idn[3] = v.w;
            } else if constexpr (IDXV && NR == 2) {
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
            if (o + 1 < 27) load_idx(o + 1);                          // one offset ahead: its LDS latency hides behind this offset's math
            bool any = false;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const bool ok = id[j] >= 0;
                any |= ok;
#pragma unroll
                for (int q = 0; q < CHUNKS; ++q) {
                    const char *src = in_lane + (uint64_t)(uint32_t)id[j] * 16u + 64 * q;
                    if constexpr (DIAG_NOGATHER) {
#pragma unroll
                        for (int e = 0; e < AV; ++e) x[st][j][q][e] = __int_as_float(id[j] + e);
                    } else if constexpr (AV == 4) {
                        const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        x[st][j][q][0] = v.x; x[st][j][q][1] = v.y; x[st][j][q][2] = v.z; x[st][j][q][3] = v.w;
                    } else {
                        const float2 v = ok ? __ldg(reinterpret_cast<const float2 *>(src)) : make_float2(0.f, 0.f);
                        x[st][j][q][0] = v.x; x[st][j][q][1] = v.y;
                    }
                }
            }
            have[st] = any;
        };
        auto math = [&](int o, int st) {
            if constexpr (!NOSKIP) {
                if (!__any_sync(0xffffffffu, have[st])) return;       // no row of the warp has this neighbour
            }
            if constexpr (DIAG_NOMMA) {
#pragma unroll
                for (int j = 0; j < NR; ++j)
#pragma unroll
                    for (int q = 0; q < CHUNKS; ++q)
#pragma unroll
                        for (int e = 0; e < AV; ++e) acc[0][j % RG][e] += x[st][j][q][e];
                return;
            }
            const float *wb = wsm + (size_t)o * W_OFF;
            float part[CT][RG][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int q = CIN >= 16 ? ks >> 1 : 0, s = CIN >= 16 ? 2 * (ks & 1) : 0;
                if constexpr (NT) {
                    float4 w[CT];
#pragma unroll
                    for (int c = 0; c < CT; ++c) w[c] = *reinterpret_cast<const float4 *>(wb + ((ks * CT + c) * 32 + lane) * 4);
                    float h[RG][4], l[RG][4];
#pragma unroll
                    for (int r = 0; r < RG; ++r) {
                        const float a[4] = {x[st][2 * r][q][s], x[st][2 * r + 1][q][s], x[st][2 * r][q][s + 1], x[st][2 * r + 1][q][s + 1]};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { h[r][e] = tf32_hi(a[e]); l[r][e] = a[e] - h[r][e]; }
                    }
                    if constexpr (SPACED) {
#pragma unroll
                        for (int c = 0; c < CT; ++c)
#pragma unroll
                            for (int r = 0; r < RG; ++r) mma_tf32_s(small[c][r], l[r][0], l[r][1], l[r][2], l[r][3], w[c].x, w[c].y);   // X_lo * W_hi
#pragma unroll
                        for (int c = 0; c < CT; ++c)
#pragma unroll
                            for (int r = 0; r < RG; ++r) {                                                                              // X_hi * W_hi
                                if (ks == 0) mma_tf32_s_zero(part[c][r], h[r][0], h[r][1], h[r][2], h[r][3], w[c].x, w[c].y);
                                else mma_tf32_s(part[c][r], h[r][0], h[r][1], h[r][2], h[r][3], w[c].x, w[c].y);
                            }
#pragma unroll
                        for (int c = 0; c < CT; ++c)
#pragma unroll
                            for (int r = 0; r < RG; ++r) mma_tf32_s(small[c][r], h[r][0], h[r][1], h[r][2], h[r][3], w[c].z, w[c].w);   // X_hi * W_lo
                    } else {
#pragma unroll
                        for (int r = 0; r < RG; ++r)
#pragma unroll
                            for (int c = 0; c < CT; ++c) {
                                mma_tf32_s(small[c][r], l[r][0], l[r][1], l[r][2], l[r][3], w[c].x, w[c].y);
                                mma_tf32_s(small[c][r], h[r][0], h[r][1], h[r][2], h[r][3], w[c].z, w[c].w);
                                if (ks == 0) mma_tf32_s_zero(part[c][r], h[r][0], h[r][1], h[r][2], h[r][3], w[c].x, w[c].y);
                                else mma_tf32_s(part[c][r], h[r][0], h[r][1], h[r][2], h[r][3], w[c].x, w[c].y);
                            }
                    }
                } else {
                    float h[RG][2], l[RG][2];
#pragma unroll
                    for (int r = 0; r < RG; ++r) {
                        const float b0 = x[st][r][q][s], b1 = x[st][r][q][s + 1];
                        h[r][0] = tf32_hi(b0); h[r][1] = tf32_hi(b1);
                        l[r][0] = b0 - h[r][0]; l[r][1] = b1 - h[r][1];
                    }
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const float4 *wp = reinterpret_cast<const float4 *>(wb + (ks * CT + c) * 256) + lane;
                        const float4 wh = wp[0], wl = wp[32];
                        if constexpr (SPACED) {
#pragma unroll
                            for (int r = 0; r < RG; ++r) mma_tf32(small[c][r], wh, l[r][0], l[r][1]);          // W_hi * X_lo
#pragma unroll
                            for (int r = 0; r < RG; ++r) {                                                     // W_hi * X_hi
                                if (ks == 0) mma_tf32_zero(part[c][r], wh, h[r][0], h[r][1]);
                                else mma_tf32(part[c][r], wh, h[r][0], h[r][1]);
                            }
#pragma unroll
                            for (int r = 0; r < RG; ++r) mma_tf32(small[c][r], wl, h[r][0], h[r][1]);          // W_lo * X_hi
                        } else {
#pragma unroll
                            for (int r = 0; r < RG; ++r) {
                                mma_tf32(small[c][r], wh, l[r][0], l[r][1]);
                                mma_tf32(small[c][r], wl, h[r][0], h[r][1]);
                                if (ks == 0) mma_tf32_zero(part[c][r], wh, h[r][0], h[r][1]);
                                else mma_tf32(part[c][r], wh, h[r][0], h[r][1]);
                            }
                        }
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
        for (int ob = 0; ob < 27; ob += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int o = ob + d;
                if (o < 27) {
                    if (o + D - 1 < 27) gather(o + D - 1, (d + D - 1) % D);
                    math(o, d);
                }
            }
        }

        // ---- epilogue
        if constexpr (NT) {
            // fragment (c, r): e=0 -> (row 16r+g, cout 8c+2t), e=1 -> (.., 8c+2t+1), e=2/3 -> row 16r+g+8
            const bool vec = ((out_ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0) && (COUT % 2 == 0) &&
                             (!residual || (((res_ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(residual) & 7) == 0)));
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const int co = 8 * c + 2 * t;
                if (co >= COUT) continue;
                const float b0 = bias ? __ldg(bias + co) : 0.f;
                const float b1 = (bias && co + 1 < COUT) ? __ldg(bias + co + 1) : 0.f;
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int64_t row = row0 + 16 * r + 8 * h + g;
                        if (row >= n) continue;
                        float v0 = acc[c][r][2 * h] + small[c][r][2 * h] + b0;
                        float v1 = acc[c][r][2 * h + 1] + small[c][r][2 * h + 1] + b1;
                        if (vec) {
                            if (residual) {
                                const float2 rv = __ldg(reinterpret_cast<const float2 *>(residual + row * res_ld + co));
                                v0 += rv.x; v1 += rv.y;
                            }
                            if (flags & PCGC_EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                            *reinterpret_cast<float2 *>(out + row * out_ld + co) = make_float2(v0, v1);
                        } else {
                            if (residual) v0 += __ldg(residual + row * res_ld + co);
                            if (flags & PCGC_EPI_RELU) v0 = fmaxf(v0, 0.f);
                            out[row * out_ld + co] = v0;
                            if (co + 1 < COUT) {
                                if (residual) v1 += __ldg(residual + row * res_ld + co + 1);
                                if (flags & PCGC_EPI_RELU) v1 = fmaxf(v1, 0.f);
                                out[row * out_ld + co + 1] = v1;
                            }
                        }
                    }
            }
        } else {
            // fragment (c, r): e=0 -> (cout 16c+g, row 8r+2t), e=1 -> (16c+g, 8r+2t+1), e=2/3 -> cout 16c+g+8
#pragma unroll
            for (int c = 0; c < CT; ++c) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int co = 16 * c + g + 8 * h;
                    if (co >= COUT) continue;
                    const float bv = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
                    for (int r = 0; r < RG; ++r)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int64_t row = row0 + 8 * r + 2 * t + e;
                            if (row >= n) continue;
                            float v = acc[c][r][2 * h + e] + small[c][r][2 * h + e] + bv;
                            if (residual) v += __ldg(residual + row * res_ld + co);
                            if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
                            out[row * out_ld + co] = v;
                        }
                }
            }
        }
    }
}

}  // namespace pcgc
