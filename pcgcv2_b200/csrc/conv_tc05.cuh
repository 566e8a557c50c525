// conv_tc05.cuh -- k=3 sparse convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// One persistent CTA per SM, warp-specialised, 128 output rows per tile:
//   * 8 GATHER warps: per kernel offset every thread copies half of one neighbour row with cp.async
//     (16-byte chunks, zero-fill for missing neighbours) into a shared-memory stage laid out in the
//     UMMA canonical K-major (no-swizzle) format, then splits its own chunks in place into the TF32 hi
//     part and writes the lo part (x - hi) to the stage's second tile, fences the async proxy and
//     arrives on the stage's "full" mbarrier.  S stages are in flight, so the gather latency is covered
//     by S offsets' worth of copies per thread without holding them in registers.
//   * 1 MMA warp (one elected lane): per offset and 8-channel k-step issues three tcgen05.mma.kind::tf32
//     (A_lo*B_hi, A_hi*B_lo into the "small" accumulator; A_hi*B_hi into one of NG "main" accumulators,
//     a new one every 27/NG offsets) with the accumulators in TENSOR MEMORY, commits to the stage's
//     "empty" mbarrier, and after the 27th offset commits to the tile's "tmem_full" mbarrier.  Several
//     short accumulation chains instead of one long one: the tensor core accumulates with truncation and
//     a 162-MMA chain biased a layer by 2e-5 (measured with mma.sync); the partial sums are joined by
//     round-to-nearest FADDs in the epilogue.
//   * 4 EPILOGUE warps: tcgen05.ld the NG+1 accumulators of their 32 TMEM lanes (= rows), add them, apply
//     bias / residual / ReLU and write each output row once.  TMEM accumulators are double buffered, so
//     the epilogue of tile i overlaps the gathers and MMAs of tile i+1.
// All 27 weight matrices (hi and lo parts, pre-packed in the canonical layout) stay in shared memory.
#pragma once
#include "common.cuh"

namespace pcgc {
namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, bool pred) {
    const int bytes = pred ? 16 : 0;
#ifdef PCGC_TC_CG
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
#else       // .ca: allocate in L1 too -- neighbouring rows re-read the same lines (80 % L1 hit rate measured)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows x 16 bytes,
// LBO = byte distance between the two K core matrices of one MMA, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);                                         // version = 1 (Blackwell), layout type 0, base offset 0
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

#ifndef PCGC_TC_P
#define PCGC_TC_P 2
#endif

template <int CIN, int NPAD>
struct Cfg {
    static_assert(CIN % 8 == 0 && (NPAD == 16 || NPAD == 32 || NPAD == 64), "tcgen05 conv: CIN % 8, NPAD in {16,32,64}");
    static constexpr int TM = 128;                       // rows per tile = UMMA M
    static constexpr int KCH = CIN / 4;                  // 16-byte chunks per row
    static constexpr int KSTEPS = CIN / 8;               // MMAs (K = 8 tf32) per term per offset
#ifndef PCGC_TC_S
#define PCGC_TC_S 6
#endif
#ifndef PCGC_TC_D
#define PCGC_TC_D 3
#endif
    static constexpr int S = PCGC_TC_S;                  // shared-memory stages (slots)
    static constexpr int D = PCGC_TC_D;                          // a thread finishes offset i-D while its copies of offset i are in flight;
                                                         // S > D leaves S-D offsets of slack for the MMA to retire before a slot is reused
    static constexpr int NG = NPAD == 16 ? 9 : 3;        // main accumulators (one per 27/NG offsets)
    static constexpr int LBO = 128, SBO = KCH * 128;     // canonical K-major no-swizzle strides (bytes)
    static constexpr int A_TILE = TM * CIN * 4;          // bytes of one A tile (hi or lo)
    static constexpr int B_TILE = NPAD * CIN * 4;        // bytes of one weight tile (hi or lo)
    static constexpr int ACC_COLS = (NG + 1) * NPAD;     // TMEM columns per accumulator buffer
    static constexpr int TMEM_COLS = 2 * ACC_COLS <= 256 ? 256 : 512;
    static_assert(2 * ACC_COLS <= 512, "accumulators exceed tensor memory");
    static constexpr int GATHER_THREADS = 256, EPI_THREADS = 128, THREADS = GATHER_THREADS + EPI_THREADS + 32;
    static constexpr int CHUNKS_PER_THREAD = TM * KCH / GATHER_THREADS;
    static_assert(CHUNKS_PER_THREAD >= 1 && (TM * KCH) % GATHER_THREADS == 0, "bad chunk split");
    static constexpr size_t OFF_A_HI = 0;
    static constexpr size_t OFF_A_LO = OFF_A_HI + (size_t)S * A_TILE;
    static constexpr size_t OFF_B = OFF_A_LO + (size_t)S * A_TILE;          // [27][hi|lo][B_TILE]
    static constexpr size_t OFF_IDX = OFF_B + (size_t)27 * 2 * B_TILE;      // [2][27][TM] int32
    static constexpr size_t OFF_BAR = OFF_IDX + (size_t)2 * 27 * TM * 4;    // mbarriers + tmem base
    static constexpr size_t SMEM = OFF_BAR + 256 + 1024;                    // + slack for 1024-byte alignment
    static constexpr size_t packed_floats() { return (size_t)27 * 2 * NPAD * CIN; }
};

// byte offset of (row r, 16-byte chunk c) inside a canonical K-major no-swizzle tile
template <int KCH>
__host__ __device__ __forceinline__ int canon_off(int r, int c) { return (r >> 3) * (KCH * 128) + c * 128 + (r & 7) * 16; }

// W [27][cin][cout] -> [27][hi|lo] tiles in the canonical layout (row = output channel, K = input channel)
template <int CIN, int NPAD>
__global__ void pack_weights_tc05_kernel(const float *__restrict__ w, int cout, float *__restrict__ packed) {
    constexpr int KCH = CIN / 4;
    const int total = 27 * NPAD * CIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % CIN, nrow = (i / CIN) % NPAD, k = i / (CIN * NPAD);
        const float x = nrow < cout ? w[((int64_t)k * CIN + ci) * cout + nrow] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const int pos = canon_off<KCH>(nrow, ci >> 2) / 4 + (ci & 3);
        packed[((int64_t)k * 2 + 0) * NPAD * CIN + pos] = hi;
        packed[((int64_t)k * 2 + 1) * NPAD * CIN + pos] = x - hi;
    }
}

template <int CIN, int NPAD, int COUT>
__global__ void __launch_bounds__(Cfg<CIN, NPAD>::THREADS, 1)
conv_k3_tc05_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                    const float *__restrict__ packed, const float *__restrict__ bias,
                    const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using C = Cfg<CIN, NPAD>;
    constexpr int S = C::S, D = C::D, TM = C::TM, KCH = C::KCH;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *a_hi = sm + C::OFF_A_HI, *a_lo = sm + C::OFF_A_LO, *b_s = sm + C::OFF_B;
    int32_t *idx_s = reinterpret_cast<int32_t *>(sm + C::OFF_IDX);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
    uint64_t *full = bars, *empty = bars + S, *tmem_full = bars + 2 * S, *tmem_empty = bars + 2 * S + 2;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int MMA_WARP = (C::GATHER_THREADS + C::EPI_THREADS) / 32;

    // ---- one-time setup: weights -> smem, barriers, tensor memory
    for (int i = threadIdx.x; i < 27 * 2 * C::B_TILE / 16; i += C::THREADS)
        cp_async16_zfill(smem_u32(b_s) + 16 * i, packed + 4 * (size_t)i, true);
    cp_async_commit_group();
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, C::GATHER_THREADS / 32); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, C::EPI_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    cp_async_wait_group<0>();
    fence_proxy_async();                                  // weights were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    const int64_t n_tiles = (n + TM - 1) / TM;

    if (warp < C::GATHER_THREADS / 32) {
        // =========================== GATHER warps ===========================
        const int p = threadIdx.x;                         // 0..255
        const int row = p / (C::GATHER_THREADS / TM);      // 2 threads per row
        const int c0 = (p % (C::GATHER_THREADS / TM)) * C::CHUNKS_PER_THREAD;
        uint32_t g = 0;                                    // running stage counter across tiles
        int titer = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
            int32_t *idx_t = idx_s + (titer & 1) * 27 * TM;
            for (int i = p; i < 27 * TM; i += C::GATHER_THREADS) {
                const int k = i / TM, r = i % TM;
                const int64_t grow = tile * TM + r;
                idx_t[i] = grow < n ? __ldg(nbr + (int64_t)k * n + grow) : -1;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(C::GATHER_THREADS) : "memory");    // gather warps only
            for (int it = 0; it < 27 + D; ++it) {
                if (it < 27) {
                    const uint32_t gi = g + it, slot = gi % S, ph = (gi / S) & 1;
                    if (lane == 0) mbar_wait(empty + slot, ph ^ 1);                  // MMAs of the previous user are done
                    __syncwarp();
                    const int32_t src_row = idx_t[it * TM + row];
                    const float *src = in + (int64_t)(src_row < 0 ? 0 : src_row) * in_ld + 4 * c0;
                    const uint32_t dst = smem_u32(a_hi + (size_t)slot * C::A_TILE);
#pragma unroll
                    for (int c = 0; c < C::CHUNKS_PER_THREAD; ++c)
                        cp_async16_zfill(dst + canon_off<KCH>(row, c0 + c), src + 4 * c, src_row >= 0);
                }
                cp_async_commit_group();
                if (it >= D) {
                    const uint32_t gi = g + it - D, slot = gi % S;
                    cp_async_wait_group<D>();                                    // this thread's copies of that stage landed
                    unsigned char *hi_t = a_hi + (size_t)slot * C::A_TILE, *lo_t = a_lo + (size_t)slot * C::A_TILE;
#ifndef PCGC_TC_NO_SPLIT
#pragma unroll
                    for (int c = 0; c < C::CHUNKS_PER_THREAD; ++c) {
                        const int off = canon_off<KCH>(row, c0 + c);
                        float4 v = *reinterpret_cast<float4 *>(hi_t + off);
                        float4 h;
                        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                        *reinterpret_cast<float4 *>(hi_t + off) = h;
                        *reinterpret_cast<float4 *>(lo_t + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
#endif
#ifndef PCGC_TC_NO_FENCE
                    fence_proxy_async();
#endif
                    //                                            // own writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full + slot);                         // one arrival per warp (8 per stage)
                }
            }
            g += 27;
        }
        cp_async_wait_group<0>();
    } else if (warp == MMA_WARP) {
        // =========================== MMA warp ===========================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TM, NPAD);
            uint32_t g = 0;
            int titer = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
                const int buf = titer & 1;
                mbar_wait(tmem_empty + buf, ((titer >> 1) & 1) ^ 1);                 // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_small = tmem_base + buf * C::ACC_COLS + C::NG * NPAD;
                for (int k = 0; k < 27; ++k) {
                    const uint32_t gi = g + k, slot = gi % S, ph = (gi / S) & 1;
                    mbar_wait(full + slot, ph);
                    tc_fence_after();
                    const uint32_t a_h = smem_u32(a_hi + (size_t)slot * C::A_TILE), a_l = smem_u32(a_lo + (size_t)slot * C::A_TILE);
                    const uint32_t b_h = smem_u32(b_s + (size_t)(2 * k) * C::B_TILE), b_l = b_h + C::B_TILE;
                    const int grp = k / (27 / C::NG);
                    const uint32_t d_main = tmem_base + buf * C::ACC_COLS + grp * NPAD;
#ifndef PCGC_TC_NO_MMA
#pragma unroll
                    for (int j = 0; j < C::KSTEPS; ++j) {
                        const uint64_t ah = umma_desc(a_h + 2 * j * C::LBO, C::LBO, C::SBO), al = umma_desc(a_l + 2 * j * C::LBO, C::LBO, C::SBO);
                        const uint64_t bh = umma_desc(b_h + 2 * j * C::LBO, C::LBO, C::SBO), bl = umma_desc(b_l + 2 * j * C::LBO, C::LBO, C::SBO);
                        umma_tf32(d_small, al, bh, idesc, (k | j) != 0);
                        umma_tf32(d_small, ah, bl, idesc, 1);
                        umma_tf32(d_main, ah, bh, idesc, !(k % (27 / C::NG) == 0 && j == 0));
                    }
#endif
                    umma_commit(empty + slot);                                       // stage reusable when these MMAs retire
                }
                umma_commit(tmem_full + buf);                                        // accumulators of this tile complete
                g += 27;
            }
        }
    } else {
        // =========================== EPILOGUE warps ===========================
        const int q = warp & 3;                             // TMEM lane quarter this warp may access
        int titer = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
            const int buf = titer & 1;
            mbar_wait(tmem_full + buf, (titer >> 1) & 1);
            tc_fence_after();
            const int64_t row = tile * TM + 32 * q + lane;
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + buf * C::ACC_COLS;
#pragma unroll
            for (int cb = 0; cb < NPAD; cb += 16) {
                float acc[16];
                tmem_ld16(t0 + cb, acc);
#pragma unroll
                for (int gidx = 1; gidx <= C::NG; ++gidx) {
                    float v[16];
                    tmem_ld16(t0 + gidx * NPAD + cb, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] += v[i];
                }
                if (cb + 16 >= NPAD) {                      // last read of this buffer: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + buf);
                }
                if (row < n) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int co = cb + i;
                        if (co < COUT) {
                            float v = acc[i];
                            if (bias) v += __ldg(bias + co);
                            if (residual) v += __ldg(residual + row * res_ld + co);
                            if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
                            acc[i] = v;
                        }
                    }
                    float *o = out + row * out_ld + cb;
                    if (COUT % 4 == 0 && (out_ld & 3) == 0 && ((uintptr_t)out & 15) == 0) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (cb + i < COUT) *reinterpret_cast<float4 *>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cb + i < COUT) o[i] = acc[i];
                    }
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// Variant 2: the gathered operand goes to TENSOR MEMORY, not shared memory.
//
// ncu on variant 1 (profiles/r01_tc05_v1.txt): the shared-memory pipe is the limiter (l1tex 71 %, 150 M bank
// conflict cycles): every gathered byte crosses shared memory ~7 times (LDGSTS write, split read, two split
// writes, three tensor-core reads).  Here a gather warp loads neighbour rows with the quad-per-row pattern
// (4 lanes x 16 bytes = one fully used 64-byte segment per row; lowest L1 tag pressure), splits hi/lo in
// registers and writes both parts with tcgen05.st.16x256b.x2 straight into TMEM, whose fragment layout
// (lane/4 -> row, 2*(lane%4) -> column pair) is exactly what the quad loads produce when the contraction
// index is permuted (column 2t+e of k-step s <-> physical channel 4t+2s+e).  tcgen05.mma then takes A from
// TMEM and only the 1 KB weight tiles from shared memory.
template <int NPAD>
struct CfgT {
    static constexpr int CIN = 16, TM = 128, KSTEPS = 2;
    static constexpr int S = 6;                          // A stages in TMEM (32 columns each: hi 16 + lo 16)
    static constexpr int P = PCGC_TC_P;                  // gather warp groups; group h handles offsets gi % P == h
    static constexpr int NG = 9;
    static constexpr int ACC_COLS = (NG + 1) * NPAD;
    static constexpr int A_COL0 = 2 * ACC_COLS;          // first A-stage column
    static_assert(NPAD == 16 && A_COL0 + 32 * S <= 512, "TMEM budget");
    static constexpr int B_TILE = NPAD * CIN * 4;
    static constexpr int GATHER_WARPS = 4 * P, EPI_WARPS = 4;
    static constexpr int THREADS = 32 * (GATHER_WARPS + EPI_WARPS + 1);
    static constexpr size_t OFF_B = 0, OFF_BAR = (size_t)27 * 2 * B_TILE, SMEM = OFF_BAR + 256 + 1024;
};

// physical input channel of contraction index kappa (CIN = 16)
__host__ __device__ __forceinline__ int tc05t_phys(int kappa) { return 4 * ((kappa & 7) >> 1) + 2 * (kappa >> 3) + (kappa & 1); }

template <int NPAD>
__global__ void pack_weights_tc05t_kernel(const float *__restrict__ w, int cout, float *__restrict__ packed) {
    constexpr int CIN = 16, KCH = 4;
    const int total = 27 * NPAD * CIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kappa = i % CIN, nrow = (i / CIN) % NPAD, k = i / (CIN * NPAD);
        const float x = nrow < cout ? w[((int64_t)k * CIN + tc05t_phys(kappa)) * cout + nrow] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const int pos = canon_off<KCH>(nrow, kappa >> 2) / 4 + (kappa & 3);
        packed[((int64_t)k * 2 + 0) * NPAD * CIN + pos] = hi;
        packed[((int64_t)k * 2 + 1) * NPAD * CIN + pos] = x - hi;
    }
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

template <int NPAD, int COUT>
__global__ void __launch_bounds__(CfgT<NPAD>::THREADS, 1)
conv_k3_tc05t_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                     const float *__restrict__ packed, const float *__restrict__ bias,
                     const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using C = CfgT<NPAD>;
    constexpr int S = C::S, TM = C::TM, P = C::P;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *b_s = sm + C::OFF_B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
    uint64_t *full = bars, *empty = bars + S, *tmem_full = bars + 2 * S, *tmem_empty = bars + 2 * S + 2;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int MMA_WARP = C::GATHER_WARPS + C::EPI_WARPS;

    for (int i = threadIdx.x; i < 27 * 2 * C::B_TILE / 16; i += C::THREADS)
        cp_async16_zfill(smem_u32(b_s) + 16 * i, packed + 4 * (size_t)i, true);
    cp_async_commit_group();
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 4); mbar_init(empty + s, 1); }      // 4 gather warps fill a stage
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, C::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    cp_async_wait_group<0>();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    const int64_t n_tiles = (n + TM - 1) / TM;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_off = my_tiles * 27;              // offsets this CTA processes, numbered gi = titer*27 + k

    if (warp < C::GATHER_WARPS) {
        // =========================== GATHER warps ===========================
        const int q = warp & 3, h = warp >> 2, g = lane >> 2, t = lane & 3;
        const char *in_lane = reinterpret_cast<const char *>(in + 4 * t);
        const int64_t ld_bytes = (int64_t)in_ld * 4;
        // rows of this warp inside a tile: 32q + 8r + g, r = 0..3
        auto load_idx = [&](int64_t gi) -> int32_t {      // lane l: index of the neighbour of row 32q + l
            if (gi >= total_off) return -1;
            const int64_t tile = blockIdx.x + (gi / 27) * gridDim.x;
            const int k = (int)(gi % 27);
            const int64_t row = tile * TM + 32 * q + lane;
            return row < n ? __ldg(nbr + (int64_t)k * n + row) : -1;
        };
        auto load_rows = [&](int32_t idx_l, float4 (&x)[4]) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int32_t src = __shfl_sync(0xffffffffu, idx_l, 8 * r + g);
                x[r] = src >= 0 ? __ldg(reinterpret_cast<const float4 *>(in_lane + src * ld_bytes)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int64_t gi = h;
        int32_t idx_cur = load_idx(gi), idx_nxt = load_idx(gi + P);
        float4 x[4];
        load_rows(idx_cur, x);
        for (; gi < total_off; gi += P) {
            // prefetch: rows of the next offset of this warp, index of the one after
            float4 xn[4];
            load_rows(idx_nxt, xn);
            const int32_t idx_nn = load_idx(gi + 2 * P);
            const uint32_t slot = (uint32_t)(gi % S), ph = (uint32_t)((gi / S) & 1);
            if (lane == 0) mbar_wait(empty + slot, ph ^ 1);      // MMAs that read this stage have retired
            __syncwarp();
            tc_fence_after();
            const uint32_t a_col = tmem_base + C::A_COL0 + 32 * slot;
#pragma unroll
            for (int half = 0; half < 2; ++half) {                // TMEM lanes 32q + 16*half + (0..15): rows 8*(2half) + g and +8
                const float4 a = x[2 * half], b = x[2 * half + 1];
                float hi[8], lo[8];
                const float va[8] = {a.x, a.y, b.x, b.y, a.z, a.w, b.z, b.w};   // {row g: c0,c1 | row g+8: c0,c1 | k-step 1 likewise}
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    hi[e] = __uint_as_float(__float_as_uint(va[e]) & 0xFFFFE000u);
                    lo[e] = va[e] - hi[e];
                }
                const uint32_t taddr = a_col + ((uint32_t)(32 * q + 16 * half) << 16);
                tmem_st_16x256b_x2(taddr, hi);
                tmem_st_16x256b_x2(taddr + 16, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + slot);
#pragma unroll
            for (int r = 0; r < 4; ++r) x[r] = xn[r];
            idx_nxt = idx_nn;
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA warp ===========================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TM, NPAD);
            int64_t gi = 0;
            for (int64_t titer = 0; titer < my_tiles; ++titer) {
                const int buf = (int)(titer & 1);
                mbar_wait(tmem_empty + buf, (uint32_t)(((titer >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_small = tmem_base + buf * C::ACC_COLS + C::NG * NPAD;
                for (int k = 0; k < 27; ++k, ++gi) {
                    const uint32_t slot = (uint32_t)(gi % S), ph = (uint32_t)((gi / S) & 1);
                    mbar_wait(full + slot, ph);
                    tc_fence_after();
                    const uint32_t a_h = tmem_base + C::A_COL0 + 32 * slot, a_l = a_h + 16;
                    const uint32_t b_h = smem_u32(b_s + (size_t)(2 * k) * C::B_TILE), b_l = b_h + C::B_TILE;
                    const uint32_t d_main = tmem_base + buf * C::ACC_COLS + (k / 3) * NPAD;
#ifndef PCGC_TC_NO_MMA
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t bh = umma_desc(b_h + 2 * j * 128, 128, 4 * 128), bl = umma_desc(b_l + 2 * j * 128, 128, 4 * 128);
                        umma_tf32_ts(d_small, a_l + 8 * j, bh, idesc, (k | j) != 0);
                        umma_tf32_ts(d_small, a_h + 8 * j, bl, idesc, 1);
                        umma_tf32_ts(d_main, a_h + 8 * j, bh, idesc, !(k % 3 == 0 && j == 0));
                    }
#endif
                    umma_commit(empty + slot);
                }
                umma_commit(tmem_full + buf);
            }
        }
    } else {
        // =========================== EPILOGUE warps ===========================
        const int q = warp & 3;
        for (int64_t titer = 0; titer < my_tiles; ++titer) {
            const int64_t tile = blockIdx.x + titer * gridDim.x;
            const int buf = (int)(titer & 1);
            mbar_wait(tmem_full + buf, (uint32_t)((titer >> 1) & 1));
            tc_fence_after();
            const int64_t row = tile * TM + 32 * q + lane;
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + buf * C::ACC_COLS;
            float acc[16];
            tmem_ld16(t0, acc);
#pragma unroll
            for (int gidx = 1; gidx <= C::NG; ++gidx) {
                float v[16];
                tmem_ld16(t0 + gidx * NPAD, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] += v[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + buf);
            if (row < n) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (i < COUT) {
                        float v = acc[i];
                        if (bias) v += __ldg(bias + i);
                        if (residual) v += __ldg(residual + row * res_ld + i);
                        if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
                        acc[i] = v;
                    }
                }
                float *o = out + row * out_ld;
                if (COUT % 4 == 0 && (out_ld & 3) == 0 && ((uintptr_t)out & 15) == 0) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        if (i < COUT) *reinterpret_cast<float4 *>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < COUT) o[i] = acc[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace tc05
}  // namespace pcgc
