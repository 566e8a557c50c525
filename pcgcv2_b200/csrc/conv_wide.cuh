// conv_wide.cuh -- k=3 sparse convolution of the WIDE layers (cin >= 16) on the 5th-generation tensor cores:
// tcgen05.mma (kind::f16) with the accumulators in tensor memory, weights streamed by the TMA engine
// (cp.async.bulk), features gathered straight into the swizzled operand layout.  sm_100a only.
//
// Data flow (one persistent CTA per SM, tile = 128 Morton-consecutive output rows, 27 kernel offsets per tile):
//
//   * GATHER warps (10, each owning every 10th offset): for kernel offset k, copy the neighbour row of every output row of the tile from the pre-split
//     half-precision ("h2", conv_h2.cuh) feature tensor into a shared-memory stage with cp.async (16 bytes per lane,
//     zero fill for a missing neighbour).  An h2 row of CIN channels is CIN/4 groups of 16 bytes
//     {hi0 hi1 hi2 hi3 | lo0 lo1 lo2 lo3} (f16): copied VERBATIM it is a K-major operand row with 2 CIN contraction
//     slots -- no split, no shuffle, no register staging between global memory and the tensor core.  The stage uses
//     the canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk index XOR row & 7): the eight
//     lanes that copy one 128-byte line of a row write one conflict-free 128-byte wavefront, and a row costs one L1
//     tag lookup per 128 bytes.
//   * PRODUCER warp (one lane): streams the 27 weight tiles of the layer through the same stages with
//     cp.async.bulk (TMA, completes on the stage's mbarrier with complete_tx).  The weight tile of an offset is
//     B = [ W_hi ; W_lo ] stacked along N (2 x cout rows), every row holding the weight of a channel TWICE (against
//     the hi and the lo slot of that channel), pre-packed in global memory as the exact swizzled shared-memory image.
//   * MMA warp (one lane): per offset CIN/8 tcgen05.mma of M = 128, N = 2 cout, K = 16.  ONE pass over the gathered
//     tile yields (x_hi + x_lo) W_hi in accumulator columns [0, cout) and (x_hi + x_lo) W_lo in [cout, 2 cout):
//     the "main" and the "small" term of the split product in separate accumulators, the gathered operand read once.
//     Accumulators live in TENSOR MEMORY; the offsets of a tile are spread over NG accumulator groups so that no
//     accumulation chain is long (the tensor core adds with truncation), and the groups are joined by round-to-nearest
//     FADDs in the epilogue.  tcgen05.commit releases the stage to the gather warps and, after the 27th offset,
//     hands the accumulators to the epilogue.
//   * EPILOGUE warps (4): tcgen05.ld their 32 TMEM lanes (= rows), join the groups, multiply by the inverse weight
//     scale, add bias / residual, ReLU, and write the fp32 row and/or its h2 split (the next layer's operand) once.
//     With two accumulator buffers the epilogue of tile i overlaps the gathers and MMAs of tile i + 1.
//
// Algorithmic bytes are those of any k=3 layer; what this kernel changes is the on-chip cost per gathered
// byte: one shared-memory write and one tensor-core read, against LDG -> registers -> 3 mma.sync per 16 channels in
// conv_h2.cuh, and one launch for cout = 64 where the mma.sync path runs four 16-wide slices that re-gather every row.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_h2.cuh"      // split_pair_h2, kH2Limit

namespace pcgc {
namespace wide {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
// one lane of a converged warp (the others fall through): lets the compiler keep warp-uniform operands of the elected lane's
// tcgen05 instructions in uniform registers instead of moving them there one by one
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, bool pred) {
    const int bytes = pred ? 16 : 0;       // .ca: keep the line in L1 too -- neighbouring output rows gather the same rows
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void *src, bool pred) {
    const int bytes = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// TMA bulk copy global -> shared, completion reported to an mbarrier as transaction bytes
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor, K-major, swizzled: 8-row x RB-byte atoms (RB = 128 / 64 / 32: SWIZZLE_128B / 64B / 32B,
// layout type 2 / 4 / 6 at bits 61..63), SBO = 8 RB bytes between 8-row groups, LBO unused (1), descriptor version 1
template <int RB>
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr) {
    static_assert(RB == 128 || RB == 64 || RB == 32, "swizzle span");
    constexpr uint64_t layout = RB == 128 ? 2 : RB == 64 ? 4 : 6;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((8 * RB) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int W>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[W]) {
    static_assert(W == 8 || W == 16, "tmem_ld: 8 or 16 columns");
    uint32_t r[W];
    if constexpr (W == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = __uint_as_float(r[i]);
}

#ifndef PCGC_WIDE_NG
#define PCGC_WIDE_NG 0          // 0: per-shape default below
#endif
#ifndef PCGC_WIDE_G
#define PCGC_WIDE_G 10
#endif
#ifndef PCGC_WIDE_SMEM_KB
#define PCGC_WIDE_SMEM_KB 224   // stage budget (measured: more stages in flight beat a bigger L1, profiles/r02_wide_tcgen05_*.txt)
#endif


template <int CIN, int COUT>
struct WCfg {
    static_assert(CIN == 8 || CIN == 16 || CIN == 32 || CIN == 64, "wide conv: CIN in {8, 16, 32, 64}");
    static_assert(COUT == 1 || COUT == 4 || COUT == 8 || COUT == 16 || COUT == 32 || COUT == 64, "wide conv: COUT");
    static constexpr int TM = 128;                          // output rows per tile = UMMA M
    static constexpr int RB = 4 * CIN < 128 ? 4 * CIN : 128; // bytes of a row inside one K block = the swizzle span
    static constexpr int KB = 4 * CIN / RB;                 // K blocks per row (a block = up to 32 channels x (hi + lo) f16)
    static constexpr int CPR = RB / 16;                     // 16-byte chunks per row and block
    static constexpr int KSTEPS = RB / 32;                  // tcgen05.mma (K = 16 f16 = 32 bytes) per block
    static constexpr int NP = COUT < 8 ? 8 : COUT;          // padded output channels
    static constexpr int N2 = 2 * NP;                       // UMMA N: main | small
    static constexpr int EW = NP >= 16 ? 16 : 8;            // accumulator columns per tcgen05.ld
    static constexpr int A_BYTES = KB * TM * RB, B_BYTES = KB * N2 * RB, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int GATHER_WARPS = PCGC_WIDE_G;        // each owns every GATHER_WARPS-th offset
    static constexpr int NBUF = 2;                          // accumulator buffers in tensor memory
    static constexpr int NG = PCGC_WIDE_NG ? PCGC_WIDE_NG : (N2 == 128 ? 2 : N2 == 64 ? 4 : N2 == 32 ? 8 : 9);
    static constexpr int ACC_COLS = NG * N2;
    static_assert(NBUF * ACC_COLS <= 512, "accumulators exceed tensor memory");
    static constexpr int TMEM_COLS = NBUF * ACC_COLS <= 32 ? 32 : NBUF * ACC_COLS <= 64 ? 64 : NBUF * ACC_COLS <= 128 ? 128 :
                                     NBUF * ACC_COLS <= 256 ? 256 : 512;
    static constexpr int BAR_BYTES = 512;                   // 2 S + 2 NBUF mbarriers + the TMEM base address
    static constexpr int IDX_BYTES = CIN >= 32 ? 2 * 27 * TM * 4 : 0;                // lockstep mode: two kernel-map slices
    static constexpr int BUDGET = 227 * 1024 - 1024 - BAR_BYTES - IDX_BYTES;
    static constexpr int WANT = (PCGC_WIDE_SMEM_KB * 1024 < BUDGET ? PCGC_WIDE_SMEM_KB * 1024 : BUDGET) / STAGE_BYTES;
    static constexpr int S = WANT > 16 ? 16 : (WANT < 2 ? 2 : WANT);                   // stages = offsets in flight per SM
    static_assert(S * STAGE_BYTES <= BUDGET, "wide conv: stages exceed shared memory");
    // Gather organisation (measured, profiles/r02_wide_tcgen05_*.txt):
    //  * LOCKSTEP (cin >= 32: few, big stages): all gather warps share the rows of every offset and each keeps D older
    //    offsets in flight (cp.async groups); the tile's kernel-map slice is staged in shared memory;
    //  * otherwise (cin <= 16: many small stages): ONE warp per offset, min(GATHER_WARPS, S) offsets in flight per SM -- a
    //    warp's wait-stage / issue / wait-data / proxy-fence / arrive round trip overlaps with the other warps' offsets.
    static constexpr bool LOCKSTEP = CIN >= 32;
    static constexpr int D = S - 1 > 4 ? 4 : S - 1;
    static constexpr int TEAM = LOCKSTEP ? GATHER_WARPS : 1;                          // arrivals per stage from the gather side
    static constexpr int NT = GATHER_WARPS < S ? GATHER_WARPS : S;                    // per-warp mode: active warps (<= S: parity waits)
    static constexpr int EPI_WARPS = 4, MMA_WARP = 4, PROD_WARP = 5, FIRST_GATHER = 6;
    static constexpr int THREADS = 32 * (FIRST_GATHER + GATHER_WARPS);
    static constexpr size_t OFF_STAGE = 0, OFF_IDX = (size_t)S * STAGE_BYTES, OFF_BAR = OFF_IDX + IDX_BYTES;
    static constexpr size_t SMEM = OFF_BAR + BAR_BYTES + 1024;   // + slack for the 1024-byte alignment SWIZZLE_128B needs
    static_assert((2 * S + 2 * NBUF) * 8 + 8 <= BAR_BYTES, "barrier region too small");
    static constexpr size_t packed_bytes() { return (size_t)27 * B_BYTES; }
    __host__ __device__ static constexpr int group_of(int k) { return k * NG / 27; }
};

// byte offset of (row r, 16-byte chunk c of K block kb) in a K-major swizzled tile of ROWS rows of RB bytes: the chunk
// index is XORed with address bits [7, 7 + log2(RB / 16)) = the row index scaled to 128-byte lines
template <int ROWS, int RB>
__host__ __device__ __forceinline__ int sw_off(int r, int kb, int c) {
    return kb * (ROWS * RB) + r * RB + ((c ^ ((r * RB >> 7) & (RB / 16 - 1))) << 4);
}

// W [27][cin][cout] (scaled by a power of two) -> per offset the shared-memory image of B = [W_hi ; W_lo]:
// row nrow < NP holds f16(W s) of output channel nrow, row NP + co holds f16(W s - hi); contraction slot
// 8 (CPR kb + c) + e addresses channel 4 (CPR kb + c) + (e & 3) -- e < 4 meets the hi half of the gathered group, e >= 4 the lo half
template <int CIN, int COUT>
__global__ void pack_weights_wide_kernel(const float *__restrict__ w, float scale, __half *__restrict__ packed) {
    using C = WCfg<CIN, COUT>;
    const int per_off = C::B_BYTES / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 27 * per_off; i += gridDim.x * blockDim.x) {
        const int k = i / per_off, q = i % per_off;
        const int e = q & 7, c = (q >> 3) % C::CPR, nrow = (q / (8 * C::CPR)) % C::N2, kb = q / (8 * C::CPR * C::N2);
        const int ch = 4 * (C::CPR * kb + c) + (e & 3), co = nrow % C::NP;
        float v = 0.f;
        if (co < COUT) {
            const float x = w[((int64_t)k * CIN + ch) * COUT + co] * scale;
            const float hi = __half2float(__float2half_rn(x));
            v = nrow < C::NP ? hi : x - hi;
        }
        packed[(int64_t)k * per_off + (sw_off<C::N2, C::RB>(nrow, kb, c) >> 1) + e] = __float2half_rn(v);
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(WCfg<CIN, COUT>::THREADS, 1)
conv_k3_wide_kernel(const uint32_t *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                    const unsigned char *__restrict__ packed, float inv_scale, const float *__restrict__ bias,
                    const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld,
                    uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    using C = WCfg<CIN, COUT>;
    constexpr int S = C::S, TM = C::TM, KB = C::KB, NP = C::NP, N2 = C::N2, NG = C::NG, NBUF = C::NBUF, RB = C::RB, CPR = C::CPR;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
    uint64_t *full = bars, *empty = bars + S, *tmem_full = bars + 2 * S, *tmem_empty = bars + 2 * S + NBUF;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 2 * S + 2 * NBUF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: barriers, tensor memory
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, C::TEAM + 1); mbar_init(empty + s, 1); }   // full: the gather team + the weight producer
        for (int b = 0; b < NBUF; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, C::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == C::MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    const int64_t n_tiles = (n + TM - 1) / TM;

    if (warp >= C::FIRST_GATHER) {
        // =========================== GATHER warps ===========================
        constexpr int RPI = 32 / CPR;                            // rows per warp instruction (4 x 128 / 8 x 64 / 16 x 32 bytes)
        const int gw = warp - C::FIRST_GATHER, sub = lane / CPR, c = lane % CPR;
        const uint32_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
        if constexpr (C::LOCKSTEP) {
            constexpr int GT = 32 * C::GATHER_WARPS, D = C::D;
            constexpr int INSTR = TM / RPI * KB;                 // warp instructions per offset, dealt round the warps
            int32_t *idx_s = reinterpret_cast<int32_t *>(sm + C::OFF_IDX);
            const int gt = threadIdx.x - 32 * C::FIRST_GATHER;
            auto stage_idx = [&](int64_t tile, int32_t *dst) {   // kernel-map slice of a tile -> shared memory (async)
                for (int i = gt; i < 27 * TM; i += GT) {
                    const int k = i / TM, r = i % TM;
                    const int64_t grow = tile * TM + r;
                    const bool ok = grow < n;                    // rows past the end read as index 0: gathered, never stored
                    cp_async4_zfill(smem_u32(dst + i), nbr + (ok ? (int64_t)k * n + grow : 0), ok);
                }
            };
            if (my_tiles) stage_idx(blockIdx.x, idx_s);
            cp_async_commit();
            cp_async_wait<0>();
            asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");
            uint32_t g = 0;                                      // offsets issued so far (stage = g % S)
            for (uint32_t titer = 0; titer < my_tiles; ++titer) {
                const int64_t tile = blockIdx.x + (int64_t)titer * gridDim.x;
                const int32_t *idx_t = idx_s + (titer & 1) * 27 * TM;
                if (titer + 1 < my_tiles) stage_idx(tile + gridDim.x, idx_s + ((titer + 1) & 1) * 27 * TM);   // joins the first group below
                for (int it = 0; it < 27 + D; ++it) {
                    if (it < 27) {
                        const uint32_t gi = g + it, slot = gi % S, ph = (gi / S) & 1;
                        if (lane == 0) mbar_wait(empty + slot, ph ^ 1);      // the MMAs that read this stage have retired
                        __syncwarp();
                        const uint32_t a_s = smem_u32(sm + C::OFF_STAGE + (size_t)slot * C::STAGE_BYTES);
#pragma unroll
                        for (int i = 0; i < (INSTR + C::GATHER_WARPS - 1) / C::GATHER_WARPS; ++i) {
                            const int q = gw + C::GATHER_WARPS * i;
                            if (q < INSTR) {
                                const int row = RPI * (q / KB) + sub, kb = q % KB;
                                const int32_t src_row = idx_t[it * TM + row];
                                const uint32_t *src = in + (int64_t)(src_row < 0 ? 0 : src_row) * in_ld + 4 * (CPR * kb + c);
                                cp_async16_zfill(a_s + sw_off<TM, RB>(row, kb, c), src, src_row >= 0);
                            }
                        }
                    }
                    cp_async_commit();
                    if (it >= D) {
                        cp_async_wait<D>();                                  // this thread's copies of offset it - D have landed
                        fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full + (g + it - D) % S);
                    }
                }
                g += 27;
                asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");        // next tile's kernel-map slice is complete for all
            }
            cp_async_wait<0>();
        } else {
            // one warp per offset: warp gw owns g = gw, gw + NT, ... (g = 27 * tile iteration + k) and copies all 128 rows.
            // Only NT = min(GATHER_WARPS, S) warps take part: a warp waits for its stage by PHASE PARITY, which is sound only
            // while no waiter is more than one phase ahead of the barrier -- with NT <= S the stage of g + NT was last used by
            // g + NT - S <= g, which this warp has itself seen retire.
            constexpr uint32_t NT = C::NT;
            const uint32_t total = my_tiles * 27u;
            auto load_idx = [&](uint32_t g, int32_t (&idx)[4]) { // kernel-map entries of rows lane, lane + 32, ... of offset g
                const int64_t tile = blockIdx.x + (int64_t)(g / 27u) * gridDim.x;
                const uint32_t k = g % 27u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t grow = tile * TM + 32 * j + lane;
                    idx[j] = (g < total && grow < n) ? __ldg(nbr + (int64_t)k * n + grow) : -1;
                }
            };
            int32_t idx_cur[4], idx_nxt[4];
            load_idx((uint32_t)gw, idx_cur);
            for (uint32_t g = (uint32_t)gw; (uint32_t)gw < NT && g < total; g += NT) {
                load_idx(g + NT, idx_nxt);                       // in flight during this offset's copies
                const uint32_t slot = g % S, ph = (g / S) & 1;
                if (lane == 0) mbar_wait(empty + slot, ph ^ 1);  // the MMAs that read this stage have retired
                __syncwarp();
                const uint32_t a_s = smem_u32(sm + C::OFF_STAGE + (size_t)slot * C::STAGE_BYTES);
#pragma unroll
                for (int q = 0; q < TM / RPI; ++q) {             // RPI rows x one K block per instruction
                    const int row = RPI * q + sub;
                    const int32_t src_row = __shfl_sync(0xffffffffu, idx_cur[(RPI * q) / 32], (RPI * q) % 32 + sub);
                    const uint32_t *src = in + (int64_t)(src_row < 0 ? 0 : src_row) * in_ld + 4 * c;
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        cp_async16_zfill(a_s + sw_off<TM, RB>(row, kb, c), src + 4 * CPR * kb, src_row >= 0);
                }
                cp_async_commit();
                cp_async_wait<0>();                              // this warp's copies of the offset have landed
                fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(full + slot);
#pragma unroll
                for (int j = 0; j < 4; ++j) idx_cur[j] = idx_nxt[j];
            }
        }
    } else if (warp == C::PROD_WARP) {
        // =========================== weight PRODUCER (TMA bulk copies) ===========================
        if (lane == 0) {
            uint32_t g = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int k = 0; k < 27; ++k, ++g) {
                    const uint32_t slot = g % S, ph = (g / S) & 1;
                    mbar_wait(empty + slot, ph ^ 1);
                    mbar_arrive_expect_tx(full + slot, C::B_BYTES);
                    tma_bulk_g2s(smem_u32(sm + C::OFF_STAGE + (size_t)slot * C::STAGE_BYTES + C::A_BYTES),
                                 packed + (size_t)k * C::B_BYTES, C::B_BYTES, full + slot);
                }
            }
        }
    } else if (warp == C::MMA_WARP) {
        // =========================== MMA issuer ===========================
        // The whole warp runs the loop converged and one ELECTED lane issues, so that descriptors and TMEM addresses are
        // provably warp-uniform and live in uniform registers (a lone `lane == 0` branch costs ~18 instructions of
        // register -> uniform-register moves plus an elect loop per tcgen05.mma: the issuing thread, not the tensor pipe,
        // was the bottleneck of the first version -- profiles/r02_octet_tcgen05_v1_ncu.txt).
        constexpr uint32_t idesc = umma_idesc_f16(TM, N2);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t stage0 = smem_u32(sm + C::OFF_STAGE);
        const uint64_t desc0 = umma_desc_sw<RB>(0);                          // everything but the start address
        uint32_t g = 0;
        int titer = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
            const int buf = titer % NBUF;
            mbar_wait(tmem_empty + buf, ((titer / NBUF) & 1) ^ 1);           // the epilogue has drained this buffer
            tc_fence_after();
            const uint32_t d0 = tmem_u + buf * C::ACC_COLS;
            for (int k = 0; k < 27; ++k, ++g) {
                const uint32_t slot = g % S, ph = (g / S) & 1;
                mbar_wait(full + slot, ph);                                  // gathered rows + weight tile have landed
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_s = stage0 + slot * C::STAGE_BYTES, b_s = a_s + C::A_BYTES;
                    const uint64_t adesc = desc0 | (uint64_t)((a_s & 0x3FFFF) >> 4), bdesc = desc0 | (uint64_t)((b_s & 0x3FFFF) >> 4);
                    const int grp = C::group_of(k);
                    const bool first = k == 0 || C::group_of(k - 1) != grp;
                    const uint32_t d = d0 + grp * N2;
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                        for (int j = 0; j < C::KSTEPS; ++j)                  // K = 16 f16 = 32 bytes inside the swizzled row
                            umma_f16(d, adesc + (uint64_t)((kb * TM * RB + 32 * j) >> 4), bdesc + (uint64_t)((kb * N2 * RB + 32 * j) >> 4),
                                     idesc, !(first && kb == 0 && j == 0));
                    umma_commit(empty + slot);                               // stage reusable when these MMAs retire
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(tmem_full + buf);              // accumulators of this tile complete
            __syncwarp();
        }
    } else {
        // =========================== EPILOGUE warps (warp w owns TMEM lanes 32 w .. 32 w + 31) ===========================
        constexpr int EW = C::EW;
        bool over = false;
        const bool vec_out = out && (out_ld & 3) == 0 && ((uintptr_t)out & 15) == 0;
        const bool vec_res = residual && (res_ld & 3) == 0 && ((uintptr_t)residual & 15) == 0;
        const bool vec_h2 = out_h2 && (out_h2_ld & 3) == 0 && ((uintptr_t)out_h2 & 15) == 0;
        int titer = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
            const int buf = titer % NBUF;
            mbar_wait(tmem_full + buf, (titer / NBUF) & 1);
            tc_fence_after();
            const int64_t row = tile * TM + 32 * warp + lane;
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * warp) << 16) + buf * C::ACC_COLS;
#pragma unroll
            for (int cb = 0; cb < NP; cb += EW) {
                float small[EW], acc[EW];
                tmem_ld<EW>(t0 + NP + cb, small);
                tmem_ld<EW>(t0 + cb, acc);
#pragma unroll
                for (int gi = 1; gi < NG; ++gi) {
                    float v[EW];
                    tmem_ld<EW>(t0 + gi * N2 + NP + cb, v);
#pragma unroll
                    for (int i = 0; i < EW; ++i) small[i] += v[i];
                    tmem_ld<EW>(t0 + gi * N2 + cb, v);
#pragma unroll
                    for (int i = 0; i < EW; ++i) acc[i] += v[i];
                }
                if (cb + EW >= NP) {                                         // last read of this buffer: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + buf);
                }
                if (row < n) {
#pragma unroll
                    for (int i = 0; i < EW; ++i) {
                        const int co = cb + i;
                        float v = (acc[i] + small[i]) * inv_scale;
                        if (co < COUT) {
                            if (bias) v += __ldg(bias + co);
                            if (residual && !vec_res) v += __ldg(residual + row * res_ld + co);
                        }
                        acc[i] = v;
                    }
                    if (vec_res && COUT % 4 == 0) {
#pragma unroll
                        for (int i = 0; i < EW; i += 4)
                            if (cb + i < COUT) {
                                const float4 r = __ldg(reinterpret_cast<const float4 *>(residual + row * res_ld + cb + i));
                                acc[i] += r.x; acc[i + 1] += r.y; acc[i + 2] += r.z; acc[i + 3] += r.w;
                            }
                    } else if (vec_res) {
#pragma unroll
                        for (int i = 0; i < EW; ++i)
                            if (cb + i < COUT) acc[i] += __ldg(residual + row * res_ld + cb + i);
                    }
                    if (flags & PCGC_EPI_RELU) {
#pragma unroll
                        for (int i = 0; i < EW; ++i) acc[i] = fmaxf(acc[i], 0.f);
                    }
                    if (out) {
                        float *o = out + row * out_ld + cb;
                        if (vec_out && COUT % 4 == 0) {
#pragma unroll
                            for (int i = 0; i < EW; i += 4)
                                if (cb + i < COUT) *reinterpret_cast<float4 *>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < EW; ++i)
                                if (cb + i < COUT) o[i] = acc[i];
                        }
                    }
                    if (out_h2 && COUT % 4 == 0) {
                        uint32_t *o = out_h2 + row * out_h2_ld + cb;
#pragma unroll
                        for (int i = 0; i < EW; i += 4)
                            if (cb + i < COUT) {
                                over |= !(fabsf(acc[i]) <= kH2Limit) || !(fabsf(acc[i + 1]) <= kH2Limit) || !(fabsf(acc[i + 2]) <= kH2Limit) ||
                                        !(fabsf(acc[i + 3]) <= kH2Limit);
                                uint4 h;
                                split_pair_h2(acc[i], acc[i + 1], h.x, h.z);
                                split_pair_h2(acc[i + 2], acc[i + 3], h.y, h.w);
                                if (vec_h2) *reinterpret_cast<uint4 *>(o + i) = h;
                                else { o[i] = h.x; o[i + 1] = h.y; o[i + 2] = h.z; o[i + 3] = h.w; }
                            }
                    }
                }
            }
        }
        if (over && overflow) *overflow = 1;
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == C::MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

}  // namespace wide
}  // namespace pcgc
