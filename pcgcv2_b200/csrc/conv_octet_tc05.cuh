// conv_octet_tc05.cuh -- k=3 sparse convolution on FULL-OCTET sets (8-child expansions: the three synthesis levels, where 55 %
// of the algorithmic bytes of the codec are) on the 5th-generation tensor cores, WITHOUT materialising a gathered operand:
// the tcgen05 shared-memory descriptors address the 27 shifted views of ONE staged halo directly.
//
// Geometry.  Rows of a full-octet set are ordered 8 p + c (child c = cx + 2 cy + 4 cz of parent p).  The output children of
// an octet at kernel offset d = (dx, dy, dz) read the positions (cx + dx + 1, cy + dy + 1, cz + dz + 1) of the octet's
// 4 x 4 x 4 voxel halo.  A tile is 32 consecutive octets (256 output rows).  Its halos are staged in shared memory as
//
//     halo[pz][py][px][grp][kc][j] : 16 bytes        octet o = 8 grp + j,  kc = 16-byte chunk of the h2 row (4 channels)
//
// i.e. for one halo position and one K chunk the eight octets of a group are 128 contiguous bytes -- exactly one 8-row x
// 16-byte "core matrix" of the canonical K-major (no-swizzle) UMMA layout, with LBO = 128 (next K chunk) and
// SBO = E (next octet group, and -- four groups on -- the next px; E = KC x 144: 16 bytes of padding per chunk).  For a fixed (cy, cz, d) the 64 rows
// (cx, o) therefore form ONE valid M = 64 operand: row group rg = 4 cx + grp sits at base + rg E, because
// px = cx + dx + 1 advances the address by exactly 4 E.  The kernel offset is nothing but a different descriptor start
// address: 27 offsets x 4 (cy, cz) combinations x CIN/8 K steps = 216 tcgen05.mma (M 64, N 2 cout, K 16) per tile and not a
// single byte is gathered, shuffled or re-staged per offset.  (mma.sync path, conv_octet_h2.cuh: per octet 27 x LDS.128
// fragment reads + 81 HMMA at 8 cycles each per SM sub-partition -- the issue rate and the shared-memory wavefronts bound it.)
//
// Accumulators live in tensor memory.  M = 64 uses half of the TMEM lanes (rows 16 q + i -> lane 32 q + i); the (cy = 0) and
// (cy = 1) combinations are interleaved in the two lane halves (lane offset 0 / 16) of the same columns, so one 32-lane
// tcgen05.ld returns both.  Per combination three accumulator groups (one per dz: chains of 18 MMAs -- the tensor core adds
// with truncation), main | small columns as in conv_wide.cuh (B = [W_hi ; W_lo] stacked along N, every weight against the
// hi and the lo slot of its channel); the groups are joined by round-to-nearest FADDs in the epilogue.
//
// Pipeline (one persistent CTA per SM): the halo is four z-plane buffers of 16 positions; the MMAs of a tile visit the
// planes in order (plane p serves the (cz, dz) pairs with cz + dz + 1 = p), so plane p of tile t + 1 is refilled as soon as
// its MMAs of tile t retire -- a 4-deep ring without a second halo buffer.  8 FILL warps copy halo rows with cp.async (one
// instruction = one position x 8 octets x 64 bytes, conflict-free 128-byte wavefronts; the source row of a position is a
// child of one of the 27 PARENT neighbours: the child set's own kernel map is never built or read), 2 MMA warps issue (one
// per cy: their accumulators are disjoint), 4 EPILOGUE warps drain the double-buffered accumulators, the weights of all 27
// offsets are brought in once by TMA bulk copies and stay resident.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_wide.cuh"      // mbarrier / tcgen05 / TMA helpers

namespace pcgc {
namespace otc {

using namespace wide;

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes (128 contiguous bytes), LBO = byte
// distance between the two K core matrices of one MMA, SBO = byte distance between 8-row groups; descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_none(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

template <int CIN, int COUT>
struct OCfg {
    static_assert(CIN == 16, "octet tcgen05 conv: CIN == 16");
    static_assert(COUT == 1 || COUT == 4 || COUT == 8 || COUT == 16, "octet tcgen05 conv: COUT in {1, 4, 8, 16}");
    static constexpr int TO = 32;                           // octets per tile (4 groups of 8)
    static constexpr int KC = CIN / 4;                      // 16-byte chunks per h2 row
    static constexpr int KSTEPS = KC / 2;                   // tcgen05.mma (K = 16 f16 = two chunks) per offset
    // bytes between the K chunks of a (position, group) block: 128 of payload (8 octets x 16 bytes) + 16 of padding, so that the
    // four chunks of ONE source row (which arrive together from L2 and are written together) fall into four different bank
    // quads (measured without the padding: 21.5 M shared-memory bank conflicts of the cp.async writes per launch)
    static constexpr int LBO_A = 144;
    static constexpr int E = KC * LBO_A;                    // bytes of one (position, group): 8 octets x the whole row
    static constexpr int EB = KC * 128;                     // the same for the weight tiles (bulk-copied: no padding needed)
    static constexpr int PLANE_BYTES = 16 * 4 * E;          // one z-plane of the halo: 16 positions x 4 groups
    static constexpr int NP = COUT < 8 ? 8 : COUT, N2 = 2 * NP, EW = NP >= 16 ? 16 : 8;
    static constexpr int B_BYTES = N2 / 8 * EB;             // weight tile of one offset: [N2 rows][2 CIN f16], canonical layout
    static constexpr int W_BYTES = 27 * B_BYTES;
    static constexpr int NG = 3, NBUF = 2;                  // accumulator groups (one per dz) and buffers
    static constexpr int ACC_COLS = 2 * NG * N2;            // per buffer: [cz][dz group][main | small]; cy = the lane half
    static_assert(NBUF * ACC_COLS <= 512, "accumulators exceed tensor memory");
    static constexpr int FILL_WARPS = 8, EPI_WARPS = 4, MMA_WARP0 = 4, FIRST_FILL = 6;
    static constexpr int THREADS = 32 * (FIRST_FILL + FILL_WARPS);
    static constexpr int IDX_BYTES = 2 * 27 * TO * 4;       // parent-neighbour rows of a tile, double buffered
    static constexpr size_t OFF_HALO = 0, OFF_W = 4 * PLANE_BYTES, OFF_IDX = OFF_W + W_BYTES, OFF_BAR = OFF_IDX + IDX_BYTES;
    static constexpr size_t SMEM = OFF_BAR + 256 + 1024;
    static_assert(SMEM <= 227 * 1024, "octet tcgen05 conv: shared memory budget");
    static constexpr size_t packed_bytes() { return (size_t)W_BYTES; }
};

// W [27][cin][cout] (scaled) -> per offset the canonical no-swizzle image of B = [W_hi ; W_lo]: row n, contraction slot
// 8 kc + e (channel 4 kc + (e & 3); e < 4 meets the hi half of the h2 group, e >= 4 the lo half)
template <int CIN, int COUT>
__global__ void pack_weights_otc_kernel(const float *__restrict__ w, float scale, __half *__restrict__ packed) {
    using C = OCfg<CIN, COUT>;
    const int per_off = C::B_BYTES / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 27 * per_off; i += gridDim.x * blockDim.x) {
        const int k = i / per_off, q = i % per_off;
        const int e = q & 7, j = (q >> 3) & 7, kc = (q >> 6) % C::KC, ng = q / (64 * C::KC);
        const int nrow = 8 * ng + j, ch = 4 * kc + (e & 3), co = nrow % C::NP;
        float v = 0.f;
        if (co < COUT) {
            const float x = w[((int64_t)k * CIN + ch) * COUT + co] * scale;
            const float hi = __half2float(__float2half_rn(x));
            v = nrow < C::NP ? hi : x - hi;
        }
        packed[i] = __float2half_rn(v);                      // i = ((k * N2/8 + ng) * KC + kc) * 64 + 8 j + e: already the image
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(OCfg<CIN, COUT>::THREADS, 1)
conv_k3_octet_tc05_kernel(const uint32_t *__restrict__ in, int in_ld, const int32_t *__restrict__ pnbr, int64_t n_par,
                          const unsigned char *__restrict__ packed, float inv_scale, const float *__restrict__ bias,
                          const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld,
                          uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    using C = OCfg<CIN, COUT>;
    constexpr int TO = C::TO, KC = C::KC, E = C::E, NP = C::NP, N2 = C::N2, NG = C::NG, NBUF = C::NBUF;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    int32_t *idx_s = reinterpret_cast<int32_t *>(sm + C::OFF_IDX);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
    uint64_t *full = bars, *empty = bars + 4, *tmem_full = bars + 8, *tmem_empty = bars + 10, *w_ready = bars + 12;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bars + 13);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int p = 0; p < 4; ++p) { mbar_init(full + p, C::FILL_WARPS); mbar_init(empty + p, 2); }        // two MMA warps release a plane
        for (int b = 0; b < NBUF; ++b) { mbar_init(tmem_full + b, 2); mbar_init(tmem_empty + b, C::EPI_WARPS); }
        mbar_init(w_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == C::MMA_WARP0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    const int64_t n_tiles = (n_par + TO - 1) / TO;
    const uint32_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;

    if (warp >= C::FIRST_FILL) {
        // =========================== FILL warps ===========================
        constexpr int FT = 32 * C::FILL_WARPS;
        const int fw = warp - C::FIRST_FILL, ft = threadIdx.x - 32 * C::FIRST_FILL;
        const int kc = lane >> 3, j = lane & 7;                  // a warp instruction: one position, one group, 8 octets x 4 chunks
        if (fw == 0 && lane == 0) {                              // the weights of all 27 offsets, once, by TMA bulk copies
            mbar_arrive_expect_tx(w_ready, C::W_BYTES);
            for (int k = 0; k < 27; ++k)
                tma_bulk_g2s(smem_u32(sm + C::OFF_W + (size_t)k * C::B_BYTES), packed + (size_t)k * C::B_BYTES, C::B_BYTES, w_ready);
        }
        // parent-neighbour rows of a tile: idx[nd][o] = pnbr[nd][p0 + o] (-1 past the end), prefetched one tile ahead
        constexpr int IPT = (27 * TO + FT - 1) / FT;
        int32_t nxt[IPT];
        auto fetch_idx = [&](int64_t tile) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int e = ft + FT * i, nd = e / TO, o = e % TO;
                const int64_t p = tile * TO + o;
                nxt[i] = (e < 27 * TO && tile < n_tiles && p < n_par) ? __ldg(pnbr + (int64_t)nd * n_par + p) : -1;
            }
        };
        auto store_idx = [&](int32_t *dst) {
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                if (ft + FT * i < 27 * TO) dst[ft + FT * i] = nxt[i];
        };
        fetch_idx(blockIdx.x);
        uint32_t g = 0;                                          // planes issued so far (plane = g & 3)
        for (uint32_t titer = 0; titer < my_tiles; ++titer) {
            const int64_t tile = blockIdx.x + (int64_t)titer * gridDim.x;
            int32_t *idx_t = idx_s + (titer & 1) * 27 * TO;
            store_idx(idx_t);
            fetch_idx(tile + gridDim.x);                         // lands while this tile is filled
            asm volatile("bar.sync 1, %0;" ::"n"(FT) : "memory");
            for (int it = 0; it < 4 + 1; ++it) {                 // one older plane in flight per thread
                if (it < 4) {
                    const uint32_t ph = ((g + it) >> 2) & 1;
                    if (lane == 0) mbar_wait(empty + it, ph ^ 1);            // plane `it` of the previous tile has been consumed
                    __syncwarp();
                    const uint32_t plane_s = smem_u32(sm + C::OFF_HALO + (size_t)it * C::PLANE_BYTES);
                    const int ndz = it == 0 ? -1 : (it == 3 ? 1 : 0), iz = it == 0 ? 1 : (it == 3 ? 0 : it - 1);
#pragma unroll
                    for (int i = 0; i < 64 / C::FILL_WARPS; ++i) {           // 16 positions x 4 groups dealt round the warps
                        const int q = fw + C::FILL_WARPS * i, pos = q >> 2, grp = q & 3, px = pos & 3, py = pos >> 2;
                        const int ndx = px == 0 ? -1 : (px == 3 ? 1 : 0), ix = px == 0 ? 1 : (px == 3 ? 0 : px - 1);
                        const int ndy = py == 0 ? -1 : (py == 3 ? 1 : 0), iy = py == 0 ? 1 : (py == 3 ? 0 : py - 1);
                        const int nd = (ndz + 1) * 9 + (ndy + 1) * 3 + (ndx + 1), child = ix + 2 * iy + 4 * iz;
                        const int32_t par = idx_t[nd * TO + 8 * grp + j];
                        const uint32_t *src = in + ((int64_t)(par < 0 ? 0 : par) * 8 + child) * in_ld + 4 * kc;
                        wide::cp_async16_zfill(plane_s + (pos * 4 + grp) * E + kc * C::LBO_A + j * 16, src, par >= 0);
                    }
                }
                wide::cp_async_commit();
                if (it >= 1) {
                    wide::cp_async_wait<1>();                                      // this thread's copies of plane it - 1 have landed
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full + (it - 1));
                }
            }
            g += 4;
        }
        wide::cp_async_wait<0>();
    } else if (warp >= C::MMA_WARP0) {
        // =========================== MMA issuers: warp MMA_WARP0 + cy issues the (cy, *) combinations ===========================
        // The whole warp runs this loop converged and one ELECTED lane issues: every operand (descriptors, TMEM addresses) is
        // then provably warp-uniform and stays in uniform registers -- issued from a lone `lane == 0` branch each tcgen05.mma
        // cost ~18 instructions of register -> uniform-register moves and an elect loop (measured: the two issuing threads
        // were busy 75 % of the time while the tensor pipe was 20 % active).
        const int cy = __shfl_sync(0xffffffffu, warp - C::MMA_WARP0, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        constexpr uint32_t idesc = umma_idesc_f16(64, N2);
        mbar_wait(w_ready, 0);
        const uint32_t w_s = smem_u32(sm + C::OFF_W), halo_s = smem_u32(sm + C::OFF_HALO);
        const uint64_t desc_a = umma_desc_none(0, C::LBO_A, E), desc_b = umma_desc_none(0, 128, C::EB);   // all but the start address
        for (uint32_t titer = 0; titer < my_tiles; ++titer) {
            const uint32_t buf = titer % NBUF;
            mbar_wait(tmem_empty + buf, ((titer / NBUF) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_u + ((uint32_t)(16 * cy) << 16) + buf * C::ACC_COLS;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                mbar_wait(full + p, titer & 1);                              // plane p of this tile has landed
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t plane_s = halo_s + p * C::PLANE_BYTES;
#pragma unroll
                    for (int cz = 0; cz < 2; ++cz) {
                        const int dz = p - 1 - cz;                           // cz + dz + 1 == p
                        if (dz < -1 || dz > 1) continue;
                        const uint32_t d = d0 + (cz * NG + (dz + 1)) * N2;
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy) {
                            const uint32_t row_s = plane_s + (uint32_t)((cy + dy + 1) * 16) * E;
#pragma unroll
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int k = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
                                const uint64_t adesc = desc_a | (uint64_t)(((row_s + (dx + 1) * 4 * E) & 0x3FFFF) >> 4);
                                const uint64_t bdesc = desc_b | (uint64_t)(((w_s + k * C::B_BYTES) & 0x3FFFF) >> 4);
#pragma unroll
                                for (int ks = 0; ks < C::KSTEPS; ++ks)                      // two K chunks on
                                    umma_f16(d, adesc + (2 * C::LBO_A / 16) * ks, bdesc + 16 * ks, idesc, !(dy == -1 && dx == -1 && ks == 0));
                            }
                        }
                    }
                    umma_commit(empty + p);                                  // plane reusable when these MMAs retire
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(tmem_full + buf);
            __syncwarp();
        }
    } else {
        // =========================== EPILOGUE warps ===========================
        // M = 64 accumulators: row m = 16 q + i lives in TMEM lane 32 q + i (+ 16 for the cy = 1 half); warp q reads its 32 lanes
        constexpr int EW = C::EW;
        bool over = false;
        const bool vec_out = out && (out_ld & 3) == 0 && ((uintptr_t)out & 15) == 0;
        const bool vec_res = residual && (res_ld & 3) == 0 && ((uintptr_t)residual & 15) == 0;
        const bool vec_h2 = out_h2 && (out_h2_ld & 3) == 0 && ((uintptr_t)out_h2 & 15) == 0;
        const int cy = lane >> 4, m = 16 * warp + (lane & 15), cx = m >> 5, o = m & 31;
        for (uint32_t titer = 0; titer < my_tiles; ++titer) {
            const int64_t tile = blockIdx.x + (int64_t)titer * gridDim.x;
            const int buf = titer % NBUF;
            mbar_wait(tmem_full + buf, (titer / NBUF) & 1);
            tc_fence_after();
            const int64_t par = tile * TO + o;
            const uint32_t t0 = tmem_base + ((uint32_t)(32 * warp) << 16) + buf * C::ACC_COLS;
#pragma unroll
            for (int cz = 0; cz < 2; ++cz) {
                const int64_t row = par * 8 + (cx + 2 * cy + 4 * cz);
#pragma unroll
                for (int cb = 0; cb < NP; cb += EW) {
                    float small[EW], acc[EW];
                    tmem_ld<EW>(t0 + (cz * NG) * N2 + NP + cb, small);
                    tmem_ld<EW>(t0 + (cz * NG) * N2 + cb, acc);
#pragma unroll
                    for (int gi = 1; gi < NG; ++gi) {
                        float v[EW];
                        tmem_ld<EW>(t0 + (cz * NG + gi) * N2 + NP + cb, v);
#pragma unroll
                        for (int i = 0; i < EW; ++i) small[i] += v[i];
                        tmem_ld<EW>(t0 + (cz * NG + gi) * N2 + cb, v);
#pragma unroll
                        for (int i = 0; i < EW; ++i) acc[i] += v[i];
                    }
                    if (cz == 1 && cb + EW >= NP) {                          // last read of this buffer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tmem_empty + buf);
                    }
                    if (par < n_par) {
#pragma unroll
                        for (int i = 0; i < EW; ++i) {
                            const int co = cb + i;
                            float v = (acc[i] + small[i]) * inv_scale;
                            if (co < COUT) {
                                if (bias) v += __ldg(bias + co);
                                if (residual && !(vec_res && COUT % 4 == 0)) v += __ldg(residual + row * res_ld + co);
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
                        }
                        if (flags & PCGC_EPI_RELU) {
#pragma unroll
                            for (int i = 0; i < EW; ++i) acc[i] = fmaxf(acc[i], 0.f);
                        }
                        if (out) {
                            float *op = out + row * out_ld + cb;
                            if (vec_out && COUT % 4 == 0) {
#pragma unroll
                                for (int i = 0; i < EW; i += 4)
                                    if (cb + i < COUT) *reinterpret_cast<float4 *>(op + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                            } else {
#pragma unroll
                                for (int i = 0; i < EW; ++i)
                                    if (cb + i < COUT) op[i] = acc[i];
                            }
                        }
                        if (out_h2 && COUT % 4 == 0) {
                            uint32_t *op = out_h2 + row * out_h2_ld + cb;
#pragma unroll
                            for (int i = 0; i < EW; i += 4)
                                if (cb + i < COUT) {
                                    over |= !(fabsf(acc[i]) <= kH2Limit) || !(fabsf(acc[i + 1]) <= kH2Limit) || !(fabsf(acc[i + 2]) <= kH2Limit) ||
                                            !(fabsf(acc[i + 3]) <= kH2Limit);
                                    uint4 h;
                                    split_pair_h2(acc[i], acc[i + 1], h.x, h.z);
                                    split_pair_h2(acc[i + 2], acc[i + 3], h.y, h.w);
                                    if (vec_h2) *reinterpret_cast<uint4 *>(op + i) = h;
                                    else { op[i] = h.x; op[i + 1] = h.y; op[i + 2] = h.z; op[i + 3] = h.w; }
                                }
                        }
                    }
                }
            }
        }
        if (over && overflow) *overflow = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::MMA_WARP0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace otc
}  // namespace pcgc
