// conv_octet_h2.cuh -- k=3 sparse convolution on FULL-OCTET sets over pre-split half-precision features:
// the halo staging of conv_octet.cuh (4 x 4 x 4 voxel halo of each octet in shared memory, filled by cp.async
// from contiguous sibling runs, addressed by the PARENT's kernel map) feeding the arithmetic of conv_h2.cuh
// (mma.sync.m16n8k16 f16, three products per term pair, operands straight from 16-byte loads).  The halo rows
// are h2 rows (four channels = {hi0 hi1 | hi2 hi3 | lo0 lo1 | lo2 lo3}), so an LDS.128 of a halo row IS the
// lane's MMA fragments: the 27-offset loop is LDS.128 + HMMA only -- no split, no index loads, no predicates.
//
//  NT (COUT <= 8):  16 rows = the same child g of two octets are the A operand, weights the B operand;
//  T  (COUT % 16 == 0): 8 rows = the children of one octet are the B operand, weights the A operand.
#pragma once
#include "conv_h2.cuh"
#include "conv_octet.cuh"

namespace pcgc {

template <int CIN, int COUT, bool NT_, int RG_, int WARPS_, bool DB_ = false>
struct OctetH2Cfg {
    static constexpr bool DB = DB_;                               // two halo buffers per warp: the next tile is staged during the MMAs
    static_assert(CIN == 8 || CIN == 16 || CIN == 32, "octet h2 kernel: CIN in {8, 16, 32}");
    static_assert(NT_ || COUT % 16 == 0, "octet h2 kernel, T formulation: COUT must be a multiple of 16");
    static constexpr bool NT = NT_;
    static constexpr bool C8 = CIN == 8;                          // 32-byte rows: [hi | lo] of the 8 channels = one k-step (conv_h2.cuh)
    static constexpr int KS = C8 ? 1 : CIN / 16;                  // k-steps = 64-byte chunks per row
    static constexpr int CT = NT ? (COUT + 7) / 8 : COUT / 16;
    static constexpr int RG = RG_, WARPS = WARPS_, THREADS = 32 * WARPS_;
    static constexpr int OW = NT ? 2 * RG : RG;                   // octets per warp iteration
    static constexpr int NR = OW;                                 // halo rows read per lane per offset
    static constexpr int ROWB = CIN * 4;
    static constexpr int PPR = CIN / 4;
    static constexpr int SY = C8 ? 6 : 4, SZ = 4 * SY;            // CIN 8: y rows 16 banks apart (LDS.64 of rows {0,1,6,7} is conflict free)
    static constexpr int HROWS = 3 * SZ + 3 * SY + 4;
    static constexpr int HB = HROWS * ROWB;
    static constexpr int W_OFF = KS * CT * (NT ? 128 : 256);
    static constexpr size_t packed_words() { return (size_t)27 * W_OFF; }
    static constexpr size_t buf_bytes() { return ((size_t)OW * HB + (size_t)27 * OW * 4 + 127) / 128 * 128; }
    static constexpr size_t warp_bytes() { return (DB ? 2 : 1) * buf_bytes(); }
    static constexpr size_t smem_bytes() { return packed_words() * 4 + (size_t)WARPS * warp_bytes(); }
    static constexpr int OCTETS_PER_CTA = WARPS * OW;
};

// wait until the z-planes that kernel-offset plane iz reads have landed, with EXTRA younger cp.async groups (the next
// tile's four planes) allowed to stay in flight
template <int EXTRA>
__device__ __forceinline__ void halo_wait_db(int iz) {
    if (iz == 0) cp_async_wait<2 + EXTRA>();
    else if (iz == 1) cp_async_wait<1 + EXTRA>();
    else cp_async_wait<EXTRA>();
    __syncwarp();
}

template <int CIN, int COUT, bool NT, int RG, int WARPS, int MINB, bool DB = false>
__global__ void __launch_bounds__(32 * WARPS, MINB)
conv_k3_octet_h2_kernel(const uint32_t *__restrict__ in, int in_ld, const int32_t *__restrict__ pnbr, int64_t n_par,
                        const uint32_t *__restrict__ packed, float inv_scale, const float *__restrict__ bias,
                        const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld,
                        uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    using C = OctetH2Cfg<CIN, COUT, NT, RG, WARPS, DB>;
    constexpr int KS = C::KS, CT = C::CT, OW = C::OW, NR = C::NR, ROWB = C::ROWB, PPR = C::PPR;
    constexpr int SY = C::SY, SZ = C::SZ, HB = C::HB, W_OFF = C::W_OFF;
    constexpr bool C8 = C::C8;
    extern __shared__ __align__(128) unsigned char smem_oh2[];
    uint32_t *wsm = reinterpret_cast<uint32_t *>(smem_oh2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    unsigned char *halo0 = smem_oh2 + C::packed_words() * 4 + (size_t)warp * C::warp_bytes();   // per buffer: halos, then [27][OW] parent rows

    for (int i = threadIdx.x; i < 27 * W_OFF / 4; i += C::THREADS) cp_async16(wsm + 4 * i, packed + 4 * i, true);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // lane's read base: child g of an octet, 16-byte piece t of each 64-byte chunk (chunk index XOR x parity for wide rows)
    const int cx = g & 1, cy = (g >> 1) & 1, cz = g >> 2;
    int slot[KS];                                                   // byte offsets inside a halo buffer
#pragma unroll
    for (int s = 0; s < KS; ++s) slot[s] = (cx + SY * cy + SZ * cz) * ROWB + (KS >= 2 ? ((s ^ cx) * 64) : 0) + t * (C8 ? 8 : 16);

    const int64_t n_tiles = (n_par + C::OCTETS_PER_CTA - 1) / C::OCTETS_PER_CTA;
    const char *in_bytes = reinterpret_cast<const char *>(in);
    const uint32_t ldb = (uint32_t)in_ld * 4u;
    H2Epilogue epi{bias, residual, out, out_h2, res_ld, out_ld, out_h2_ld, flags, inv_scale};
    int32_t prow[OW];                                               // parent rows of the next tile to stage (lane k: neighbour k)
#ifndef PCGC_KNOCKOUT                                               // tuning builds only (tools/knockout.sh): drop one phase, time the rest
#define PCGC_KNOCKOUT 0
#endif
    constexpr bool ko_fill = PCGC_KNOCKOUT & 1, ko_xlds = PCGC_KNOCKOUT & 2, ko_wlds = PCGC_KNOCKOUT & 4, ko_mma = PCGC_KNOCKOUT & 8;
    auto stage = [&](unsigned char *buf) {                          // rows -> this buffer's index slice -> cp.async of the halos
        int32_t *sidx = reinterpret_cast<int32_t *>(buf + (size_t)OW * HB);
        store_parent_rows<OW>(sidx, prow, lane);
        __syncwarp();
        if constexpr (!ko_fill) halo_fill<PPR, OW, HB, ROWB, SY, SZ, (KS >= 2)>(buf, sidx, in_bytes, ldb, lane);
    };
    // tile order: strided (tile = blockIdx.x + i * gridDim.x) or, with PCGC_TILES_CHUNKED in `flags`, one contiguous run of
    // tiles per CTA: Morton-consecutive tiles share a face of their halos, which then still sits in this SM's L1
    const bool chunked = flags & PCGC_TILES_CHUNKED;
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t tile_begin = chunked ? blockIdx.x * per_cta : blockIdx.x, tile_step = chunked ? 1 : gridDim.x;
    const int64_t tile_end = chunked ? min(n_tiles, tile_begin + per_cta) : n_tiles;
    load_parent_rows<OW>(prow, pnbr, n_par, tile_begin * C::OCTETS_PER_CTA + warp * OW, lane);
    int cur = 0;
    if constexpr (DB) {                                             // prologue: the first tile's halos
        stage(halo0);
        load_parent_rows<OW>(prow, pnbr, n_par, (tile_begin + tile_step) * C::OCTETS_PER_CTA + warp * OW, lane);
    }
    for (int64_t tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int64_t oct0 = tile * C::OCTETS_PER_CTA + warp * OW;                // first octet of this warp
        __syncwarp();                                                             // previous iteration's readers are done
        unsigned char *halo;
        if constexpr (DB) {                // stage tile + gridDim.x into the other buffer: its loads fly during this tile's MMAs
            halo = halo0 + (size_t)cur * C::buf_bytes();
            stage(halo0 + (size_t)(cur ^ 1) * C::buf_bytes());                    // (past the end: rows are -1, zero fill, no traffic)
            load_parent_rows<OW>(prow, pnbr, n_par, (tile + 2 * tile_step) * C::OCTETS_PER_CTA + warp * OW, lane);
            cur ^= 1;
        } else {
            halo = halo0;
            stage(halo0);
            load_parent_rows<OW>(prow, pnbr, n_par, (tile + tile_step) * C::OCTETS_PER_CTA + warp * OW, lane);
        }

        float acc[CT][RG][4], small[CT][RG][4];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[c][r][e] = small[c][r][e] = 0.f;
        halo_wait_db<DB ? 4 : 0>(0);

        uint4 xb[2][NR][KS];
        auto load_frags = [&](int o, uint4 (&x)[NR][KS]) {
            const int ix = o % 3, iy = (o / 3) % 3, iz = o / 9;
            const int doff = (ix + SY * iy + SZ * iz) * ROWB;
#pragma unroll
            for (int j = 0; j < NR; ++j)
#pragma unroll
                for (int q = 0; q < KS; ++q) {
                    if constexpr (C8) {
                        const uint2 v = *reinterpret_cast<const uint2 *>(halo + slot[0] + j * HB + doff);
                        x[j][q] = make_uint4(v.x, v.y, 0u, 0u);
                    } else {
                        x[j][q] = *reinterpret_cast<const uint4 *>(halo + slot[KS >= 2 ? (q ^ (ix & 1)) : 0] + j * HB + doff);
                    }
                }
        };
        load_frags(0, xb[0]);
#pragma unroll
        for (int o = 0; o < 27; ++o) {
            if ((o + 1) % 9 == 0 && o + 1 < 27) halo_wait_db<DB ? 4 : 0>((o + 1) / 9);   // next z-plane of the halo
            if (o + 1 < 27 && !ko_xlds) load_frags(o + 1, xb[(o + 1) & 1]);
            uint4 (&x)[NR][KS] = xb[ko_xlds ? 0 : (o & 1)];                       // (ko_xlds is a compile-time constant)
            const uint32_t *wb = wsm + (size_t)o * W_OFF;
            float part[CT][RG][4];
#pragma unroll
            for (int q = 0; q < KS; ++q) {
                if constexpr (C8 && NT) {                                                              // two MMAs per offset (conv_h2.cuh)
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(wb + (c * 32 + lane) * 4);
#pragma unroll
                        for (int r = 0; r < RG; ++r)
                            mma_f16_zero(part[c][r], x[2 * r][0].x, x[2 * r + 1][0].x, x[2 * r][0].y, x[2 * r + 1][0].y, w.x, w.y);
#pragma unroll
                        for (int r = 0; r < RG; ++r)
                            mma_f16(part[c][r], x[2 * r][0].x, x[2 * r + 1][0].x, x[2 * r][0].y, x[2 * r + 1][0].y, w.z, w.w);
                    }
                } else if constexpr (C8) {
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const uint4 *wp = reinterpret_cast<const uint4 *>(wb + c * 256) + lane;
                        const uint4 wa = wp[0], wl = wp[32];
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16_zero(part[c][r], wa.x, wa.y, wa.z, wa.w, x[r][0].x, x[r][0].y);
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16(part[c][r], wl.x, wl.y, wl.z, wl.w, x[r][0].x, x[r][0].y);
                    }
                } else if constexpr (NT) {
                    uint4 w[CT];
#pragma unroll
                    for (int c = 0; c < CT; ++c) w[c] = *reinterpret_cast<const uint4 *>(wb + ((q * CT + c) * 32 + lane) * 4);
#pragma unroll
                    for (int c = 0; c < CT; ++c)
#pragma unroll
                        for (int r = 0; r < RG; ++r)                                                   // X_lo * W_hi
                            mma_f16(small[c][r], x[2 * r][q].z, x[2 * r + 1][q].z, x[2 * r][q].w, x[2 * r + 1][q].w, w[c].x, w[c].y);
#pragma unroll
                    for (int c = 0; c < CT; ++c)
#pragma unroll
                        for (int r = 0; r < RG; ++r) {                                                 // X_hi * W_hi
                            if (q == 0) mma_f16_zero(part[c][r], x[2 * r][q].x, x[2 * r + 1][q].x, x[2 * r][q].y, x[2 * r + 1][q].y, w[c].x, w[c].y);
                            else mma_f16(part[c][r], x[2 * r][q].x, x[2 * r + 1][q].x, x[2 * r][q].y, x[2 * r + 1][q].y, w[c].x, w[c].y);
                        }
#pragma unroll
                    for (int c = 0; c < CT; ++c)
#pragma unroll
                        for (int r = 0; r < RG; ++r)                                                   // X_hi * W_lo
                            mma_f16(small[c][r], x[2 * r][q].x, x[2 * r + 1][q].x, x[2 * r][q].y, x[2 * r + 1][q].y, w[c].z, w[c].w);
                } else {
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const uint4 *wp = reinterpret_cast<const uint4 *>((ko_wlds ? wsm : wb) + (q * CT + c) * 256) + lane;
                        const uint4 wh = wp[0], wl = wp[32];
                        if constexpr (ko_mma) {
#pragma unroll
                            for (int r = 0; r < RG; ++r) {
                                part[c][r][0] = __uint_as_float(wh.x ^ x[r][q].x); part[c][r][1] = __uint_as_float(wh.y ^ x[r][q].y);
                                part[c][r][2] = __uint_as_float(wl.z ^ x[r][q].z); part[c][r][3] = __uint_as_float(wl.w ^ x[r][q].w);
                            }
                            continue;
                        }
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16(small[c][r], wh.x, wh.y, wh.z, wh.w, x[r][q].z, x[r][q].w);      // W_hi * X_lo
#pragma unroll
                        for (int r = 0; r < RG; ++r) {                                                                     // W_hi * X_hi
                            if (q == 0) mma_f16_zero(part[c][r], wh.x, wh.y, wh.z, wh.w, x[r][q].x, x[r][q].y);
                            else mma_f16(part[c][r], wh.x, wh.y, wh.z, wh.w, x[r][q].x, x[r][q].y);
                        }
#pragma unroll
                        for (int r = 0; r < RG; ++r) mma_f16(small[c][r], wl.x, wl.y, wl.z, wl.w, x[r][q].x, x[r][q].y);      // W_lo * X_hi
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[c][r][e] += part[c][r][e];
        }

        // ---- epilogue (as conv_h2.cuh)
        const int64_t n = n_par * 8, row0 = oct0 * 8;
        if constexpr (NT) {
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

// ---- CIN = 4 (the 4 -> 8 and 4 -> 4 layers of the finest InceptionResNet blocks) ----------------------------------
// An h2 row of four channels is exactly one 16-byte group {hi01 hi23 lo01 lo23}.  All three products of the split
// arithmetic fit ONE m16n8k16 MMA: the sixteen contraction slots carry [hi01 hi23 lo01 lo23 | hi01 hi23 0 0] on the
// feature side and [Whi01 Whi23 Whi01 Whi23 | Wlo01 Wlo23 0 0] on the weight side, i.e. hi*Whi + lo*Whi + hi*Wlo.
// Lane (g, t) supplies slot t (the t-th 32-bit word of its row: one LDS.32, bank-conflict free) and slot t + 4 (the
// same word again for t < 2, zero otherwise).  One accumulator chain per kernel z-plane (9 MMAs), joined by RN FADDs.
template <int COUT, int RG_, int WARPS_>
struct OctetH2C4Cfg {
    static_assert(COUT == 4 || COUT == 8 || COUT == 16, "octet h2 CIN=4 kernel: COUT in {4, 8, 16}");
    static constexpr int CT = (COUT + 7) / 8;
    static constexpr int RG = RG_, WARPS = WARPS_, THREADS = 32 * WARPS_;
    static constexpr int OW = 2 * RG;                              // octets per warp iteration (16 rows per group)
    static constexpr int SY = 4, SZ = 18;                          // rows {0,1,4,5,18,19,22,23} x 4 words: 32 distinct banks
    static constexpr int HROWS = 3 * SZ + 3 * SY + 4;
    static constexpr int HB = HROWS * 16;
    static constexpr int W_OFF = CT * 64;                          // packed words per offset: [CT][32 lanes][2]
    static constexpr size_t packed_words() { return (size_t)27 * W_OFF; }
    static constexpr size_t weight_bytes() { return (packed_words() * 4 + 127) / 128 * 128; }
    static constexpr size_t warp_bytes() { return ((size_t)OW * HB + (size_t)27 * OW * 4 + 127) / 128 * 128; }
    static constexpr size_t smem_bytes() { return weight_bytes() + (size_t)WARPS * warp_bytes(); }
    static constexpr int OCTETS_PER_CTA = WARPS * OW;
};

// W [27][4][cout] -> [27][CT][32 lanes] x {b0, b1}: b0 = slot t: (t & 1 ? Whi[2,3] : Whi[0,1])[8c+g], b1 = slot t+4: Wlo (t < 2) or 0
static __global__ void pack_weights_h2c4_kernel(const float *__restrict__ w, int cout, float scale, uint32_t *__restrict__ packed) {
    const int CT = (cout + 7) / 8;
    const int total = 27 * CT * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int lane = i & 31, c = (i >> 5) % CT, k = (i >> 5) / CT;
        const int g = lane >> 2, t = lane & 3, co = 8 * c + g, ci = 2 * (t & 1);
        const float x0 = co < cout ? w[(k * 4 + ci) * cout + co] * scale : 0.f;
        const float x1 = co < cout ? w[(k * 4 + ci + 1) * cout + co] * scale : 0.f;
        uint32_t hi, lo;
        split_pair_h2(x0, x1, hi, lo);
        packed[2 * i] = hi;
        packed[2 * i + 1] = t < 2 ? lo : 0u;
    }
}

// K1TAIL (COUT = 4 only): the k=1 convolution that follows the layer in the InceptionResNet block (conv1_2 after conv1_1,
// autoencoder.py:55) runs in the epilogue: y = relu(conv_k3(x)) is never written; out = y W1 + b1 + residual with W1
// [4][8] fp32 (`tail_w`), b1 [8] (`tail_b`).  A row's four values sit in lanes t = 0, 1 of its quad: four shuffles hand them
// to all four lanes, lane t finishes output channels 2t, 2t+1 (8 FFMA) and stores the fp32 and the h2 pair.
template <int COUT, int RG, int WARPS, int MINB, bool K1TAIL = false>
__global__ void __launch_bounds__(32 * WARPS, MINB)
conv_k3_octet_h2c4_kernel(const uint32_t *__restrict__ in, int in_ld, const int32_t *__restrict__ pnbr, int64_t n_par,
                          const uint32_t *__restrict__ packed, float inv_scale, const float *__restrict__ bias,
                          const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld,
                          uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow,
                          const float *__restrict__ tail_w = nullptr, const float *__restrict__ tail_b = nullptr) {
    static_assert(!K1TAIL || COUT == 4, "k=1 tail: 4 -> 8 only");
    using C = OctetH2C4Cfg<COUT, RG, WARPS>;
    constexpr int CT = C::CT, OW = C::OW, SY = C::SY, SZ = C::SZ, HB = C::HB, W_OFF = C::W_OFF;
    extern __shared__ __align__(128) unsigned char smem_oh2[];
    uint32_t *wsm = reinterpret_cast<uint32_t *>(smem_oh2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    unsigned char *halo = smem_oh2 + C::weight_bytes() + (size_t)warp * C::warp_bytes();
    int32_t *sidx = reinterpret_cast<int32_t *>(halo + (size_t)OW * HB);

    for (int i = threadIdx.x; i < (int)C::packed_words(); i += C::THREADS) wsm[i] = __ldg(packed + i);
    __syncthreads();

    const int cx = g & 1, cy = (g >> 1) & 1, cz = g >> 2;
    const unsigned char *base = halo + (cx + SY * cy + SZ * cz) * 16 + 4 * t;       // child g, word t of the row
    const bool dup = t < 2;                                                         // slots 4, 5 repeat the hi words

    const int64_t n_tiles = (n_par + C::OCTETS_PER_CTA - 1) / C::OCTETS_PER_CTA;
    const char *in_bytes = reinterpret_cast<const char *>(in);
    const uint32_t ldb = (uint32_t)in_ld * 4u;
    H2Epilogue epi{K1TAIL ? tail_b : bias, residual, out, out_h2, res_ld, out_ld, out_h2_ld, K1TAIL ? 0 : flags, K1TAIL ? 1.f : inv_scale};
    float tw[4][2], b0 = 0.f, b1 = 0.f;                             // K1TAIL: W1[i][2t], W1[i][2t+1]; first-stage bias of this lane's pair
    if constexpr (K1TAIL) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { tw[i][0] = __ldg(tail_w + i * 8 + 2 * t); tw[i][1] = __ldg(tail_w + i * 8 + 2 * t + 1); }
        if (t < 2 && bias) { b0 = __ldg(bias + 2 * t); b1 = __ldg(bias + 2 * t + 1); }
    }
    int32_t prow[OW];
    const bool chunked = flags & PCGC_TILES_CHUNKED;                // see conv_k3_octet_h2_kernel
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t tile_begin = chunked ? blockIdx.x * per_cta : blockIdx.x, tile_step = chunked ? 1 : gridDim.x;
    const int64_t tile_end = chunked ? min(n_tiles, tile_begin + per_cta) : n_tiles;
    load_parent_rows<OW>(prow, pnbr, n_par, tile_begin * C::OCTETS_PER_CTA + warp * OW, lane);
    for (int64_t tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int64_t oct0 = tile * C::OCTETS_PER_CTA + warp * OW;
        __syncwarp();
        store_parent_rows<OW>(sidx, prow, lane);
        __syncwarp();
        halo_fill<1, OW, HB, 16, SY, SZ, false>(halo, sidx, in_bytes, ldb, lane);
        load_parent_rows<OW>(prow, pnbr, n_par, (tile + tile_step) * C::OCTETS_PER_CTA + warp * OW, lane);

        float acc[CT][RG][4];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[c][r][e] = 0.f;

#pragma unroll
        for (int iz = 0; iz < 3; ++iz) {
            halo_wait(iz);
            float plane[CT][RG][4];
#pragma unroll
            for (int oo = 0; oo < 9; ++oo) {
                const int o = 9 * iz + oo, ix = oo % 3, iy = oo / 3;
                const int doff = (ix + SY * iy + SZ * iz) * 16;
                uint2 w[CT];
#pragma unroll
                for (int c = 0; c < CT; ++c) w[c] = *reinterpret_cast<const uint2 *>(wsm + o * W_OFF + (c * 32 + lane) * 2);
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const uint32_t a0 = *reinterpret_cast<const uint32_t *>(base + (2 * r) * HB + doff);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t *>(base + (2 * r + 1) * HB + doff);
                    const uint32_t a2 = dup ? a0 : 0u, a3 = dup ? a1 : 0u;
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        if (oo == 0) mma_f16_zero(plane[c][r], a0, a1, a2, a3, w[c].x, w[c].y);
                        else mma_f16(plane[c][r], a0, a1, a2, a3, w[c].x, w[c].y);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[c][r][e] += plane[c][r][e];
        }

        const int64_t n = n_par * 8, row0 = oct0 * 8;
        if constexpr (K1TAIL) {
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // first stage on the lanes that own channels (t = 0: 0, 1; t = 1: 2, 3): scale, bias, ReLU
                    const float v0 = fmaxf(acc[0][r][2 * h] * inv_scale + b0, 0.f), v1 = fmaxf(acc[0][r][2 * h + 1] * inv_scale + b1, 0.f);
                    const int q0 = lane & ~3;
                    const float c0 = __shfl_sync(0xffffffffu, v0, q0), c1 = __shfl_sync(0xffffffffu, v1, q0);
                    const float c2 = __shfl_sync(0xffffffffu, v0, q0 + 1), c3 = __shfl_sync(0xffffffffu, v1, q0 + 1);
                    const int64_t row = row0 + 16 * r + 8 * h + g;
                    if (row >= n) continue;
                    float y0 = c0 * tw[0][0], y1 = c0 * tw[0][1];
                    y0 = fmaf(c1, tw[1][0], y0); y1 = fmaf(c1, tw[1][1], y1);
                    y0 = fmaf(c2, tw[2][0], y0); y1 = fmaf(c2, tw[2][1], y1);
                    y0 = fmaf(c3, tw[3][0], y0); y1 = fmaf(c3, tw[3][1], y1);
                    epi.store_pair(row, 2 * t, y0, y1);               // + b1 + residual, fp32 and h2 stores
                }
        } else {
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
                        epi.store_pair(row, co, acc[c][r][2 * h], acc[c][r][2 * h + 1]);
                    }
            }
        }
    }
    if (epi.over && overflow) *overflow = 1;
}

// ---- both second-stage layers of a 16-channel InceptionResNet block in one kernel --------------------------------------
// conv0_1 (k=3, 4 -> 8, + x[:, :8]) reads a, conv1_1 (k=3, 4 -> 4, ReLU) -> conv1_2 (k=1, 4 -> 8, + x[:, 8:]) reads b, and a | b are
// the two 16-byte halves of the rows the merged first layer wrote (PCGC_IRN_MERGED_FIRST).  One pass stages every 32-byte row
// ONCE (one sector; the two single-branch kernels each fetched the same sector for their half) into two CIN = 4 halo planes
// and runs both MMA chains on them; the block's 16 output channels leave from one epilogue.  Arithmetic per branch is that of
// conv_k3_octet_h2c4_kernel (<8> and <4, K1TAIL>), bit for bit.
template <int RG_, int WARPS_>
struct OctetH2C4DualCfg {
    using A = OctetH2C4Cfg<8, RG_, WARPS_>;
    static constexpr int RG = RG_, WARPS = WARPS_, THREADS = 32 * WARPS_, OW = 2 * RG_;
    static constexpr int SY = A::SY, SZ = A::SZ, HB = A::HB, W_OFF = 64;
    static constexpr int PLANE = OW * HB;                          // bytes of one branch's halos of a warp
    static constexpr size_t weight_bytes() { return ((size_t)2 * 27 * W_OFF * 4 + 127) / 128 * 128; }
    static constexpr size_t warp_bytes() { return ((size_t)2 * PLANE + (size_t)27 * OW * 4 + 127) / 128 * 128; }
    static constexpr size_t smem_bytes() { return weight_bytes() + (size_t)WARPS * warp_bytes(); }
    static constexpr int OCTETS_PER_CTA = WARPS * OW;
};

template <int RG, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
conv_k3_octet_h2c4_dual_kernel(const uint32_t *__restrict__ in, int in_ld, const int32_t *__restrict__ pnbr, int64_t n_par,
                               const uint32_t *__restrict__ packed_a, float inv_scale_a, const float *__restrict__ bias_a,
                               const uint32_t *__restrict__ packed_b, float inv_scale_b, const float *__restrict__ bias_b,
                               const float *__restrict__ tail_w, const float *__restrict__ tail_b,
                               const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld,
                               uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    using C = OctetH2C4DualCfg<RG, WARPS>;
    constexpr int OW = C::OW, SY = C::SY, SZ = C::SZ, HB = C::HB, W_OFF = C::W_OFF, PLANE = C::PLANE;
    extern __shared__ __align__(128) unsigned char smem_oh2[];
    uint32_t *wsm_a = reinterpret_cast<uint32_t *>(smem_oh2), *wsm_b = wsm_a + 27 * W_OFF;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    unsigned char *halo = smem_oh2 + C::weight_bytes() + (size_t)warp * C::warp_bytes();
    int32_t *sidx = reinterpret_cast<int32_t *>(halo + (size_t)2 * PLANE);

    for (int i = threadIdx.x; i < 27 * W_OFF; i += C::THREADS) { wsm_a[i] = __ldg(packed_a + i); wsm_b[i] = __ldg(packed_b + i); }
    __syncthreads();

    const int cx = g & 1, cy = (g >> 1) & 1, cz = g >> 2;
    const unsigned char *base = halo + (cx + SY * cy + SZ * cz) * 16 + 4 * t;       // child g, word t of the a row; b: + PLANE
    const bool dup = t < 2;

    const int64_t n_tiles = (n_par + C::OCTETS_PER_CTA - 1) / C::OCTETS_PER_CTA;
    const char *in_bytes = reinterpret_cast<const char *>(in);
    const uint32_t ldb = (uint32_t)in_ld * 4u;
    // output channels 0..7 (branch a) and 8..15 (branch b: the k=1 tail) of the block's rows
    H2Epilogue epi_a{bias_a, residual, out, out_h2, res_ld, out_ld, out_h2_ld, 0, inv_scale_a};
    H2Epilogue epi_b{tail_b, residual ? residual + 8 : nullptr, out ? out + 8 : nullptr, out_h2 ? out_h2 + 8 : nullptr, res_ld, out_ld,
                     out_h2_ld, 0, 1.f};
    float tw[4][2], b0 = 0.f, b1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { tw[i][0] = __ldg(tail_w + i * 8 + 2 * t); tw[i][1] = __ldg(tail_w + i * 8 + 2 * t + 1); }
    if (t < 2 && bias_b) { b0 = __ldg(bias_b + 2 * t); b1 = __ldg(bias_b + 2 * t + 1); }

    int32_t prow[OW];
    const bool chunked = flags & PCGC_TILES_CHUNKED;
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t tile_begin = chunked ? blockIdx.x * per_cta : blockIdx.x, tile_step = chunked ? 1 : gridDim.x;
    const int64_t tile_end = chunked ? min(n_tiles, tile_begin + per_cta) : n_tiles;
    load_parent_rows<OW>(prow, pnbr, n_par, tile_begin * C::OCTETS_PER_CTA + warp * OW, lane);
    for (int64_t tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int64_t oct0 = tile * C::OCTETS_PER_CTA + warp * OW;
        __syncwarp();
        store_parent_rows<OW>(sidx, prow, lane);
        __syncwarp();
        halo_fill<2, OW, HB, 16, SY, SZ, false, PLANE>(halo, sidx, in_bytes, ldb, lane);   // piece 0 -> plane a, piece 1 -> plane b
        load_parent_rows<OW>(prow, pnbr, n_par, (tile + tile_step) * C::OCTETS_PER_CTA + warp * OW, lane);

        float acc_a[RG][4], acc_b[RG][4];
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc_a[r][e] = acc_b[r][e] = 0.f;

#pragma unroll
        for (int iz = 0; iz < 3; ++iz) {
            halo_wait(iz);
            float pl_a[RG][4], pl_b[RG][4];
#pragma unroll
            for (int oo = 0; oo < 9; ++oo) {
                const int o = 9 * iz + oo, ix = oo % 3, iy = oo / 3;
                const int doff = (ix + SY * iy + SZ * iz) * 16;
                const uint2 wa = *reinterpret_cast<const uint2 *>(wsm_a + o * W_OFF + lane * 2);
                const uint2 wb = *reinterpret_cast<const uint2 *>(wsm_b + o * W_OFF + lane * 2);
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const uint32_t a0 = *reinterpret_cast<const uint32_t *>(base + (2 * r) * HB + doff);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t *>(base + (2 * r + 1) * HB + doff);
                    const uint32_t c0 = *reinterpret_cast<const uint32_t *>(base + PLANE + (2 * r) * HB + doff);
                    const uint32_t c1 = *reinterpret_cast<const uint32_t *>(base + PLANE + (2 * r + 1) * HB + doff);
                    if (oo == 0) {
                        mma_f16_zero(pl_a[r], a0, a1, dup ? a0 : 0u, dup ? a1 : 0u, wa.x, wa.y);
                        mma_f16_zero(pl_b[r], c0, c1, dup ? c0 : 0u, dup ? c1 : 0u, wb.x, wb.y);
                    } else {
                        mma_f16(pl_a[r], a0, a1, dup ? a0 : 0u, dup ? a1 : 0u, wa.x, wa.y);
                        mma_f16(pl_b[r], c0, c1, dup ? c0 : 0u, dup ? c1 : 0u, wb.x, wb.y);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) { acc_a[r][e] += pl_a[r][e]; acc_b[r][e] += pl_b[r][e]; }
        }

        const int64_t n = n_par * 8, row0 = oct0 * 8;
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float v0 = fmaxf(acc_b[r][2 * h] * inv_scale_b + b0, 0.f), v1 = fmaxf(acc_b[r][2 * h + 1] * inv_scale_b + b1, 0.f);
                const int q0 = lane & ~3;
                const float c0 = __shfl_sync(0xffffffffu, v0, q0), c1 = __shfl_sync(0xffffffffu, v1, q0);
                const float c2 = __shfl_sync(0xffffffffu, v0, q0 + 1), c3 = __shfl_sync(0xffffffffu, v1, q0 + 1);
                const int64_t row = row0 + 16 * r + 8 * h + g;
                if (row >= n) continue;
                epi_a.store_pair(row, 2 * t, acc_a[r][2 * h], acc_a[r][2 * h + 1]);          // conv0_1 + bias + x[:, :8]
                float y0 = c0 * tw[0][0], y1 = c0 * tw[0][1];
                y0 = fmaf(c1, tw[1][0], y0); y1 = fmaf(c1, tw[1][1], y1);
                y0 = fmaf(c2, tw[2][0], y0); y1 = fmaf(c2, tw[2][1], y1);
                y0 = fmaf(c3, tw[3][0], y0); y1 = fmaf(c3, tw[3][1], y1);
                epi_b.store_pair(row, 2 * t, y0, y1);                                        // conv1_2 + bias + x[:, 8:]
            }
    }
    if ((epi_a.over || epi_b.over) && overflow) *overflow = 1;
}

}  // namespace pcgc
