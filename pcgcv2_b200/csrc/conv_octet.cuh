// conv_octet.cuh -- k=3 sparse convolution on FULL-OCTET coordinate sets (round-1 v6).
//
// Every set the synthesis network convolves on is the 8-child expansion of a parent set
// (ME.MinkowskiGenerativeConvolutionTranspose, autoencoder.py:155,182,209): n = 8 P rows, row
// 8 i + c is child c = cx + 2 cy + 4 cz of parent row i (Morton order keeps the 8 siblings of an
// octet contiguous: 8 * CIN * 4 bytes).  The 27 neighbours of the 8 children of one octet all lie
// in the 4 x 4 x 4 voxel "halo" around it -- 64 rows instead of 8 x 27 = 216 gathers -- and that
// halo is addressed by the PARENT's kernel map alone (27 entries per octet = 13.5 B per row
// instead of 108 B per row; the child-level map of the 1.69 M-row level is never built).
//
// A warp stages the halos of its octets in shared memory with cp.async (16-byte pieces, zero
// fill for missing parent neighbours; sibling rows are fetched as contiguous 32..512-byte runs:
// ~48 cache lines per octet instead of ~175 for per-row gathers, measured L1 tag pressure was the
// gather limit), then walks the 27 offsets with every operand address a compile-time offset
// from one lane base: no index loads, no 64-bit address arithmetic, no predicates and no
// long-scoreboard stalls in the math loop.
//
//  * conv_k3_octet_mma_kernel (CIN 8/16/32): 3xTF32 mma.sync.m16n8k8, D[row, cout] += X_k W_k with
//    16 rows (two octets) as the A operand read straight from the halo (the permuted contraction
//    index of conv_mma.cuh makes a lane's 16-byte LDS its A fragments for two k-steps) and the
//    pre-split weights as B fragments (one LDS.128 per k-step and 8-wide output tile).
//    Arithmetic identical to conv_pipe.cuh (main term joined by an RN FADD per offset).
//  * conv_k3_octet_ffma_kernel (CIN 4): one lane per output row, FP32 FFMA in ascending
//    (offset, channel) order -- bit-identical to conv_rowlane.cuh.
//
// Shared-memory bank layout: halo row index = hx + SY*hy + SZ*hz with SY, SZ chosen per row width so
// that the 8 (or 16) lanes of one LDS phase hit distinct banks; rows of >= 128 bytes XOR the 64-byte
// chunk index with the x parity.
#pragma once
#include "common.cuh"
#include "conv_mma.cuh"
#include "conv_pipe.cuh"

namespace pcgc {

// halo rows are re-read by the neighbouring octets of the same CTA: allocate them in L1 (.ca)
__device__ __forceinline__ void cp_async16_ca(void *smem_dst, const void *gmem_src, bool pred) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int bytes = pred ? 16 : 0;            // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}

// parent offset index (0..2) and child bit of halo coordinate h in 0..3 (voxel offset h - 1 from the octet origin)
__host__ __device__ __forceinline__ constexpr int halo_parent(int h) { return (h + 1) >> 1; }
__host__ __device__ __forceinline__ constexpr int halo_child(int h) { return (h + 1) & 1; }


// ---- halo staging shared by both kernels ------------------------------------------------------------
// One warp fills the halos of its OW octets, one z-plane (16 positions per octet) per call so that each plane
// is its own cp.async group: the math on offsets iz = 0 starts when planes 0 and 1 have landed.  A plane is
// OW * 16 positions x PPR 16-byte pieces, 32 pieces per warp instruction.  Piece e = 32*it + lane: the
// piece-in-row j = lane % PPR and the low bits of (octet, y, x) come from the lane, everything else from the
// (compile-time, fully unrolled) iteration -- so the parent neighbour index q, the child bits cc and both
// addresses are "lane constant + immediate" and a piece costs LDS (parent row) + max + compare + one
// 32x32->64 multiply-add + LDGSTS.
// PIECE_STRIDE: byte distance between the destinations of consecutive 16-byte pieces of a row (16 = the row is stored whole;
// a plane size = every piece goes to its own copy of the halo layout: the two 4-channel halves of an 8-channel row feed two
// independent CIN = 4 halos, conv_k3_octet_h2c4_dual_kernel)
template <int PPR, int OW, int HB, int ROWB, int SY, int SZ, bool SWZ, int PIECE_STRIDE = 16>
__device__ __forceinline__ void halo_fill_plane(int hz, unsigned char *halo, const int32_t *sidx, const char *in_bytes,
                                                uint32_t ldb, int lane) {
    constexpr int LPI = 32 / PPR;                                   // halo positions per warp instruction
    static_assert(LPI >= 4 && LPI <= 32 && (OW * 16) % LPI == 0, "halo_fill: 1..8 pieces per row");
    const int j = lane % PPR, u = lane / PPR;
    const uint32_t stride8 = 8u * ldb;                              // bytes between the first children of consecutive parents
#pragma unroll
    for (int it = 0; it < OW * 16 / LPI; ++it) {
        const int c = it * LPI;                                     // compile-time after unrolling; bits disjoint from u
        const int hx = u & 3;
        const int hy = ((u >> 2) & 3) | ((c >> 2) & 3);
        const int o = (u >> 4) | (c >> 4);
        const int q = halo_parent(hx) + 3 * halo_parent(hy) + 9 * halo_parent(hz);
        const int cc = halo_child(hx) | (halo_child(hy) << 1) | (halo_child(hz) << 2);
        const int32_t p = sidx[q * OW + o];
        const int piece = SWZ ? ((((j >> 2) ^ (hx & 1)) << 2) | (j & 3)) : j;
        unsigned char *dst = halo + o * HB + (hx + SY * hy + SZ * hz) * ROWB + piece * PIECE_STRIDE;
        const char *src = in_bytes + (size_t)cc * ldb + 16 * j + (uint64_t)(uint32_t)max(p, 0) * stride8;
        cp_async16_ca(dst, src, p >= 0);
    }
}
template <int PPR, int OW, int HB, int ROWB, int SY, int SZ, bool SWZ, int PIECE_STRIDE = 16>
__device__ __forceinline__ void halo_fill(unsigned char *halo, const int32_t *sidx, const char *in_bytes, uint32_t ldb,
                                          int lane) {
#pragma unroll
    for (int hz = 0; hz < 4; ++hz) {
        halo_fill_plane<PPR, OW, HB, ROWB, SY, SZ, SWZ, PIECE_STRIDE>(hz, halo, sidx, in_bytes, ldb, lane);
        cp_async_commit();                                          // group hz
    }
}
// wait until the z-planes that kernel-offset plane iz (0..2) reads (iz and iz + 1) have landed
__device__ __forceinline__ void halo_wait(int iz) {
    if (iz == 0) cp_async_wait<2>();
    else if (iz == 1) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncwarp();
}

// lane k < 27 fetches the parent rows of neighbour k of the warp's OW octets (-1: missing / past the end)
template <int OW>
__device__ __forceinline__ void load_parent_rows(int32_t (&v)[OW], const int32_t *__restrict__ pnbr, int64_t n_par,
                                                 int64_t oct0, int lane) {
#pragma unroll
    for (int o = 0; o < OW; ++o) {
        v[o] = -1;
        if (lane < 27 && oct0 + o < n_par) v[o] = lane == 13 ? (int32_t)(oct0 + o) : __ldg(pnbr + (int64_t)lane * n_par + oct0 + o);
    }
}
template <int OW>
__device__ __forceinline__ void store_parent_rows(int32_t *sidx, const int32_t (&v)[OW], int lane) {
    if (lane < 27) {
#pragma unroll
        for (int o = 0; o < OW; ++o) sidx[lane * OW + o] = v[o];
    }
}

template <int CIN, int COUT, int RG_, int WARPS_>
struct OctetMmaCfg {
    static_assert(CIN == 8 || CIN == 16 || CIN == 32 || CIN == 64, "octet mma kernel: CIN in {8,16,32,64}");
    static constexpr int KS = CIN / 8;                            // k-steps per offset
    static constexpr int CT = (COUT + 7) / 8;                     // 8-wide output tiles
    static constexpr int CHUNKS = CIN >= 16 ? CIN / 16 : 1;       // 64-byte chunks per row (loads per row per lane)
    static constexpr int AV = CIN >= 16 ? 4 : 2;                  // floats per lane load
    static constexpr int RG = RG_, WARPS = WARPS_, THREADS = 32 * WARPS_;
    static constexpr int OW = 2 * RG;                             // octets per warp iteration
    static constexpr int ROWB = CIN * 4;                          // halo row bytes
    static constexpr int PPR = CIN / 4;                           // 16-byte pieces per row
    static constexpr int SY = CIN == 8 ? 6 : 4, SZ = 4 * SY;      // halo strides in rows (CIN 8: y rows 16 banks apart)
    static constexpr int HROWS = 3 * SZ + 3 * SY + 4;
    static constexpr int HB = HROWS * ROWB;                       // halo bytes per octet
    static constexpr int W_OFF = KS * CT * 128;                   // packed floats per offset (NT layout, hi + lo)
    static constexpr size_t packed_floats() { return (size_t)27 * W_OFF; }
    static constexpr size_t warp_bytes() { return ((size_t)OW * HB + (size_t)27 * OW * 4 + 127) / 128 * 128; }
    static constexpr size_t smem_bytes() { return packed_floats() * 4 + (size_t)WARPS * warp_bytes(); }
    static constexpr int OCTETS_PER_CTA = WARPS * OW;
};

template <int CIN, int COUT, int RG, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
conv_k3_octet_mma_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ pnbr, int64_t n_par,
                         const float *__restrict__ packed, const float *__restrict__ bias,
                         const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags,
                         uint32_t hi_mask) {
    // hi_mask = 0xFFFFE000 arrives as a kernel argument on purpose: with a literal mask ptxas folds the AND into
    // the TF32 operand (the tensor core ignores those bits anyway) and then has to assemble every A-fragment quad
    // from the raw registers of two different loads with MOVs (65 moves per 24 HMMAs, measured in SASS); an
    // opaque mask makes the hi parts computed values that are allocated straight into the aligned quad.
    using C = OctetMmaCfg<CIN, COUT, RG, WARPS>;
    constexpr int KS = C::KS, CT = C::CT, CHUNKS = C::CHUNKS, AV = C::AV, OW = C::OW, ROWB = C::ROWB, PPR = C::PPR;
    constexpr int SY = C::SY, SZ = C::SZ, HB = C::HB, W_OFF = C::W_OFF, NR = 2 * RG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *wsm = reinterpret_cast<float *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    unsigned char *halo = smem_raw + C::packed_floats() * 4 + (size_t)warp * C::warp_bytes();
    int32_t *sidx = reinterpret_cast<int32_t *>(halo + (size_t)OW * HB);            // [27][OW] parent rows of the neighbours

    for (int i = threadIdx.x; i < 27 * W_OFF / 4; i += C::THREADS) cp_async16(wsm + 4 * i, packed + 4 * i, true);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // lane's read base: child g of an octet, 16-byte piece t (8-byte piece for CIN 8) of the row
    const int cx = g & 1, cy = (g >> 1) & 1, cz = g >> 2;
    const unsigned char *slot[CHUNKS];                              // physical chunk (s ^ cx) of the lane's row
#pragma unroll
    for (int s = 0; s < CHUNKS; ++s)
        slot[s] = halo + (cx + SY * cy + SZ * cz) * ROWB + (CHUNKS >= 2 ? ((s ^ cx) * 64) : 0) + t * (AV * 4);

    const int64_t n_tiles = (n_par + C::OCTETS_PER_CTA - 1) / C::OCTETS_PER_CTA;
    const char *in_bytes = reinterpret_cast<const char *>(in);
    const uint32_t ldb = (uint32_t)in_ld * 4u;
    int32_t prow[OW];                                               // parent rows for the NEXT tile (lane k: neighbour k)
    load_parent_rows<OW>(prow, pnbr, n_par, (int64_t)blockIdx.x * C::OCTETS_PER_CTA + warp * OW, lane);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t oct0 = tile * C::OCTETS_PER_CTA + warp * OW;                // first octet of this warp
        __syncwarp();                                                             // previous iteration's readers are done
        store_parent_rows<OW>(sidx, prow, lane);
        __syncwarp();
        halo_fill<PPR, OW, HB, ROWB, SY, SZ, (CHUNKS >= 2)>(halo, sidx, in_bytes, ldb, lane);
        // the next tile's parent rows travel while this tile is multiplied
        load_parent_rows<OW>(prow, pnbr, n_par, (tile + gridDim.x) * C::OCTETS_PER_CTA + warp * OW, lane);

        float acc[CT][RG][4], small[CT][RG][4];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[c][r][e] = small[c][r][e] = 0.f;
        halo_wait(0);

        // ---- 27 offsets, fully unrolled: every address is lane base + immediate; the fragments of offset
        //      o + 1 are read (register double buffer) before the MMAs of offset o are issued
        float xb[2][NR][CHUNKS][AV];
        auto load_frags = [&](int o, float (&x)[NR][CHUNKS][AV]) {
            const int ix = o % 3, iy = (o / 3) % 3, iz = o / 9;
            const int doff = (ix + SY * iy + SZ * iz) * ROWB;
#pragma unroll
            for (int j = 0; j < NR; ++j)
#pragma unroll
                for (int q = 0; q < CHUNKS; ++q) {
                    const unsigned char *a = slot[CHUNKS >= 2 ? (q ^ (ix & 1)) : 0] + j * HB + doff;
                    if constexpr (AV == 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(a);
                        x[j][q][0] = v.x; x[j][q][1] = v.y; x[j][q][2] = v.z; x[j][q][3] = v.w;
                    } else {
                        const float2 v = *reinterpret_cast<const float2 *>(a);
                        x[j][q][0] = v.x; x[j][q][1] = v.y;
                    }
                }
        };
        load_frags(0, xb[0]);
#pragma unroll
        for (int o = 0; o < 27; ++o) {
            if ((o + 1) % 9 == 0 && o + 1 < 27) halo_wait((o + 1) / 9);           // next z-plane of the halo
            if (o + 1 < 27) load_frags(o + 1, xb[(o + 1) & 1]);
            float (&x)[NR][CHUNKS][AV] = xb[o & 1];
            const float *wb = wsm + (size_t)o * W_OFF;
            float part[CT][RG][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int q = CIN >= 16 ? ks >> 1 : 0, s = CIN >= 16 ? 2 * (ks & 1) : 0;
                float4 w[CT];
#pragma unroll
                for (int c = 0; c < CT; ++c) w[c] = *reinterpret_cast<const float4 *>(wb + ((ks * CT + c) * 32 + lane) * 4);
                float h[RG][4], l[RG][4];
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const float a[4] = {x[2 * r][q][s], x[2 * r + 1][q][s], x[2 * r][q][s + 1], x[2 * r + 1][q][s + 1]};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { h[r][e] = __uint_as_float(__float_as_uint(a[e]) & hi_mask); l[r][e] = a[e] - h[r][e]; }
                }
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
            }
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[c][r][e] += part[c][r][e];
        }

        // ---- epilogue: fragment (c, r): e=0 -> (row 16r+g, cout 8c+2t), e=1 -> (.., 8c+2t+1), e=2/3 -> row 16r+g+8
        const int64_t n = n_par * 8, row0 = oct0 * 8;
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
                for (int hh = 0; hh < 2; ++hh) {
                    const int64_t row = row0 + 16 * r + 8 * hh + g;
                    if (row >= n) continue;
                    float v0 = acc[c][r][2 * hh] + small[c][r][2 * hh] + b0;
                    float v1 = acc[c][r][2 * hh + 1] + small[c][r][2 * hh + 1] + b1;
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
    }
}

// ---- CIN = 4: FP32 FFMA, one lane per output row ----------------------------------------------------
template <int COUT, int RPL_, int WARPS_>
struct OctetFfmaCfg {
    static_assert(COUT % 4 == 0 && COUT <= 16, "octet ffma kernel: COUT in {4, 8, 12, 16}");
    static constexpr int CIN = 4, RPL = RPL_, WARPS = WARPS_, THREADS = 32 * WARPS_;
    static constexpr int OW = 4 * RPL;                             // octets per warp iteration (32 * RPL rows)
    static constexpr int SY = 4, SZ = 18;                          // z rows 8 banks apart from the y rows
    static constexpr int HROWS = 3 * SZ + 3 * SY + 4;
    static constexpr int HB = HROWS * 16;
    static constexpr size_t weight_floats() { return (size_t)27 * 4 * COUT; }
    static constexpr size_t weight_bytes() { return (weight_floats() * 4 + 127) / 128 * 128; }
    static constexpr size_t warp_bytes() { return ((size_t)OW * HB + (size_t)27 * OW * 4 + 127) / 128 * 128; }
    static constexpr size_t smem_bytes() { return weight_bytes() + (size_t)WARPS * warp_bytes(); }
    static constexpr int OCTETS_PER_CTA = WARPS * OW;
};

template <int COUT, int RPL, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
conv_k3_octet_ffma_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ pnbr, int64_t n_par,
                          const float *__restrict__ weight, const float *__restrict__ bias,
                          const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using C = OctetFfmaCfg<COUT, RPL, WARPS>;
    constexpr int OW = C::OW, SY = C::SY, SZ = C::SZ, HB = C::HB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *wsm = reinterpret_cast<float *>(smem_raw);              // [27][4][COUT], the reference layout
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *halo = smem_raw + C::weight_bytes() + (size_t)warp * C::warp_bytes();
    int32_t *sidx = reinterpret_cast<int32_t *>(halo + (size_t)OW * HB);

    for (int i = threadIdx.x; i < (int)C::weight_floats(); i += C::THREADS) wsm[i] = __ldg(weight + i);
    __syncthreads();

    const int child = lane & 7, lo = lane >> 3;                    // row = octet (4*grp + lo), child
    const int cx = child & 1, cy = (child >> 1) & 1, cz = child >> 2;
    const unsigned char *base = halo + lo * HB + (cx + SY * cy + SZ * cz) * 16;
    const bool io_vec = (((uintptr_t)out & 15) == 0) && (out_ld % 4 == 0) &&
                        (!residual || ((((uintptr_t)residual & 15) == 0) && (res_ld % 4 == 0)));

    const int64_t n_tiles = (n_par + C::OCTETS_PER_CTA - 1) / C::OCTETS_PER_CTA;
    const char *in_bytes = reinterpret_cast<const char *>(in);
    const uint32_t ldb = (uint32_t)in_ld * 4u;
    int32_t prow[OW];
    load_parent_rows<OW>(prow, pnbr, n_par, (int64_t)blockIdx.x * C::OCTETS_PER_CTA + warp * OW, lane);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t oct0 = tile * C::OCTETS_PER_CTA + warp * OW;
        __syncwarp();
        store_parent_rows<OW>(sidx, prow, lane);
        __syncwarp();
        halo_fill<1, OW, HB, 16, SY, SZ, false>(halo, sidx, in_bytes, ldb, lane);
        load_parent_rows<OW>(prow, pnbr, n_par, (tile + gridDim.x) * C::OCTETS_PER_CTA + warp * OW, lane);
        float acc[RPL][COUT];
#pragma unroll
        for (int r = 0; r < RPL; ++r)
#pragma unroll
            for (int i = 0; i < COUT; ++i) acc[r][i] = 0.f;

#pragma unroll
        for (int o = 0; o < 27; ++o) {
            const int ix = o % 3, iy = (o / 3) % 3, iz = o / 9;
            if (o % 9 == 0) halo_wait(iz);
            const int doff = (ix + SY * iy + SZ * iz) * 16;
            float v[RPL][4];
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const float4 t4 = *reinterpret_cast<const float4 *>(base + r * 4 * HB + doff);
                v[r][0] = t4.x; v[r][1] = t4.y; v[r][2] = t4.z; v[r][3] = t4.w;
            }
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                for (int qv = 0; qv < COUT / 4; ++qv) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(wsm + (o * 4 + ci) * COUT + 4 * qv);
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        acc[r][4 * qv + 0] = fmaf(v[r][ci], w4.x, acc[r][4 * qv + 0]);
                        acc[r][4 * qv + 1] = fmaf(v[r][ci], w4.y, acc[r][4 * qv + 1]);
                        acc[r][4 * qv + 2] = fmaf(v[r][ci], w4.z, acc[r][4 * qv + 2]);
                        acc[r][4 * qv + 3] = fmaf(v[r][ci], w4.w, acc[r][4 * qv + 3]);
                    }
                }
        }

#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int64_t oct = oct0 + 4 * r + lo;
            if (oct >= n_par) continue;
            const int64_t row = oct * 8 + child;
            float *o_ptr = out + row * out_ld;
            const float *r_ptr = residual ? residual + row * res_ld : nullptr;
#pragma unroll
            for (int i = 0; i < COUT; ++i) {
                float y = acc[r][i];
                if (bias) y += __ldg(bias + i);
                if (r_ptr && !io_vec) y += __ldg(r_ptr + i);
                acc[r][i] = y;
            }
            if (io_vec) {
#pragma unroll
                for (int i = 0; i < COUT; i += 4) {
                    float4 y = make_float4(acc[r][i], acc[r][i + 1], acc[r][i + 2], acc[r][i + 3]);
                    if (r_ptr) {
                        const float4 rv = __ldg(reinterpret_cast<const float4 *>(r_ptr + i));
                        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
                    }
                    if (flags & PCGC_EPI_RELU) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    *reinterpret_cast<float4 *>(o_ptr + i) = y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < COUT; ++i) o_ptr[i] = (flags & PCGC_EPI_RELU) ? fmaxf(acc[r][i], 0.f) : acc[r][i];
            }
        }
    }
}

}  // namespace pcgc
