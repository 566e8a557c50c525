// conv_ones.cu -- the FIRST layer of the analysis network (encoder.conv0, autoencoder.py:71,138: k=3, 1 -> 16, ReLU).
//
// Its input features are the constant 1 the reference attaches to every voxel (data_utils.py:94, coder.py:131:
// `feats = torch.ones(...)`), so   out[u] = bias + sum over the PRESENT neighbours k of W[k][0][:]   -- the layer needs the 27-bit
// occupancy of each voxel's neighbourhood, not a kernel map.  That occupancy comes out of the PARENT set's kernel map and the
// parents' child-occupancy bytes exactly as in pcgc_kernel_map_k3_from_parent (a voxel's 3x3x3 neighbourhood touches 2x2x2 parent
// cells: 7 map reads + 8 info reads per voxel), so the 27 x N int32 kernel map of the finest analysis level -- 86 MB at 795 k
// voxels, written once and read once, by this layer only -- is never built.
//
// One warp = 32 voxels: each lane derives the mask of its voxel, then the warp walks its voxels two at a time with lanes =
// output channels (27 weights per lane in registers): 64-byte fp32 rows and 16-byte h2 groups leave fully coalesced.
#include "common.cuh"
#include "conv_h2.cuh"

namespace pcgc {

__global__ void __launch_bounds__(256)
conv_k3_ones_from_parent_kernel(const uint64_t *__restrict__ child_keys, const int32_t *__restrict__ parent_of,
                                const uint64_t *__restrict__ info, const int32_t *__restrict__ pnbr, int64_t n_parents, int64_t n,
                                const float *__restrict__ weight, const float *__restrict__ bias, float *__restrict__ out, int out_ld,
                                uint32_t *__restrict__ out_h2, int out_h2_ld, int flags, int *__restrict__ overflow) {
    constexpr int COUT = 16;
    const int lane = threadIdx.x & 31, c = lane & 15, half = lane >> 4;
    float w[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) w[k] = __ldg(weight + k * COUT + c);
    const float b = bias ? __ldg(bias + c) : 0.f;
    bool over = false;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5, warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int64_t base = warp0 * 32; base < n; base += warps * 32) {
        // ---- lane = voxel: 27-bit presence mask of its neighbourhood
        const int64_t i = base + lane;
        uint32_t mask = 0;
        if (i < n) {
            const int cc0 = (int)(child_keys[i] & 7);
            const int64_t p = parent_of[i];
            const int cx = cc0 & 1, cy = (cc0 >> 1) & 1, cz = cc0 >> 2;
            const int lx = (cx - 1) >> 1, ly = (cy - 1) >> 1, lz = (cz - 1) >> 1;          // lower of the two parent offsets per axis
            uint64_t occ = 0;                                                               // child-occupancy bytes of the 2x2x2 parent cells
#pragma unroll
            for (int cell = 0; cell < 8; ++cell) {
                const int kp = (lx + (cell & 1) + 1) + 3 * (ly + ((cell >> 1) & 1) + 1) + 9 * (lz + (cell >> 2) + 1);
                const int64_t q = kp == 13 ? p : (int64_t)__ldg(pnbr + (int64_t)kp * n_parents + p);
                const uint64_t o = q < 0 ? 0ull : (__ldg(info + q) & 0xFFull);
                occ |= o << (8 * cell);
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                const int tx = cx + k % 3 - 1, ty = cy + (k / 3) % 3 - 1, tz = cz + k / 9 - 1;    // in {-1, 0, 1, 2}
                const int cell = ((tx >> 1) - lx) | (((ty >> 1) - ly) << 1) | (((tz >> 1) - lz) << 2);
                const int cc = (tx & 1) | ((ty & 1) << 1) | ((tz & 1) << 2);
                mask |= (uint32_t)((occ >> (8 * cell + cc)) & 1ull) << k;
            }
        }
        // ---- lanes = channels: two voxels per step (one per half warp)
#pragma unroll 1
        for (int it = 0; it < 16; ++it) {
            const int r = 2 * it + half;
            const uint32_t m = __shfl_sync(0xffffffffu, mask, r);
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 27; ++k) acc += (m >> k) & 1u ? w[k] : 0.f;                  // ascending k, like the gather kernels
            float v = acc + b;
            if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
            const int64_t row = base + r;
            const float v1 = __shfl_down_sync(0xffffffffu, v, 1);                            // channel c + 1
            uint32_t hi, lo;
            split_pair_h2(v, v1, hi, lo);                                                    // valid on even c
            const uint32_t hi2 = __shfl_down_sync(0xffffffffu, hi, 2), lo2 = __shfl_down_sync(0xffffffffu, lo, 2);
            if (row < n) {
                if (out) out[row * out_ld + c] = v;
                if (out_h2) {
                    over |= !(fabsf(v) <= kH2Limit);
                    if ((c & 3) == 0) *reinterpret_cast<uint4 *>(out_h2 + row * out_h2_ld + c) = make_uint4(hi, hi2, lo, lo2);
                }
            }
        }
    }
    if (over && overflow) *overflow = 1;
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

int pcgc_conv_k3_ones_from_parent_fwd(const uint64_t *child_keys, const int32_t *parent_of, const uint64_t *parent_info,
                                      const int32_t *parent_nbr, int64_t n_parents, int64_t n, const float *weight, const float *bias,
                                      int32_t cout, float *out, int32_t out_ld, uint32_t *out_h2, int32_t out_h2_ld, int32_t flags,
                                      int32_t *overflow, void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF && n_parents >= 0 && cout == 16, "pcgc_conv_k3_ones_from_parent_fwd: bad shape (cout must be 16)");
    if (n == 0) return PCGC_OK;
    PCGC_REQUIRE(child_keys && parent_of && parent_info && parent_nbr && weight && (out || out_h2), "pcgc_conv_k3_ones_from_parent_fwd: null pointer");
    PCGC_REQUIRE(!out || out_ld >= cout, "pcgc_conv_k3_ones_from_parent_fwd: out_ld < cout");
    PCGC_REQUIRE(!out_h2 || (out_h2_ld >= cout && out_h2_ld % 4 == 0 && ((uintptr_t)out_h2 & 15) == 0),
                 "pcgc_conv_k3_ones_from_parent_fwd: h2 output rows must be 16-byte aligned");
    conv_k3_ones_from_parent_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(child_keys, parent_of, parent_info, parent_nbr,
                                                                                          n_parents, n, weight, bias, out, out_ld, out_h2,
                                                                                          out_h2_ld, flags, overflow);
    return check_launch("conv_k3_ones_from_parent");
}

}  // extern "C"
