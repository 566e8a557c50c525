// conv_tile.cuh -- tiled implicit-GEMM k=3 sparse convolution for WIDE layers whose 27 weight
// matrices do not fit in shared memory (CIN*COUT > ~1500: 64x64, 64x32, 32x64, 128x128...).
// A CTA owns a tile of TM=64 consecutive output rows.  For each kernel offset k the gathered
// input rows A_k [TM x CIN] (zero rows for missing neighbours) and the weight matrix
// W_k [CIN x COUT] are staged in shared memory with cp.async (double buffered: the gather of
// offset k+1 overlaps the math of offset k) and the product is accumulated output-stationary
// in registers (4 x COUT/16 micro-tile per thread).  The output tile is written once, fused
// with bias / residual / ReLU.  FP32 FFMA math keeps the 1e-4 activation tolerance.
#pragma once
#include "common.cuh"

namespace pcgc {

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool pred) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int bytes = pred ? 16 : 0;            // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int CIN, int COUT>
struct TileCfg {
    static constexpr int TM = 64;
    static constexpr int THREADS = 256;
    static constexpr int A_LD = CIN + 4;             // +4 floats: rows 4 apart land 16 banks apart
    static constexpr int RM = 4;
    static constexpr int RN = COUT / 16;
    static_assert(CIN % 4 == 0 && COUT % 32 == 0 && RN >= 2 && RN <= 8, "tile kernel: CIN%4, COUT in {32,64,128}");
    static constexpr size_t smem_bytes() {
        return sizeof(int32_t) * 27 * TM + sizeof(float) * 2 * (TM * A_LD + CIN * COUT);
    }
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv_k3_tile_kernel(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ nbr, int64_t n,
                    const float *__restrict__ weight, const float *__restrict__ bias,
                    const float *__restrict__ residual, int res_ld, float *__restrict__ out, int out_ld, int flags) {
    using T = TileCfg<CIN, COUT>;
    constexpr int TM = T::TM, A_LD = T::A_LD, RM = T::RM, RN = T::RN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *idx_s = reinterpret_cast<int32_t *>(smem_raw);                 // [27][TM]
    float *a_s = reinterpret_cast<float *>(idx_s + 27 * TM);                // [2][TM][A_LD]
    float *w_s = a_s + 2 * TM * A_LD;                                       // [2][CIN][COUT]
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int64_t n_tiles = (n + TM - 1) / TM;

    auto prefetch = [&](int k, int buf) {
        constexpr int A_CHUNKS = TM * (CIN / 4);
        float *a_dst = a_s + buf * TM * A_LD;
        for (int c = t; c < A_CHUNKS; c += T::THREADS) {
            const int r = c / (CIN / 4), c4 = c % (CIN / 4);
            const int32_t src = idx_s[k * TM + r];
            cp_async16(a_dst + r * A_LD + 4 * c4, in + (int64_t)(src < 0 ? 0 : src) * in_ld + 4 * c4, src >= 0);
        }
        constexpr int W_CHUNKS = CIN * COUT / 4;
        float *w_dst = w_s + buf * CIN * COUT;
        const float *w_src = weight + (int64_t)k * CIN * COUT;
        for (int c = t; c < W_CHUNKS; c += T::THREADS) cp_async16(w_dst + 4 * c, w_src + 4 * c, true);
        cp_async_commit();
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        __syncthreads();                                   // previous tile's readers are done
        for (int i = t; i < 27 * TM; i += T::THREADS) {
            const int k = i / TM, r = i % TM;
            idx_s[i] = (row0 + r < n) ? __ldg(nbr + (int64_t)k * n + row0 + r) : -1;
        }
        __syncthreads();
        float acc[RM][RN];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
        prefetch(0, 0);
        for (int k = 0; k < 27; ++k) {
            const int buf = k & 1;
            if (k + 1 < 27) {
                prefetch(k + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float *a_b = a_s + buf * TM * A_LD + (ty * RM) * A_LD;
            const float *w_b = w_s + buf * CIN * COUT + tx * RN;
#pragma unroll 4
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
                float4 a4[RM];
#pragma unroll
                for (int i = 0; i < RM; ++i) a4[i] = *reinterpret_cast<const float4 *>(a_b + i * A_LD + 4 * c4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float b[RN];
                    const float *wp = w_b + (4 * c4 + e) * COUT;
                    if constexpr (RN == 2) {
                        const float2 v = *reinterpret_cast<const float2 *>(wp);
                        b[0] = v.x; b[1] = v.y;
                    } else {
#pragma unroll
                        for (int j = 0; j < RN; j += 4) {
                            const float4 v = *reinterpret_cast<const float4 *>(wp + j);
                            b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < RM; ++i) {
                        const float a = e == 0 ? a4[i].x : e == 1 ? a4[i].y : e == 2 ? a4[i].z : a4[i].w;
#pragma unroll
                        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a, b[j], acc[i][j]);
                    }
                }
            }
            __syncthreads();                               // buffer `buf` may be refilled next iteration
        }
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int64_t row = row0 + ty * RM + i;
            if (row >= n) continue;
            float *o = out + row * out_ld + tx * RN;
#pragma unroll
            for (int j = 0; j < RN; ++j) {
                float v = acc[i][j];
                if (bias) v += __ldg(bias + tx * RN + j);
                if (residual) v += __ldg(residual + row * res_ld + tx * RN + j);
                if (flags & PCGC_EPI_RELU) v = fmaxf(v, 0.f);
                o[j] = v;
            }
        }
    }
}

}  // namespace pcgc
