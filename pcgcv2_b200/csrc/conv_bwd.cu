// conv_bwd.cu -- backward passes of the sparse convolutions (the reference drives
// them through trainer.py:136 `sum_loss.backward()` into MinkowskiEngine's Convolution*Backward).
//
//   grad wrt INPUT of k=3 / k=1 convolutions is itself a forward convolution (stride-1 kernel maps are
//   symmetric: nbr[k][u] = v  <=>  nbr[26-k][v] = u), so the host layer calls the forward kernels with
//   the transposed, offset-flipped weights.  This file holds what has no forward twin:
//     * weight gradients  gW[k] = sum over pairs (a, b) of  A[a]^T (x) G[b]   (all four conv types),
//     * input gradients of the k=2 s=2 down conv and the generative k=2 s=2 up conv.
//   Generic in the channel counts, FP32 FFMA, shared-memory row tiles.  DETERMINISTIC: every block writes its partial
//   [kvol][ca][cb] sums to the caller's workspace and a second pass adds the partials in block order (no atomics), so the
//   gradients -- and with them a data-parallel run's all-reduced update -- are bit-reproducible.
#include "common.cuh"

namespace pcgc {

constexpr int kWgradTileRows = 64;

enum PairMode { PAIR_K3 = 0, PAIR_IDENT = 1, PAIR_DOWN = 2, PAIR_UP = 3 };

// one (blockIdx.y = k) x (chunk of rows): gw[k][ca][cb] += sum_rows A[ia]^T (x) B[ib]
//   K3   : ia = nbr[k*n + u] (skip < 0), ib = u            (A = layer input, B = grad_out)
//   IDENT: ia = ib = u, k = 0                              (k = 1 convolution)
//   DOWN : ia = u (child row), ib = parent_of[u], only rows with (keys[u] & 7) == k
//   UP   : ia = u, ib = 8u + k
template <int E>
__global__ void __launch_bounds__(256)
weight_grad_kernel(int mode, const float *__restrict__ A, int a_ld, const float *__restrict__ B, int b_ld,
                   const int32_t *__restrict__ nbr, const int32_t *__restrict__ parent_of,
                   const uint64_t *__restrict__ keys, int64_t n, int ca, int cb, int rows_per_block,
                   float *__restrict__ partial) {
    constexpr int TR = kWgradTileRows;     // rows per shared-memory tile: 64 FMAs per element between two barriers (16 left the
    extern __shared__ float sm[];          // kernel barrier-bound: 64 of the 94 ms of GPU time of a config-5 training step)
    float *as = sm;                    // [TR][ca]
    float *bs = sm + TR * ca;          // [TR][cb]
    __shared__ int64_t ia_s[TR], ib_s[TR];
    const int k = blockIdx.y;
    const int total = ca * cb;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(n, r_begin + rows_per_block);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += TR) {
        if (threadIdx.x < TR) {
            const int64_t u = r0 + threadIdx.x;
            int64_t ia = -1, ib = -1;
            if (u < r_end) {
                if (mode == PAIR_K3) { ia = nbr[(int64_t)k * n + u]; ib = u; }
                else if (mode == PAIR_IDENT) { ia = u; ib = u; }
                else if (mode == PAIR_DOWN) { if ((int)(keys[u] & 7) == k) { ia = u; ib = parent_of[u]; } }
                else { ia = u; ib = 8 * u + k; }
            }
            ia_s[threadIdx.x] = ia;
            ib_s[threadIdx.x] = ia < 0 ? -1 : ib;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TR * ca; i += blockDim.x) {
            const int r = i / ca, c = i % ca;
            as[i] = ia_s[r] >= 0 ? A[ia_s[r] * a_ld + c] : 0.f;
        }
        for (int i = threadIdx.x; i < TR * cb; i += blockDim.x) {
            const int r = i / cb, c = i % cb;
            bs[i] = ib_s[r] >= 0 ? B[ib_s[r] * b_ld + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int idx = threadIdx.x + e * 256;
            if (idx < total) {
                const int ci = idx / cb, co = idx % cb;
                float s = 0.f;
#pragma unroll 16
                for (int r = 0; r < TR; ++r) s = fmaf(as[r * ca + ci], bs[r * cb + co], s);
                acc[e] += s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int idx = threadIdx.x + e * 256;
        if (idx < total) partial[((int64_t)blockIdx.x * gridDim.y + k) * total + idx] = acc[e];
    }
}

// The same sums with register micro-tiles, for channel counts that are multiples of 4 with at most 256 micro-tiles (every layer
// of the codec except the 1-channel ends): thread (m, sl) owns the 4 x 4 block m of gw[k] and the rows sl, sl + S, ... of each
// shared-memory tile (S = 256 / micro-tiles row slices), i.e. two LDS.128 feed 16 FMAs (the element-per-thread kernel above
// issues two LDS per FMA); the S slices are added in slice order through shared memory at the end -- deterministic like the rest.
__global__ void __launch_bounds__(256)
weight_grad_mt_kernel(int mode, const float *__restrict__ A, int a_ld, const float *__restrict__ B, int b_ld,
                      const int32_t *__restrict__ nbr, const int32_t *__restrict__ parent_of, const uint64_t *__restrict__ keys,
                      int64_t n, int ca, int cb, int rows_per_block, float *__restrict__ partial) {
    constexpr int TR = kWgradTileRows;
    extern __shared__ __align__(16) float sm[];
    float *as = sm;                    // [TR][ca]
    float *bs = sm + TR * ca;          // [TR][cb]
    __shared__ int64_t ia_s[TR], ib_s[TR];
    const int k = blockIdx.y, total = ca * cb;
    const int nbj = cb >> 2, nm = (ca >> 2) * nbj, S = 256 / nm;
    const int m = threadIdx.x % nm, sl = threadIdx.x / nm;
    const int mi = m / nbj, mj = m % nbj;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(n, r_begin + rows_per_block);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += TR) {
        if (threadIdx.x < TR) {
            const int64_t u = r0 + threadIdx.x;
            int64_t ia = -1, ib = -1;
            if (u < r_end) {
                if (mode == PAIR_K3) { ia = nbr[(int64_t)k * n + u]; ib = u; }
                else if (mode == PAIR_IDENT) { ia = u; ib = u; }
                else if (mode == PAIR_DOWN) { if ((int)(keys[u] & 7) == k) { ia = u; ib = parent_of[u]; } }
                else { ia = u; ib = 8 * u + k; }
            }
            ia_s[threadIdx.x] = ia;
            ib_s[threadIdx.x] = ia < 0 ? -1 : ib;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TR * ca; i += blockDim.x) {
            const int r = i / ca, c = i % ca;
            as[i] = ia_s[r] >= 0 ? A[ia_s[r] * a_ld + c] : 0.f;
        }
        for (int i = threadIdx.x; i < TR * cb; i += blockDim.x) {
            const int r = i / cb, c = i % cb;
            bs[i] = ib_s[r] >= 0 ? B[ib_s[r] * b_ld + c] : 0.f;
        }
        __syncthreads();
        if (sl < S) {
            for (int r = sl; r < TR; r += S) {
                const float4 a = *reinterpret_cast<const float4 *>(as + r * ca + 4 * mi);
                const float4 b = *reinterpret_cast<const float4 *>(bs + r * cb + 4 * mj);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    // slice sums -> shared memory [S][total] (element ci * cb + co), then added in slice order
    float *red = sm;
    if (sl < S) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(size_t)sl * total + (4 * mi + i) * cb + 4 * mj + j] = acc[i][j];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        float sum = 0.f;
        for (int q = 0; q < S; ++q) sum += red[(size_t)q * total + idx];
        partial[((int64_t)blockIdx.x * gridDim.y + k) * total + idx] = sum;
    }
}

// out[i] = partial[0][i] + partial[1][i] + ... in block order
__global__ void sum_partials_kernel(const float *__restrict__ partial, int blocks, int64_t count, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * count + i];
        out[i] = s;
    }
}

// gi[u][ci] = sum_co G[parent_of[u]][co] * W[keys[u]&7][ci][co]
__global__ void down_bwd_data_kernel(const float *__restrict__ go, int go_ld, const uint64_t *__restrict__ keys,
                                     const int32_t *__restrict__ parent_of, int64_t n, const float *__restrict__ w,
                                     int cin, int cout, float *__restrict__ gi, int gi_ld) {
    const int64_t total = n * cin;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / cin;
        const int ci = (int)(i % cin);
        const float *g = go + (int64_t)parent_of[u] * go_ld;
        const float *wk = w + ((int64_t)(keys[u] & 7) * cin + ci) * cout;
        float s = 0.f;
        for (int co = 0; co < cout; ++co) s = fmaf(g[co], wk[co], s);
        gi[u * gi_ld + ci] = s;
    }
}

// gi[u][ci] = sum_k sum_co G[8u+k][co] * W[k][ci][co]
__global__ void up_bwd_data_kernel(const float *__restrict__ go, int go_ld, int64_t n, const float *__restrict__ w,
                                   int cin, int cout, float *__restrict__ gi, int gi_ld) {
    const int64_t total = n * cin;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / cin;
        const int ci = (int)(i % cin);
        float s = 0.f;
        for (int k = 0; k < 8; ++k) {
            const float *g = go + (8 * u + k) * go_ld;
            const float *wk = w + ((int64_t)k * cin + ci) * cout;
            for (int co = 0; co < cout; ++co) s = fmaf(g[co], wk[co], s);
        }
        gi[u * gi_ld + ci] = s;
    }
}

// partial[block][c] = sum over the block's rows of x[r][c]  (bias gradients; summed in block order by sum_partials_kernel)
__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ x, int ld, int64_t n, int c, float *__restrict__ partial) {
    __shared__ float red[256];
    const int col = threadIdx.x % c, lane_row = threadIdx.x / c, rows_per_iter = blockDim.x / c;
    float s = 0.f;
    if (lane_row < rows_per_iter)
        for (int64_t r = (int64_t)blockIdx.x * rows_per_iter + lane_row; r < n; r += (int64_t)gridDim.x * rows_per_iter)
            s += x[r * ld + col];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < c) {
        float t = 0.f;
        for (int j = 0; j < rows_per_iter; ++j) t += red[j * c + threadIdx.x];
        partial[(int64_t)blockIdx.x * c + threadIdx.x] = t;
    }
}

constexpr int kColsumBlocks = 2 * kNumSMs;

static int weight_grad_rows_per_block(int64_t n) {          // at most 256 partial blocks per kernel offset
    int64_t r = (n + 255) / 256;
    r = (r + 15) / 16 * 16;
    return (int)(r < 2048 ? 2048 : r);
}
static size_t weight_grad_ws_bytes(int64_t n, int kvol, int ca, int cb) {
    if (n <= 0) return 0;
    const int rpb = weight_grad_rows_per_block(n);
    return sizeof(float) * (size_t)((n + rpb - 1) / rpb) * kvol * ca * cb;
}

static int launch_weight_grad(int mode, const float *A, int a_ld, const float *B, int b_ld, const int32_t *nbr,
                              const int32_t *parent_of, const uint64_t *keys, int64_t n, int kvol, int ca, int cb,
                              float *gw, void *ws, size_t ws_bytes, cudaStream_t s) {
    PCGC_REQUIRE(ca >= 1 && cb >= 1 && ca * cb <= 256 * 64, "weight gradient: %dx%d channels not supported", ca, cb);
    if (n == 0) {
        PCGC_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)kvol * ca * cb, s));
        return PCGC_OK;
    }
    PCGC_REQUIRE(ws && ws_bytes >= weight_grad_ws_bytes(n, kvol, ca, cb), "weight gradient: workspace too small");
    const int rows_per_block = weight_grad_rows_per_block(n);
    dim3 grid((unsigned)((n + rows_per_block - 1) / rows_per_block), kvol);
    float *part = (float *)ws;
    const int nm = (ca / 4) * (cb / 4);
    if (ca % 4 == 0 && cb % 4 == 0 && nm >= 1 && nm <= 256) {     // register micro-tiles
        const size_t tile = sizeof(float) * kWgradTileRows * (ca + cb), red = sizeof(float) * (size_t)(256 / nm) * ca * cb;
        const size_t smem_mt = tile > red ? tile : red;
        PCGC_REQUIRE(smem_mt <= 96 * 1024, "weight gradient: %dx%d channels need too much shared memory", ca, cb);
        if (smem_mt > 48 * 1024)                                  // per call: the attribute belongs to the current device's context
            PCGC_CUDA(cudaFuncSetAttribute(weight_grad_mt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        weight_grad_mt_kernel<<<grid, 256, smem_mt, s>>>(mode, A, a_ld, B, b_ld, nbr, parent_of, keys, n, ca, cb, rows_per_block, part);
        int rc = check_launch("weight_grad_mt");
        if (rc) return rc;
        const int64_t count = (int64_t)kvol * ca * cb;
        sum_partials_kernel<<<grid_for(count, 256, 4), 256, 0, s>>>(part, (int)grid.x, count, gw);
        return check_launch("weight_grad_sum");
    }
    const size_t smem = sizeof(float) * kWgradTileRows * (ca + cb);
    const int e = (ca * cb + 255) / 256;
#define WG(E)                                                                                                                          \
    do {                                                                                                                               \
        if (smem > 48 * 1024) cudaFuncSetAttribute(weight_grad_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        weight_grad_kernel<E><<<grid, 256, smem, s>>>(mode, A, a_ld, B, b_ld, nbr, parent_of, keys, n, ca, cb, rows_per_block, part); \
    } while (0)
    if (e <= 1) WG(1); else if (e <= 4) WG(4); else if (e <= 16) WG(16); else WG(64);
#undef WG
    int rc = check_launch("weight_grad");
    if (rc) return rc;
    const int64_t count = (int64_t)kvol * ca * cb;
    sum_partials_kernel<<<grid_for(count, 256, 4), 256, 0, s>>>(part, (int)grid.x, count, gw);
    return check_launch("weight_grad_sum");
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

size_t pcgc_conv_bwd_weight_ws_bytes(int64_t n, int32_t kvol, int32_t cin, int32_t cout) {
    return weight_grad_ws_bytes(n, kvol, cin, cout);
}

size_t pcgc_colsum_ws_bytes(void) { return sizeof(float) * kColsumBlocks * 256; }

int pcgc_conv_bwd_weight(const float *in, int32_t in_ld, const int32_t *nbr, int64_t n, int32_t kvol,
                         const float *grad_out, int32_t go_ld, int32_t cin, int32_t cout, float *grad_weight,
                         void *ws, size_t ws_bytes, void *stream) {
    PCGC_REQUIRE((kvol == 27 && nbr) || (kvol == 1 && !nbr), "pcgc_conv_bwd_weight: kvol 27 needs a kernel map, kvol 1 none");
    return launch_weight_grad(kvol == 27 ? PAIR_K3 : PAIR_IDENT, in, in_ld, grad_out, go_ld, nbr, nullptr, nullptr, n, kvol,
                              cin, cout, grad_weight, ws, ws_bytes, (cudaStream_t)stream);
}

int pcgc_conv_k2s2_bwd(const float *in, int32_t in_ld, const uint64_t *in_keys, const int32_t *parent_of, int64_t n_in,
                       const float *grad_out, int32_t go_ld, const float *weight, int32_t cin, int32_t cout,
                       float *grad_in, int32_t gi_ld, float *grad_weight, void *ws, size_t ws_bytes, void *stream) {
    PCGC_REQUIRE(in_keys && parent_of, "pcgc_conv_k2s2_bwd: null map");
    cudaStream_t s = (cudaStream_t)stream;
    if (grad_weight) {
        int rc = launch_weight_grad(PAIR_DOWN, in, in_ld, grad_out, go_ld, nullptr, parent_of, in_keys, n_in, 8, cin, cout,
                                    grad_weight, ws, ws_bytes, s);
        if (rc) return rc;
    }
    if (grad_in && n_in) {
        down_bwd_data_kernel<<<grid_for(n_in * cin, 256, 8), 256, 0, s>>>(grad_out, go_ld, in_keys, parent_of, n_in, weight,
                                                                         cin, cout, grad_in, gi_ld);
        return check_launch("down_bwd_data");
    }
    return PCGC_OK;
}

int pcgc_convT_k2s2_bwd(const float *in, int32_t in_ld, int64_t n_in, const float *grad_out, int32_t go_ld,
                        const float *weight, int32_t cin, int32_t cout, float *grad_in, int32_t gi_ld,
                        float *grad_weight, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (grad_weight) {
        int rc = launch_weight_grad(PAIR_UP, in, in_ld, grad_out, go_ld, nullptr, nullptr, nullptr, n_in, 8, cin, cout,
                                    grad_weight, ws, ws_bytes, s);
        if (rc) return rc;
    }
    if (grad_in && n_in) {
        up_bwd_data_kernel<<<grid_for(n_in * cin, 256, 8), 256, 0, s>>>(grad_out, go_ld, n_in, weight, cin, cout, grad_in,
                                                                       gi_ld);
        return check_launch("up_bwd_data");
    }
    return PCGC_OK;
}

int pcgc_colsum(const float *x, int32_t ld, int64_t n, int32_t c, float *out, void *ws, size_t ws_bytes, void *stream) {
    PCGC_REQUIRE(c >= 1 && c <= 256 && ld >= c, "pcgc_colsum: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        PCGC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * c, s));
        return PCGC_OK;
    }
    PCGC_REQUIRE(ws && ws_bytes >= pcgc_colsum_ws_bytes(), "pcgc_colsum: workspace too small");
    const int rows_per_iter = 256 / c;
    int blocks = grid_for(n, rows_per_iter * 8, 2);
    if (blocks > kColsumBlocks) blocks = kColsumBlocks;
    colsum_kernel<<<blocks, 256, 0, s>>>(x, ld, n, c, (float *)ws);
    int rc = check_launch("colsum");
    if (rc) return rc;
    sum_partials_kernel<<<1, 256, 0, s>>>((const float *)ws, blocks, c, out);
    return check_launch("colsum_sum");
}

}  // extern "C"
