// select.cu -- exact top-k mask (radix select) and stable pruning (scan + compaction).
// SURVEY section 8 rows a9 (istopk, data_utils.py:77-89: torch.topk on the CPU in the reference)
// and a8 (ME.MinkowskiPruning, autoencoder.py:237,247).  Pure HBM-bound passes.
#include <cub/cub.cuh>

#include "common.cuh"

namespace pcgc {

__device__ __forceinline__ uint32_t orderable(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);       // ascending uint order == ascending float order
}

struct SelectState {
    uint32_t prefix;      // threshold bits decided so far (high bits)
    uint32_t k_rem;       // how many of the elements matching the prefix are still wanted
    uint32_t hist[256];
};

__global__ void select_init_kernel(SelectState *st, uint32_t k) {
    if (threadIdx.x == 0) { st->prefix = 0; st->k_rem = k; }
    st->hist[threadIdx.x] = 0;
}

// histogram of the byte at `shift` over the elements whose higher bits equal the prefix
__global__ void select_hist_kernel(const float *__restrict__ x, int ld, int64_t n, int shift, SelectState *st) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = st->prefix;
    const uint32_t hi_mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = orderable(__ldg(x + i * ld));
        if ((u & hi_mask) == prefix) atomicAdd(&h[(u >> shift) & 255], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

// pick the bin that holds the k_rem-th largest element; fold it into the prefix
__global__ void select_pick_kernel(SelectState *st, int shift) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = st->hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t k = st->k_rem, above = 0;
        int b = 255;
        for (; b > 0; --b) {
            if (above + h[b] >= k) break;
            above += h[b];
        }
        st->prefix |= (uint32_t)b << shift;
        st->k_rem = k - above;
    }
    __syncthreads();
    st->hist[threadIdx.x] = 0;
}

__global__ void select_eqflag_kernel(const float *__restrict__ x, int ld, int64_t n, const SelectState *st,
                                     int32_t *__restrict__ eq) {
    const uint32_t thr = st->prefix;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        eq[i] = orderable(__ldg(x + i * ld)) == thr;
}

__global__ void select_mask_kernel(const float *__restrict__ x, int ld, int64_t n, const SelectState *st,
                                   const int32_t *__restrict__ eq_rank, uint8_t *__restrict__ mask) {
    const uint32_t thr = st->prefix, need = st->k_rem;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = orderable(__ldg(x + i * ld));
        mask[i] = (u > thr) || (u == thr && (uint32_t)eq_rank[i] < need);
    }
}

__global__ void fill_u8_kernel(uint8_t *p, int64_t n, uint8_t v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- prune ----------------------------------------------------------------------------------
__global__ void mask_to_i32_kernel(const uint8_t *__restrict__ m, int64_t n, int32_t *__restrict__ o) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = m[i] != 0;
}

// one thread per (row, 16-byte chunk): consecutive threads copy consecutive chunks of a row
__global__ void prune_copy_kernel(const uint8_t *__restrict__ mask, int64_t n, const int32_t *__restrict__ pos,
                                  const uint64_t *__restrict__ keys, const float *__restrict__ feats, int ld,
                                  int channels, uint64_t *__restrict__ keys_out, float *__restrict__ feats_out,
                                  int out_ld, int32_t *__restrict__ n_kept, int vec) {
    const int chunks = vec ? channels / 4 : channels;
    const int64_t total = n * chunks;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / chunks;
        const int c = (int)(i % chunks);
        if (c == 0 && row == n - 1) *n_kept = pos[row] + (mask[row] != 0);
        if (!mask[row]) continue;
        const int64_t dst = pos[row];
        if (c == 0 && keys) keys_out[dst] = keys[row];
        if (vec)
            reinterpret_cast<float4 *>(feats_out + dst * out_ld)[c] = __ldg(reinterpret_cast<const float4 *>(feats + row * ld) + c);
        else
            feats_out[dst * out_ld + c] = __ldg(feats + row * ld + c);
    }
}

// kernel map of the pruned set = kernel map of the full set filtered through the same compaction
__global__ void prune_map_kernel(const uint8_t *__restrict__ mask, int64_t n, const int32_t *__restrict__ pos,
                                 const int32_t *__restrict__ nbr_in, const int32_t *__restrict__ n_kept,
                                 int32_t *__restrict__ nbr_out) {
    const int k = blockIdx.y;
    const int64_t stride_out = *n_kept;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!mask[i]) continue;
        const int32_t v = nbr_in[(int64_t)k * n + i];
        nbr_out[(int64_t)k * stride_out + pos[i]] = (v >= 0 && mask[v]) ? pos[v] : -1;
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t scan_temp_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int32_t *)nullptr, n > 0 ? n : 1);
    return b;
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

size_t pcgc_topk_mask_ws_bytes(int64_t n) {
    if (n < 1) n = 1;
    return align256(sizeof(SelectState)) + 2 * align256(sizeof(int32_t) * (size_t)n) + align256(scan_temp_bytes(n));
}

int pcgc_topk_mask(const float *logits, int32_t ld, int64_t n, int64_t k, uint8_t *mask, void *ws, size_t ws_bytes,
                   void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF && k >= 0 && ld >= 1, "pcgc_topk_mask: bad arguments");
    if (n == 0) return PCGC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int g = grid_for(n, 256, 8);
    if (k == 0 || k >= n) {
        fill_u8_kernel<<<g, 256, 0, s>>>(mask, n, k == 0 ? 0 : 1);
        return check_launch("fill_u8");
    }
    if (ws_bytes < pcgc_topk_mask_ws_bytes(n)) {
        set_error("pcgc_topk_mask: workspace too small");
        return PCGC_ERR_WORKSPACE;
    }
    char *p = (char *)ws;
    SelectState *st = (SelectState *)p;  p += align256(sizeof(SelectState));
    int32_t *eq = (int32_t *)p;          p += align256(sizeof(int32_t) * (size_t)n);
    int32_t *rank = (int32_t *)p;        p += align256(sizeof(int32_t) * (size_t)n);
    void *scan_ws = p;
    size_t scan_bytes = scan_temp_bytes(n);
    int rc;
    select_init_kernel<<<1, 256, 0, s>>>(st, (uint32_t)k);
    if ((rc = check_launch("select_init"))) return rc;
    for (int shift = 24; shift >= 0; shift -= 8) {
        select_hist_kernel<<<g, 256, 0, s>>>(logits, ld, n, shift, st);
        if ((rc = check_launch("select_hist"))) return rc;
        select_pick_kernel<<<1, 256, 0, s>>>(st, shift);
        if ((rc = check_launch("select_pick"))) return rc;
    }
    select_eqflag_kernel<<<g, 256, 0, s>>>(logits, ld, n, st, eq);
    if ((rc = check_launch("select_eqflag"))) return rc;
    PCGC_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, eq, rank, n, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    select_mask_kernel<<<g, 256, 0, s>>>(logits, ld, n, st, rank, mask);
    return check_launch("select_mask");
}

size_t pcgc_prune_ws_bytes(int64_t n) {
    if (n < 1) n = 1;
    return 2 * align256(sizeof(int32_t) * (size_t)n) + align256(scan_temp_bytes(n));
}

int pcgc_prune(const uint8_t *mask, int64_t n, const uint64_t *keys, const float *feats, int32_t ld, int32_t channels,
               uint64_t *keys_out, float *feats_out, int32_t out_ld, int32_t *n_kept, const int32_t *nbr_in,
               int32_t *nbr_out, void *ws, size_t ws_bytes, void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF && channels >= 1 && ld >= channels && out_ld >= channels, "pcgc_prune: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        PCGC_CUDA(cudaMemsetAsync(n_kept, 0, sizeof(int32_t), s));
        return PCGC_OK;
    }
    if (ws_bytes < pcgc_prune_ws_bytes(n)) {
        set_error("pcgc_prune: workspace too small");
        return PCGC_ERR_WORKSPACE;
    }
    char *p = (char *)ws;
    int32_t *flag = (int32_t *)p;  p += align256(sizeof(int32_t) * (size_t)n);
    int32_t *pos = (int32_t *)p;   p += align256(sizeof(int32_t) * (size_t)n);
    void *scan_ws = p;
    size_t scan_bytes = scan_temp_bytes(n);
    int rc;
    mask_to_i32_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(mask, n, flag);
    if ((rc = check_launch("mask_to_i32"))) return rc;
    PCGC_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, flag, pos, n, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int vec = (channels % 4 == 0) && (ld % 4 == 0) && (out_ld % 4 == 0) && (((uintptr_t)feats & 15) == 0) &&
                    (((uintptr_t)feats_out & 15) == 0);
    const int64_t total = n * (vec ? channels / 4 : channels);
    prune_copy_kernel<<<grid_for(total, 256, 8), 256, 0, s>>>(mask, n, pos, keys, feats, ld, channels, keys_out,
                                                             feats_out, out_ld, n_kept, vec);
    if ((rc = check_launch("prune_copy"))) return rc;
    if (nbr_in && nbr_out) {
        dim3 grid(grid_for(n, 256, 2), 27);
        prune_map_kernel<<<grid, 256, 0, s>>>(mask, n, pos, nbr_in, n_kept, nbr_out);
        rc = check_launch("prune_map");
    }
    return rc;
}

}  // extern "C"
