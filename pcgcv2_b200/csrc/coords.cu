// coords.cu -- coordinate maps: Morton keys, hash table, k=3 kernel map, stride-2 down map,
// 8-child up map, argsort.  All integer work, HBM/L2-bound (SURVEY section 8 rows a1,a4,a5,a6,a10,a11).
#include <cub/cub.cuh>

#include "common.cuh"

namespace pcgc {

// ---- pack / unpack ----------------------------------------------------------------------
__global__ void pack_keys_kernel(const int4 *__restrict__ coords, int64_t n, int32_t stride,
                                 uint64_t *__restrict__ keys, int32_t *__restrict__ err) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int4 c = coords[i];                      // (b, x, y, z): one 16-byte load per row
        int x = c.y / stride, y = c.z / stride, z = c.w / stride;
        bool bad = (c.x < 0) | (c.x > PCGC_MAX_BATCH) | (c.y < 0) | (c.z < 0) | (c.w < 0) |
                   (x > PCGC_MAX_COORD) | (y > PCGC_MAX_COORD) | (z > PCGC_MAX_COORD) |
                   (x * stride != c.y) | (y * stride != c.z) | (z * stride != c.w);
        if (bad) {
            *err = 1;
            keys[i] = 0;
        } else {
            keys[i] = make_key((uint32_t)c.x, (uint32_t)x, (uint32_t)y, (uint32_t)z);
        }
    }
}

// the same from int32 [n,3] (x, y, z) rows with one batch index for all (Codec's input: no padded copy of the cloud)
__global__ void pack_keys3_kernel(const int32_t *__restrict__ coords, int64_t n, int32_t stride, int32_t batch, int32_t hint_bits,
                                  uint64_t *__restrict__ keys, int32_t *__restrict__ err) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int cx = coords[3 * i], cy = coords[3 * i + 1], cz = coords[3 * i + 2];
        const int x = cx / stride, y = cy / stride, z = cz / stride;
        const bool bad = (cx < 0) | (cy < 0) | (cz < 0) | (x > PCGC_MAX_COORD) | (y > PCGC_MAX_COORD) | (z > PCGC_MAX_COORD) |
                         (x * stride != cx) | (y * stride != cy) | (z * stride != cz);
        if (bad) {
            *err = 1;
            keys[i] = 0;
        } else {
            if (hint_bits > 0 && ((((uint32_t)x | (uint32_t)y | (uint32_t)z) >> hint_bits) != 0u || batch != 0)) atomicOr(err, 2);
            keys[i] = make_key((uint32_t)batch, (uint32_t)x, (uint32_t)y, (uint32_t)z);
        }
    }
}

// child_map[k][p] = row of child k (= key & 7) of parent p, -1 where absent (the gather map of the k=2 stride-2 convolution)
__global__ void child_map_k2_kernel(const uint64_t *__restrict__ child_keys, const int32_t *__restrict__ parent_of, int64_t n,
                                    int64_t n_parents, int32_t *__restrict__ child_map) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        child_map[(int64_t)(child_keys[i] & 7) * n_parents + parent_of[i]] = (int32_t)i;
}

__global__ void unpack_keys_kernel(const uint64_t *__restrict__ keys, int64_t n, int32_t stride,
                                   int4 *__restrict__ coords) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t b, x, y, z;
        split_key(keys[i], b, x, y, z);
        coords[i] = make_int4((int)b, (int)x * stride, (int)y * stride, (int)z * stride);
    }
}

// scale_sparse_tensor (data_utils.py:112-118): float32 multiply, round half to even (torch.round), int cast
__global__ void scale_coords_kernel(const int32_t *__restrict__ in, int64_t count, float factor, int32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2int_rn(__fmul_rn((float)in[i], factor));
}

// ---- hash table ---------------------------------------------------------------------------
__global__ void hash_clear_kernel(uint64_t *__restrict__ tkeys, int32_t *__restrict__ tvals, int64_t cap) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        tkeys[i] = PCGC_EMPTY_KEY;
        tvals[i] = 0x7FFFFFFF;
    }
}

__global__ void hash_insert_kernel(const uint64_t *__restrict__ keys, int64_t n, uint64_t *__restrict__ tkeys,
                                   int32_t *__restrict__ tvals, uint64_t mask, int32_t *__restrict__ n_dup) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        uint64_t s = hash_slot(key, mask);
        while (true) {
            unsigned long long prev = atomicCAS((unsigned long long *)(tkeys + s), PCGC_EMPTY_KEY, key);
            if (prev == PCGC_EMPTY_KEY) {
                atomicMin(tvals + s, (int32_t)i);
                break;
            }
            if (prev == key) {                   // duplicate coordinate: keep the first-seen row
                atomicMin(tvals + s, (int32_t)i);
                atomicAdd(n_dup, 1);
                break;
            }
            s = (s + 1) & mask;
        }
    }
}

__global__ void hash_keep_kernel(const uint64_t *__restrict__ keys, int64_t n, const uint64_t *__restrict__ tkeys,
                                 const int32_t *__restrict__ tvals, uint64_t mask, uint8_t *__restrict__ keep) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keep[i] = hash_find(tkeys, tvals, mask, keys[i]) == (int32_t)i;
}

__global__ void hash_contains_kernel(const uint64_t *__restrict__ q, int64_t n, const uint64_t *__restrict__ tkeys,
                                     uint64_t mask, uint8_t *__restrict__ found) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = q[i];
        uint64_t s = hash_slot(key, mask);
        uint8_t f = 0;
        while (true) {
            uint64_t k = __ldg(tkeys + s);
            if (k == key) { f = 1; break; }
            if (k == PCGC_EMPTY_KEY) break;
            s = (s + 1) & mask;
        }
        found[i] = f;
    }
}

// ---- k=3 kernel map ---------------------------------------------------------------------
// blockIdx.y = kernel offset k (x fastest); threads sweep rows, so keys are read and the
// offset-major map is written fully coalesced; the probes of adjacent (Morton-adjacent)
// rows fall into the same hash groups.
__global__ void kernel_map_k3_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                     const uint64_t *__restrict__ tkeys, const int32_t *__restrict__ tvals,
                                     uint64_t mask, int32_t *__restrict__ nbr,
                                     unsigned long long *__restrict__ n_pairs) {
    const int k = blockIdx.y;
    const int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
    int32_t *__restrict__ out = nbr + (int64_t)k * n;
    unsigned int local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t r;
        if (k == 13) {
            r = (int32_t)i;
        } else {
            uint32_t b, x, y, z;
            split_key(keys[i], b, x, y, z);
            const int qx = (int)x + dx, qy = (int)y + dy, qz = (int)z + dz;
            if ((qx | qy | qz) < 0 || qx > PCGC_MAX_COORD || qy > PCGC_MAX_COORD || qz > PCGC_MAX_COORD)
                r = -1;
            else
                r = hash_find(tkeys, tvals, mask, make_key(b, qx, qy, qz));
        }
        out[i] = r;
        local += (r >= 0);
    }
    if (n_pairs) {
        local = __reduce_add_sync(0xffffffffu, local);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_pairs, (unsigned long long)local);
    }
}

// ---- stride-2 down map --------------------------------------------------------------------
__global__ void iota_kernel(int32_t *__restrict__ v, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = (int32_t)i;
}

__global__ void parent_heads_kernel(const uint64_t *__restrict__ sorted_keys, int64_t n, int32_t *__restrict__ head) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        head[i] = (i == 0) || (parent_key(sorted_keys[i]) != parent_key(sorted_keys[i - 1]));
}

__global__ void parent_emit_kernel(const uint64_t *__restrict__ sorted_keys, int64_t n,
                                   const int32_t *__restrict__ head_scan /* inclusive */,
                                   const int32_t *__restrict__ child_rows, uint64_t *__restrict__ parent_keys,
                                   int32_t *__restrict__ child_off, int32_t *__restrict__ n_parents,
                                   int32_t *__restrict__ parent_of) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t incl = head_scan[i];
        if (parent_of) parent_of[child_rows[i]] = incl - 1;
        const bool is_head = (i == 0) || (incl != head_scan[i - 1]);
        if (is_head) {
            parent_keys[incl - 1] = parent_key(sorted_keys[i]);
            child_off[incl - 1] = (int32_t)i;
        }
        if (i == n - 1) {
            child_off[incl] = (int32_t)n;
            *n_parents = incl;
        }
    }
}

// info[p] = (first child row << 8) | 8-bit occupancy of the children of parent p (sorted children)
__global__ void parent_info_kernel(const uint64_t *__restrict__ child_keys, const int32_t *__restrict__ child_off,
                                   int64_t n_parents, uint64_t *__restrict__ info) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_parents; p += (int64_t)gridDim.x * blockDim.x) {
        const int beg = child_off[p], end = child_off[p + 1];
        uint32_t occ = 0;
        for (int j = beg; j < end; ++j) occ |= 1u << (uint32_t)(child_keys[j] & 7);
        info[p] = ((uint64_t)beg << 8) | occ;
    }
}

// k=3 kernel map of a child set derived from its PARENT set's kernel map: no hashing.  A child at
// position c in {0,1}^3 of parent p looks, for offset d in {-1,0,1}^3, into parent neighbour
// floor((c+d)/2) at child position (c+d)&1.  Children are sorted (Morton), so the row of child c'
// of parent p' is first_child(p') + popcount(occupancy(p') & ((1<<c')-1)); full-octet sets
// (generative up-sampling output) need no tables at all: row = 8*p' + c'.
__global__ void kernel_map_from_parent_kernel(const uint64_t *__restrict__ child_keys,
                                              const int32_t *__restrict__ parent_of,
                                              const uint64_t *__restrict__ info, const int32_t *__restrict__ pnbr,
                                              int64_t n_parents, int64_t n, int32_t *__restrict__ nbr) {
    const int k = blockIdx.y;
    const int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
    int32_t *__restrict__ out = nbr + (int64_t)k * n;
    const bool full = (info == nullptr);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t r;
        if (k == 13) {
            r = (int32_t)i;
        } else {
            const int c = full ? (int)(i & 7) : (int)(child_keys[i] & 7);
            const int64_t p = full ? (i >> 3) : (int64_t)parent_of[i];
            const int tx = (c & 1) + dx, ty = ((c >> 1) & 1) + dy, tz = ((c >> 2) & 1) + dz;   // in {-1,0,1,2}
            const int kp = ((tx >> 1) + 1) + 3 * ((ty >> 1) + 1) + 9 * ((tz >> 1) + 1);        // parent offset
            const int cc = (tx & 1) | ((ty & 1) << 1) | ((tz & 1) << 2);                       // child position there
            const int64_t q = kp == 13 ? p : (int64_t)__ldg(pnbr + (int64_t)kp * n_parents + p);
            if (q < 0) {
                r = -1;
            } else if (full) {
                r = (int32_t)(q * 8 + cc);
            } else {
                const uint64_t inf = __ldg(info + q);
                const uint32_t occ = (uint32_t)(inf & 0xFF);
                r = (occ >> cc) & 1 ? (int32_t)(inf >> 8) + __popc(occ & ((1u << cc) - 1)) : -1;
            }
        }
        out[i] = r;
    }
}

__global__ void upsample_keys_kernel(const uint64_t *__restrict__ keys, int64_t n8, uint64_t *__restrict__ child) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        child[i] = child_key(keys[i >> 3], (int)(i & 7));
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace pcgc

using namespace pcgc;

extern "C" {

int pcgc_pack_keys(const int32_t *coords, int64_t n, int32_t tensor_stride, uint64_t *keys, int32_t *err_flag,
                   void *stream) {
    PCGC_REQUIRE(n >= 0 && tensor_stride >= 1, "pcgc_pack_keys: bad n=%lld stride=%d", (long long)n, tensor_stride);
    if (n == 0) return PCGC_OK;
    PCGC_REQUIRE(((uintptr_t)coords & 15) == 0, "pcgc_pack_keys: coords must be 16-byte aligned");
    pack_keys_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const int4 *)coords, n, tensor_stride,
                                                                           keys, err_flag);
    return check_launch("pack_keys");
}

int pcgc_pack_keys3(const int32_t *coords3, int64_t n, int32_t tensor_stride, int32_t batch, int32_t hint_bits, uint64_t *keys,
                    int32_t *err_flag, void *stream) {
    PCGC_REQUIRE(n >= 0 && tensor_stride >= 1 && batch >= 0 && batch <= PCGC_MAX_BATCH && hint_bits >= 0 && hint_bits <= 19,
                 "pcgc_pack_keys3: bad arguments");
    if (n == 0) return PCGC_OK;
    pack_keys3_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(coords3, n, tensor_stride, batch, hint_bits, keys, err_flag);
    return check_launch("pack_keys3");
}

int pcgc_child_map_k2(const uint64_t *child_keys, const int32_t *parent_of, int64_t n, int64_t n_parents, int32_t *child_map,
                      void *stream) {
    PCGC_REQUIRE(n >= 0 && n_parents >= 0, "pcgc_child_map_k2: bad sizes");
    if (n_parents == 0) return PCGC_OK;
    PCGC_CUDA(cudaMemsetAsync(child_map, 0xFF, sizeof(int32_t) * 8 * (size_t)n_parents, (cudaStream_t)stream));   // -1 everywhere
    if (n == 0) return PCGC_OK;
    child_map_k2_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(child_keys, parent_of, n, n_parents, child_map);
    return check_launch("child_map_k2");
}

int pcgc_unpack_keys(const uint64_t *keys, int64_t n, int32_t tensor_stride, int32_t *coords, void *stream) {
    PCGC_REQUIRE(n >= 0 && tensor_stride >= 1, "pcgc_unpack_keys: bad arguments");
    if (n == 0) return PCGC_OK;
    PCGC_REQUIRE(((uintptr_t)coords & 15) == 0, "pcgc_unpack_keys: coords must be 16-byte aligned");
    unpack_keys_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(keys, n, tensor_stride, (int4 *)coords);
    return check_launch("unpack_keys");
}

int pcgc_scale_coords(const int32_t *coords, int64_t count, float factor, int32_t *out, void *stream) {
    PCGC_REQUIRE(count >= 0, "pcgc_scale_coords: bad count");
    if (count == 0) return PCGC_OK;
    scale_coords_kernel<<<grid_for(count, 256, 8), 256, 0, (cudaStream_t)stream>>>(coords, count, factor, out);
    return check_launch("scale_coords");
}

int64_t pcgc_hash_capacity(int64_t n) {
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

int pcgc_hash_build(const uint64_t *keys, int64_t n, uint64_t *table_keys, int32_t *table_vals, int64_t cap,
                    int32_t *n_dup, void *stream) {
    PCGC_REQUIRE(cap >= 64 && (cap & (cap - 1)) == 0 && cap >= 2 * n, "pcgc_hash_build: bad capacity %lld for n=%lld",
                 (long long)cap, (long long)n);
    PCGC_REQUIRE(n < 0x7FFFFFFF, "pcgc_hash_build: too many rows");
    cudaStream_t s = (cudaStream_t)stream;
    PCGC_CUDA(cudaMemsetAsync(n_dup, 0, sizeof(int32_t), s));
    hash_clear_kernel<<<grid_for(cap, 256, 8), 256, 0, s>>>(table_keys, table_vals, cap);
    int rc = check_launch("hash_clear");
    if (rc || n == 0) return rc;
    hash_insert_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(keys, n, table_keys, table_vals, (uint64_t)cap - 1, n_dup);
    return check_launch("hash_insert");
}

int pcgc_hash_keep_flags(const uint64_t *keys, int64_t n, const uint64_t *table_keys, const int32_t *table_vals,
                         int64_t cap, uint8_t *keep, void *stream) {
    if (n == 0) return PCGC_OK;
    hash_keep_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(keys, n, table_keys, table_vals,
                                                                           (uint64_t)cap - 1, keep);
    return check_launch("hash_keep");
}

int pcgc_hash_contains(const uint64_t *query, int64_t n, const uint64_t *table_keys, int64_t cap, uint8_t *found,
                       void *stream) {
    if (n == 0) return PCGC_OK;
    hash_contains_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(query, n, table_keys,
                                                                               (uint64_t)cap - 1, found);
    return check_launch("hash_contains");
}

int pcgc_kernel_map_k3(const uint64_t *keys, int64_t n, const uint64_t *table_keys, const int32_t *table_vals,
                       int64_t cap, int32_t *nbr, unsigned long long *n_pairs, void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF, "pcgc_kernel_map_k3: bad n");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_pairs) PCGC_CUDA(cudaMemsetAsync(n_pairs, 0, sizeof(unsigned long long), s));
    if (n == 0) return PCGC_OK;
    dim3 grid(grid_for(n, 256, 2), 27);
    kernel_map_k3_kernel<<<grid, 256, 0, s>>>(keys, n, table_keys, table_vals, (uint64_t)cap - 1, nbr, n_pairs);
    return check_launch("kernel_map_k3");
}

size_t pcgc_argsort_ws_bytes(int64_t n) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, n > 0 ? n : 1);
    return align256(cub_bytes) + align256(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
}

int pcgc_argsort_u64(const uint64_t *keys, int64_t n, int end_bit, uint64_t *keys_sorted, int32_t *order, void *ws,
                     size_t ws_bytes, void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF && end_bit >= 1 && end_bit <= 64, "pcgc_argsort_u64: bad arguments");
    if (n == 0) return PCGC_OK;
    if (ws_bytes < pcgc_argsort_ws_bytes(n)) {
        set_error("pcgc_argsort_u64: workspace too small");
        return PCGC_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *iota = (int32_t *)ws;
    void *cub_ws = (char *)ws + align256(sizeof(int32_t) * (size_t)n);
    size_t cub_bytes = ws_bytes - align256(sizeof(int32_t) * (size_t)n);
    iota_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(iota, n);
    int rc = check_launch("iota");
    if (rc) return rc;
    PCGC_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys, keys_sorted, iota, order, n, 0, end_bit, s));
    g_launches.fetch_add(4, std::memory_order_relaxed);
    return PCGC_OK;
}

size_t pcgc_stride_down_ws_bytes(int64_t n) {
    if (n < 1) n = 1;
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    //   sorted keys | head flags / scan | cub scan temp | argsort ws
    return align256(sizeof(uint64_t) * (size_t)n) + 2 * align256(sizeof(int32_t) * (size_t)n) + align256(scan_bytes) +
           pcgc_argsort_ws_bytes(n);
}

int pcgc_stride_down(const uint64_t *keys, int64_t n, int32_t keys_are_sorted, uint64_t *parent_keys,
                     int32_t *n_parents, int32_t *child_rows, int32_t *child_off, int32_t *parent_of, void *ws,
                     size_t ws_bytes, void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF, "pcgc_stride_down: bad n");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        PCGC_CUDA(cudaMemsetAsync(n_parents, 0, sizeof(int32_t), s));
        PCGC_CUDA(cudaMemsetAsync(child_off, 0, sizeof(int32_t), s));
        return PCGC_OK;
    }
    if (ws_bytes < pcgc_stride_down_ws_bytes(n)) {
        set_error("pcgc_stride_down: workspace too small");
        return PCGC_ERR_WORKSPACE;
    }
    char *p = (char *)ws;
    uint64_t *sorted = (uint64_t *)p;  p += align256(sizeof(uint64_t) * (size_t)n);
    int32_t *head = (int32_t *)p;      p += align256(sizeof(int32_t) * (size_t)n);
    int32_t *scan = (int32_t *)p;      p += align256(sizeof(int32_t) * (size_t)n);
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    void *scan_ws = p;                 p += align256(scan_bytes);
    void *sort_ws = p;
    const int g = grid_for(n, 256, 8);
    int rc;
    const uint64_t *sk = keys;
    if (keys_are_sorted) {
        iota_kernel<<<g, 256, 0, s>>>(child_rows, n);
        if ((rc = check_launch("iota"))) return rc;
    } else {
        if ((rc = pcgc_argsort_u64(keys, n, 64, sorted, child_rows, sort_ws, pcgc_argsort_ws_bytes(n), stream))) return rc;
        sk = sorted;
    }
    parent_heads_kernel<<<g, 256, 0, s>>>(sk, n, head);
    if ((rc = check_launch("parent_heads"))) return rc;
    PCGC_CUDA(cub::DeviceScan::InclusiveSum(scan_ws, scan_bytes, head, scan, n, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    parent_emit_kernel<<<g, 256, 0, s>>>(sk, n, scan, child_rows, parent_keys, child_off, n_parents, parent_of);
    return check_launch("parent_emit");
}

int pcgc_parent_info(const uint64_t *child_keys, const int32_t *child_off, int64_t n_parents, uint64_t *info,
                     void *stream) {
    PCGC_REQUIRE(n_parents >= 0, "pcgc_parent_info: bad n_parents");
    if (n_parents == 0) return PCGC_OK;
    parent_info_kernel<<<grid_for(n_parents, 256, 8), 256, 0, (cudaStream_t)stream>>>(child_keys, child_off, n_parents,
                                                                                     info);
    return check_launch("parent_info");
}

int pcgc_kernel_map_k3_from_parent(const uint64_t *child_keys, const int32_t *parent_of, const uint64_t *parent_info,
                                   const int32_t *parent_nbr, int64_t n_parents, int64_t n, int32_t *nbr,
                                   void *stream) {
    PCGC_REQUIRE(n >= 0 && n < 0x7FFFFFFF && n_parents >= 0, "pcgc_kernel_map_k3_from_parent: bad sizes");
    if (parent_info == nullptr)
        PCGC_REQUIRE(n == 8 * n_parents, "pcgc_kernel_map_k3_from_parent: full-octet mode needs n == 8 * n_parents");
    else
        PCGC_REQUIRE(child_keys && parent_of, "pcgc_kernel_map_k3_from_parent: null child tables");
    if (n == 0) return PCGC_OK;
    dim3 grid(grid_for(n, 256, 2), 27);
    kernel_map_from_parent_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(child_keys, parent_of, parent_info,
                                                                         parent_nbr, n_parents, n, nbr);
    return check_launch("kernel_map_from_parent");
}

int pcgc_upsample_keys(const uint64_t *keys, int64_t n, uint64_t *child_keys, void *stream) {
    PCGC_REQUIRE(n >= 0 && 8 * n < 0x7FFFFFFF, "pcgc_upsample_keys: bad n");
    if (n == 0) return PCGC_OK;
    upsample_keys_kernel<<<grid_for(8 * n, 256, 8), 256, 0, (cudaStream_t)stream>>>(keys, 8 * n, child_keys);
    return check_launch("upsample_keys");
}

}  // extern "C"
