// common.cuh -- shared device helpers of libpcgc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "../../include/pcgc.h"

namespace pcgc {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PCGC_ERR_CUDA;
    }
    return PCGC_OK;
}

#define PCGC_CUDA(expr)                                                         \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            pcgc::set_error("%s: %s", #expr, cudaGetErrorString(_e));           \
            return PCGC_ERR_CUDA;                                               \
        }                                                                       \
    } while (0)

#define PCGC_REQUIRE(cond, ...)                                                 \
    do {                                                                        \
        if (!(cond)) {                                                          \
            pcgc::set_error(__VA_ARGS__);                                       \
            return PCGC_ERR_INVALID;                                            \
        }                                                                       \
    } while (0)

// grid for a grid-stride kernel over `work` items with `per_block` items per block per
// iteration: enough blocks to cover the work, capped at waves * 148 resident CTAs.
inline int grid_for(int64_t work, int per_block, int ctas_per_sm) {
    int64_t need = (work + per_block - 1) / per_block;
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- Morton keys ------------------------------------------------------------------------
// key = (batch << 57) | interleave(x, y, z), x in bit 0 of each triple, 19 bits per axis.
constexpr int kCoordBits = 19;
constexpr uint64_t kMortonMask = (1ull << (3 * kCoordBits)) - 1;

__host__ __device__ __forceinline__ uint64_t spread3(uint32_t v) {
    uint64_t x = v & 0x1FFFFFull;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

__host__ __device__ __forceinline__ uint32_t compact3(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10C30C30C30C30C3ull;
    x = (x ^ (x >> 4)) & 0x100F00F00F00F00Full;
    x = (x ^ (x >> 8)) & 0x1F0000FF0000FFull;
    x = (x ^ (x >> 16)) & 0x1F00000000FFFFull;
    x = (x ^ (x >> 32)) & 0x1FFFFFull;
    return (uint32_t)x;
}

__host__ __device__ __forceinline__ uint64_t make_key(uint32_t b, uint32_t x, uint32_t y, uint32_t z) {
    return ((uint64_t)b << 57) | spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

__host__ __device__ __forceinline__ void split_key(uint64_t key, uint32_t &b, uint32_t &x, uint32_t &y,
                                                   uint32_t &z) {
    b = (uint32_t)(key >> 57);
    const uint64_t m = key & kMortonMask;       // the batch bits must not leak into the de-interleave
    x = compact3(m);
    y = compact3(m >> 1);
    z = compact3(m >> 2);
}

// one level up / down the octree: the batch bits stay in place, only the Morton field shifts
__host__ __device__ __forceinline__ uint64_t parent_key(uint64_t key) {
    return (key & ~kMortonMask) | ((key & kMortonMask) >> 3);
}
__host__ __device__ __forceinline__ uint64_t child_key(uint64_t key, int k) {
    return (key & ~kMortonMask) | (((key & kMortonMask) << 3) & kMortonMask) | (uint64_t)k;
}

// ---- hash table ---------------------------------------------------------------------------
// Open addressing, linear probing.  The low 3 key bits (position inside one 2x2x2 octet) are
// kept as the low slot bits so the 8 siblings of an octet share one 64-byte line of the key
// array (the probes of adjacent voxels hit the same lines); the octet index is mixed.  (Keeping
// 6 bits -- a 4x4x4 cell per 64-slot group -- overloads the groups: measured 0.9 ms per map.)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

__host__ __device__ __forceinline__ uint64_t hash_slot(uint64_t key, uint64_t mask) {
    return ((mix64(key >> 3) << 3) | (key & 7)) & mask;
}

__device__ __forceinline__ int32_t hash_find(const uint64_t *__restrict__ tkeys,
                                             const int32_t *__restrict__ tvals, uint64_t mask,
                                             uint64_t key) {
    uint64_t s = hash_slot(key, mask);
    while (true) {
        uint64_t k = __ldg(tkeys + s);
        if (k == key) return __ldg(tvals + s);
        if (k == PCGC_EMPTY_KEY) return -1;
        s = (s + 1) & mask;
    }
}

}  // namespace pcgc
