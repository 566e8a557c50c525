// metrics.cu -- D1 (point-to-point) geometry distortion on the GPU.
//
// Replaces the pc_error_d subprocess of the reference (pc_error.py:44-54, read back at coder.py:181-184: the
// `mseF,PSNR (p2point)` line): for every voxel of cloud A the squared distance to its nearest voxel of cloud B, summed
// exactly in integers, in both directions.  Both clouds are integer voxel sets that already live on the device as
// Morton keys with a hash table (the same table the kernel maps use), so the nearest neighbour is found by probing
// growing cubes around the query: an exact hit costs one probe (the common case for a decoded cloud), a miss scans the
// shell of Chebyshev radius R = 1, 2, ... and stops as soon as the best squared distance is <= (R + 1)^2 (every
// voxel not yet seen is farther than R + 1).  Queries still open after `max_radius` are finished by a brute-force
// pass over B (one block per query), so the result is exact for any pair of clouds.
#include "common.cuh"

namespace pcgc {

__device__ __forceinline__ bool hash_has(const uint64_t *__restrict__ tkeys, uint64_t mask, uint64_t key) {
    uint64_t s = hash_slot(key, mask);
    while (true) {
        const uint64_t k = __ldg(tkeys + s);
        if (k == key) return true;
        if (k == PCGC_EMPTY_KEY) return false;
        s = (s + 1) & mask;
    }
}

// acc[0] = sum of squared NN distances, acc[1] = max squared NN distance, acc[2] = number of open queries
__global__ void d1_probe_kernel(const uint64_t *__restrict__ qkeys, int64_t nq, const uint64_t *__restrict__ tkeys, uint64_t mask,
                                int max_radius, unsigned long long *__restrict__ acc, int32_t *__restrict__ open_list) {
    unsigned long long sum = 0, mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = qkeys[i];
        if (hash_has(tkeys, mask, key)) continue;                          // distance 0
        uint32_t b, x, y, z;
        split_key(key, b, x, y, z);
        uint32_t best = 0xFFFFFFFFu;
        bool done = false;
        for (int R = 1; R <= max_radius && !done; ++R) {
            for (int dz = -R; dz <= R; ++dz)
                for (int dy = -R; dy <= R; ++dy) {
                    const bool face = dz == -R || dz == R || dy == -R || dy == R;
                    for (int dx = -R; dx <= R; dx += face ? 1 : 2 * R) {   // the shell only: interior columns touch dx = +-R
                        const int nx = (int)x + dx, ny = (int)y + dy, nz = (int)z + dz;
                        if ((nx | ny | nz) < 0 || nx > PCGC_MAX_COORD || ny > PCGC_MAX_COORD || nz > PCGC_MAX_COORD) continue;
                        const uint32_t d2 = (uint32_t)(dx * dx + dy * dy + dz * dz);
                        if (d2 < best && hash_has(tkeys, mask, make_key(b, (uint32_t)nx, (uint32_t)ny, (uint32_t)nz))) best = d2;
                    }
                }
            done = best <= (uint32_t)((R + 1) * (R + 1));
        }
        if (done) {
            sum += best;
            mx = mx > best ? mx : best;
        } else {
            open_list[atomicAdd(acc + 2, 1ull)] = (int32_t)i;
        }
    }
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = mx > m2 ? mx : m2;
    }
    if ((threadIdx.x & 31) == 0) {
        if (sum) atomicAdd(acc, sum);
        if (mx) atomicMax(acc + 1, mx);
    }
}

// one block per open query: exact minimum over all of B
__global__ void d1_brute_kernel(const uint64_t *__restrict__ qkeys, const int32_t *__restrict__ open_list, const uint64_t *__restrict__ bkeys,
                                int64_t nb, unsigned long long *__restrict__ acc) {
    __shared__ unsigned long long red[32];
    const unsigned long long n_open = acc[2];
    for (unsigned long long q = blockIdx.x; q < n_open; q += gridDim.x) {
        uint32_t b, x, y, z;
        split_key(qkeys[open_list[q]], b, x, y, z);
        unsigned long long best = ~0ull;
        for (int64_t j = threadIdx.x; j < nb; j += blockDim.x) {
            uint32_t b2, x2, y2, z2;
            split_key(__ldg(bkeys + j), b2, x2, y2, z2);
            if (b2 != b) continue;
            const long long dx = (long long)x - x2, dy = (long long)y - y2, dz = (long long)z - z2;
            const unsigned long long d2 = (unsigned long long)(dx * dx + dy * dy + dz * dz);
            best = d2 < best ? d2 : best;
        }
        for (int o = 16; o; o >>= 1) {
            const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o);
            best = v < best ? v : best;
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = red[w] < best ? red[w] : best;
            if (best != ~0ull) {
                atomicAdd(acc, best);
                atomicMax(acc + 1, best);
            }
        }
        __syncthreads();
    }
}

}  // namespace pcgc

using namespace pcgc;

extern "C" {

int pcgc_d1_sqdist(const uint64_t *query_keys, int64_t n_query, const uint64_t *table_keys, int64_t cap, const uint64_t *cloud_keys,
                   int64_t n_cloud, int32_t max_radius, uint64_t *acc3, int32_t *open_list, void *stream) {
    PCGC_REQUIRE(n_query >= 0 && n_cloud >= 0 && cap >= 64 && (cap & (cap - 1)) == 0 && max_radius >= 1 && max_radius <= 16,
                 "pcgc_d1_sqdist: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    PCGC_CUDA(cudaMemsetAsync(acc3, 0, 3 * sizeof(uint64_t), s));
    if (n_query == 0 || n_cloud == 0) return PCGC_OK;
    d1_probe_kernel<<<grid_for(n_query, 256, 8), 256, 0, s>>>(query_keys, n_query, table_keys, (uint64_t)cap - 1, max_radius,
                                                            (unsigned long long *)acc3, open_list);
    int rc = check_launch("d1_probe");
    if (rc) return rc;
    d1_brute_kernel<<<2 * kNumSMs, 256, 0, s>>>(query_keys, open_list, cloud_keys, n_cloud, (unsigned long long *)acc3);
    return check_launch("d1_brute");
}

}  // extern "C"
