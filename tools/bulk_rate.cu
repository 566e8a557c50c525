// bulk_rate.cu -- how many small cp.async.bulk (global -> shared, 64 / 128 / 512 bytes) can one SM retire per cycle?
// Decides whether the 4x4x4 halo of the full-octet kernels (48 contiguous runs of 64-128 bytes per octet) can be staged
// by the bulk-copy engine instead of 256 16-byte LDGSTS per octet.  One CTA per SM, WARPS warps, every lane issues
// COPIES copies per round into its warp's buffer and the warp waits on its own mbarrier.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES, int PER_LANE>
__global__ void __launch_bounds__(256) bulk_kernel(const char *__restrict__ src, size_t src_bytes, int rounds, unsigned long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = smem + (size_t)warp * 32 * PER_LANE * BYTES;
    const uint32_t bar = smem_u32(&bars[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(bar));
    __syncthreads();
    uint32_t phase = 0;
    uint64_t rng = (blockIdx.x * 256 + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PER_LANE * BYTES) : "memory");
#pragma unroll
        for (int c = 0; c < PER_LANE; ++c) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const size_t off = ((rng >> 20) % (src_bytes / 512)) * 512 + ((rng >> 12) & 3) * 128;      // 24 MB window: L2 resident
            const uint32_t dst = smem_u32(buf + (size_t)(c * 32 + lane) * BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src + off), "r"(BYTES), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        }
        phase ^= 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int BYTES, int PER_LANE>
__global__ void __launch_bounds__(256) ldgsts_kernel(const char *__restrict__ src, size_t src_bytes, int rounds, unsigned long long *cycles) {
    // the same bytes by 16-byte cp.async: BYTES / 16 adjacent lanes per run
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = smem + (size_t)warp * 32 * PER_LANE * BYTES;
    constexpr int LPR = BYTES / 16;                  // lanes per run
    uint64_t rng0 = (blockIdx.x * 8 + warp) * 0x9E3779B97F4A7C15ull + 777;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int c = 0; c < PER_LANE * LPR; ++c) {   // PER_LANE*32 runs per warp-round = PER_LANE*32*LPR lane-ops = PER_LANE*LPR instrs
            const int run = c * (32 / LPR) + lane / LPR;
            uint64_t h = (rng0 + run + (uint64_t)r * 4096) * 0xD6E8FEB86659FD93ull;
            h ^= h >> 32;
            const size_t off = ((h >> 8) % (src_bytes / 512)) * 512 + ((h >> 3) & 3) * 128 + (lane % LPR) * 16;
            const uint32_t dst = smem_u32(buf + (size_t)run * BYTES + (lane % LPR) * 16);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + off) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <typename K>
void run(const char *name, K kernel, int bytes, int per_lane, const char *src, size_t src_bytes, unsigned long long *cyc) {
    const int rounds = 200;
    const size_t smem = (size_t)8 * 32 * per_lane * bytes;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<148, 256, smem>>>(src, src_bytes, 10, cyc);
    kernel<<<148, 256, smem>>>(src, src_bytes, rounds, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double runs = (double)rounds * 8 * 32 * per_lane;
    printf("%-8s %4d B x %d/lane: %8.0f cycles, %.2f cycles per run per SM, %.1f B/cycle/SM\n", name, bytes, per_lane, avg, avg / runs,
           runs * bytes / avg);
}

int main() {
    const size_t src_bytes = 24u << 20;
    char *src;
    unsigned long long *cyc;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    cudaMalloc(&cyc, 148 * sizeof(unsigned long long));
    run("bulk", bulk_kernel<64, 4>, 64, 4, src, src_bytes, cyc);
    run("bulk", bulk_kernel<64, 8>, 64, 8, src, src_bytes, cyc);
    run("bulk", bulk_kernel<128, 4>, 128, 4, src, src_bytes, cyc);
    run("bulk", bulk_kernel<512, 1>, 512, 1, src, src_bytes, cyc);
    run("bulk", bulk_kernel<16, 8>, 16, 8, src, src_bytes, cyc);
    run("ldgsts", ldgsts_kernel<64, 4>, 64, 4, src, src_bytes, cyc);
    run("ldgsts", ldgsts_kernel<64, 8>, 64, 8, src, src_bytes, cyc);
    run("ldgsts", ldgsts_kernel<128, 4>, 128, 4, src, src_bytes, cyc);
    run("ldgsts", ldgsts_kernel<512, 1>, 512, 1, src, src_bytes, cyc);
    return 0;
}
