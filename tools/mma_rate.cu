// Issue-rate micro-benchmark of the legacy tensor path on sm_100a: mma.sync m16n8k8 TF32 against m16n8k16
// F16 / BF16 (f32 accumulate), plus the float -> half2 pack conversion.  One CTA per SM, W warps per CTA, every warp
// runs ILP independent accumulator chains.  Prints cycles per warp-instruction per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int KIND, int ILP>
__global__ void rate_kernel(float *out, long long *cycles, int iters) {
    float d[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) d[i][e] = threadIdx.x * 1e-9f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 3)   // m16n8k8 f16 (half the K of 16816)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
            else {                // KIND 4: cvt.rn.f16x2.f32 pack, 4 per "instruction group"
                uint32_t p;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(p) : "f"(d[i][0]), "f"(d[i][1]));
                d[i][0] = __uint_as_float(p ^ b0);
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(p) : "f"(d[i][2]), "f"(d[i][3]));
                d[i][2] = __uint_as_float(p ^ b1);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND, int ILP>
static int run(const char *name, int warps, float *out, long long *cyc) {
    const int iters = 4096;
    rate_kernel<KIND, ILP><<<148, warps * 32>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    rate_kernel<KIND, ILP><<<148, warps * 32>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double per_smsp = (double)iters * ILP * (KIND == 4 ? 2 : 1) * warps / 4.0;   // warp-instructions per sub-partition
    printf("%-28s warps/CTA %2d ILP %d : %.2f cycles per warp-instruction per sub-partition\n", name, warps, ILP, avg / per_smsp);
    return 0;
}

int main() {
    float *out; long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    for (int warps : {4, 8, 16}) {
        if (warps == 4) { run<0, 4>("mma m16n8k8  tf32", 4, out, cyc); run<1, 4>("mma m16n8k16 f16", 4, out, cyc); run<2, 4>("mma m16n8k16 bf16", 4, out, cyc); run<3, 4>("mma m16n8k8  f16", 4, out, cyc); run<4, 4>("cvt.rn.f16x2.f32", 4, out, cyc); }
        if (warps == 8) { run<0, 4>("mma m16n8k8  tf32", 8, out, cyc); run<1, 4>("mma m16n8k16 f16", 8, out, cyc); run<2, 4>("mma m16n8k16 bf16", 8, out, cyc); run<3, 4>("mma m16n8k8  f16", 8, out, cyc); run<4, 4>("cvt.rn.f16x2.f32", 8, out, cyc); }
        if (warps == 16) { run<0, 4>("mma m16n8k8  tf32", 16, out, cyc); run<1, 4>("mma m16n8k16 f16", 16, out, cyc); run<2, 4>("mma m16n8k16 bf16", 16, out, cyc); run<3, 4>("mma m16n8k8  f16", 16, out, cyc); run<4, 4>("cvt.rn.f16x2.f32", 16, out, cyc); }
    }
    return 0;
}
