// Microbenchmark (sm_100a): does the legacy mma.sync TF32 path run beside packed FP32 (FFMA2) work, or do they share
// a pipe?  Half of the warps of every CTA loop on FFMA2, the other half on mma.sync.m16n8k8 TF32; each kind is also
// timed alone (the other half idle).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// mode bit 0: FFMA2 warps active, bit 1: MMA warps active
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool is_mma = warp & 1;
    float acc = 0.f;
    const long long t0 = clock64();
    if (!is_mma && (mode & 1)) {
        uint64_t p[8]; for (int i = 0; i < 8; ++i) p[i] = (uint64_t)(threadIdx.x + i);
        const uint64_t s = 0x3f8000013f800001ull, t = 0x3a0000003a000000ull;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], s, t);
        }
        for (int i = 0; i < 8; ++i) acc += (float)p[i];
    }
    if (is_mma && (mode & 2)) {
        float c[4][4] = {};
        uint32_t a[4] = {threadIdx.x, threadIdx.x + 1u, 0x3f800000u, 0x3f000000u};
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) mma_tf32(c[u], a, 0x3f800000u, 0x3e800000u);      // 4 independent accumulators
        }
        for (int u = 0; u < 4; ++u) acc += c[u][0] + c[u][3];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
    const int iters = 4000;
    for (int mode = 1; mode <= 3; ++mode) {
        k<<<148, 512>>>(out, cyc, 10, mode);
        k<<<148, 512>>>(out, cyc, iters, mode);
        cudaDeviceSynchronize();
        static long long h[148 * 16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double f = 0, m = 0;
        for (int b = 0; b < 148; ++b) for (int w = 0; w < 16; ++w) { if (w & 1) m += h[b * 16 + w]; else f += h[b * 16 + w]; }
        f /= 148 * 8; m /= 148 * 8;
        // per sub-partition: 2 FFMA2 warps x 32 FFMA2 / iter, 2 MMA warps x 4 HMMA / iter
        printf("mode %d (%s%s): FFMA2 warps %.2f cyc per FFMA2 per sub-partition | MMA warps %.2f cyc per HMMA.1688 per sub-partition\n",
               mode, (mode & 1) ? "FFMA2 " : "", (mode & 2) ? "MMA" : "",
               (mode & 1) ? f / (iters * 32.0 * 2) : 0.0, (mode & 2) ? m / (iters * 4.0 * 2) : 0.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
