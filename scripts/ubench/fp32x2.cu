// Microbenchmark: scalar FADD/FFMA vs packed f32x2 (sm_100a) issue/throughput.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    constexpr int C = 8;   // independent chains (x2 lanes when packed)
    float a[2 * C];
    uint64_t p[C];
    for (int i = 0; i < 2 * C; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < C; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const uint64_t ps = pk(s, s), pt = pk(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) { for (int i = 0; i < 2 * C; ++i) a[i] = fmaf(a[i], s, 0.5f); }          // 16 FFMA
            if (MODE == 1) { for (int i = 0; i < 2 * C; ++i) a[i] = a[i] + s; }                     // 16 FADD
            if (MODE == 2) { for (int i = 0; i < C; ++i) p[i] = fma2(p[i], ps, pt); }               // 8 FFMA2
            if (MODE == 3) { for (int i = 0; i < C; ++i) p[i] = add2(p[i], ps); }                   // 8 FADD2
            if (MODE == 4) { for (int i = 0; i < C; ++i) p[i] = mul2(p[i], ps); }                   // 8 FMUL2
            if (MODE == 5) { for (int i = 0; i < C; ++i) { p[i] = fma2(p[i], ps, pt); a[2*i] = fmaf(a[2*i], s, 0.5f); a[2*i+1] = a[2*i+1] + s; } }  // mix
        }
    }
    float acc = 0;
    for (int i = 0; i < 2 * C; ++i) acc += a[i];
    for (int i = 0; i < C; ++i) { float x, y; upk(p[i], x, y); acc += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char* name, int flops_per_inner, float* out) {
    const int iters = 20000, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(out, 100, 1.0001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * 256 * iters * 4 * flops_per_inner;   // fp lane-ops (an FMA counts 1)
    printf("%-8s %8.3f ms  %7.2f T lane-op/s  (%.1f lane-op/clk/SM @1.9GHz)\n", name, ms, lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FFMA", 16, out); run<1>("FADD", 16, out); run<2>("FFMA2", 16, out); run<3>("FADD2", 16, out); run<4>("FMUL2", 16, out); run<5>("MIX", 32, out);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
