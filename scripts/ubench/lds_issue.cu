// Microbenchmark (sm_100a): shared-memory load cost by access pattern, and whether other instructions issue in the
// shadow of packed FFMA2.  Reports cycles per warp-instruction per SM sub-partition (4 per SM), clock64 based.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ float4 lds128(const float* p) { float4 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds64(const float* p) { float2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(const float* p) { float v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(float* p, float4 v) { unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
__device__ __forceinline__ void sts64(float* p, float2 v) { unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// MODE 0: LDS.128 one address per warp      1: LDS.128 two addresses per warp (half-warps)   2: LDS.128 lane-distinct
//      3: LDS.64 one address per warp       4: LDS.64 lane-distinct                            5: LDS.32 one address
//      6: 8 FFMA2                           7: 8 FFMA2 + 8 IADD3/LOP                           8: 8 FFMA2 + 4 LDS.64 distinct
//      9: 5 FFMA2 (.F32 broadcast of LDS.128 values) + LDS.64 distinct + 2 LDS.128 uniform  (the DCT item loop)
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* sf = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sf[i] = i * 1e-6f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t p[8]; for (int i = 0; i < 8; ++i) p[i] = (uint64_t)(threadIdx.x + i);
    float acc = 0.f; unsigned ia = threadIdx.x, ib = 17;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int base = (it * 64 + warp * 16) & 4095;      // floats
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { float4 v = lds128(sf + base + 4 * u); acc += v.x + v.w; }
            if (MODE == 1) { float4 v = lds128(sf + base + 4 * u + 64 * (lane >> 4)); acc += v.x + v.w; }
            if (MODE == 2) { float4 v = lds128(sf + ((base + 4 * lane + 128 * u) & 8191)); acc += v.x + v.w; }
            if (MODE == 3) { float2 v = lds64(sf + base + 2 * u); acc += v.x + v.y; }
            if (MODE == 4) { float2 v = lds64(sf + ((base + 2 * lane + 64 * u) & 8191)); acc += v.x + v.y; }
            if (MODE == 5) { acc += lds32(sf + base + u); }
            if (MODE == 10) { sts128(sf + ((base + 4 * lane + 128 * u) & 8191), make_float4(acc, acc, acc, acc)); }
            if (MODE == 11) { sts64(sf + ((base + 2 * lane + 64 * u) & 8191), make_float2(acc, acc)); }
            // FFT exchange read: lane l of a half-warp reads record (17 l + u) -- 16-byte records vs two 8-byte planes
            if (MODE == 12) { float4 v = lds128(sf + 4 * ((lane & 15) * 17 + u) + 1088 * (lane >> 4) + ((warp * 2176 + (it & 1) * 1088) & 4095)); acc += v.x + v.y + v.z + v.w; }
            if (MODE == 13) { const float* b = sf + 2 * ((lane & 15) * 17 + u) + 1088 * (lane >> 4) + ((warp * 2176 + (it & 1) * 1088) & 4095);
                              float2 v = lds64(b), w = lds64(b + 544); acc += v.x + v.y + w.x + w.y; }
            if (MODE == 14) { float* b = sf + 4 * ((lane & 15) + 17 * u) + 1088 * (lane >> 4) + ((warp * 2176 + (it & 1) * 1088) & 4095); sts128(b, make_float4(acc, acc, acc, acc)); }
            if (MODE == 15) { float* b = sf + 2 * ((lane & 15) + 17 * u) + 1088 * (lane >> 4) + ((warp * 2176 + (it & 1) * 1088) & 4095); sts64(b, make_float2(acc, acc)); sts64(b + 544, make_float2(acc, acc)); }
        }
        if (MODE == 6 || MODE == 7 || MODE == 8) {
            const uint64_t s = 0x3f8000013f800001ull, t = 0x3a0000003a000000ull;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], s, t);
                if (MODE == 7) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ia = ia * 3u + ib; ib = (ib ^ ia) + 5u; }
                }
                if (MODE == 8) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { float2 v = lds64(sf + ((base + 2 * lane + 64 * (4 * u + i)) & 8191)); acc += v.x; }
                }
            }
        }
        if (MODE == 9) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 xf = lds64(sf + ((base + 2 * lane * 129 + 2 * u) & 8191)); uint64_t x; asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(xf.x), "f"(xf.y));
                const float4 d0 = lds128(sf + ((base + 8 * u) & 4095));
                const float4 d1 = lds128(sf + ((base + 8 * u + 4) & 4095));
                uint64_t b0, b1, b2, b3, b4;
                asm("mov.b64 %0, {%1,%1};" : "=l"(b0) : "f"(d0.x)); asm("mov.b64 %0, {%1,%1};" : "=l"(b1) : "f"(d0.y));
                asm("mov.b64 %0, {%1,%1};" : "=l"(b2) : "f"(d0.z)); asm("mov.b64 %0, {%1,%1};" : "=l"(b3) : "f"(d0.w));
                asm("mov.b64 %0, {%1,%1};" : "=l"(b4) : "f"(d1.x));
                p[0] = fma2(x, b0, p[0]); p[1] = fma2(x, b1, p[1]); p[2] = fma2(x, b2, p[2]); p[3] = fma2(x, b3, p[3]); p[4] = fma2(x, b4, p[4]);
            }
        }
    }
    const long long t1 = clock64();
    for (int i = 0; i < 8; ++i) acc += (float)p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + ia + ib;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int instr_per_iter, int warps, float* out, long long* cyc) {
    const int iters = 4000, blocks = 148;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<MODE><<<blocks, 32 * warps, 65536>>>(out, cyc, 10);
    k<MODE><<<blocks, 32 * warps, 65536>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double wi = (double)iters * instr_per_iter * warps;     // warp-instructions of interest per SM
    printf("%-34s warps %2d  %7.2f cyc / warp-instr / SM   (%.2f per sub-partition)\n", name, warps, avg / wi, 4 * avg / wi);
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int w : {4, 16}) {
        run<0>("LDS.128 uniform", 8, w, out, cyc); run<1>("LDS.128 2 addr/warp", 8, w, out, cyc); run<2>("LDS.128 distinct", 8, w, out, cyc);
        run<3>("LDS.64 uniform", 8, w, out, cyc); run<4>("LDS.64 distinct", 8, w, out, cyc); run<5>("LDS.32 uniform", 8, w, out, cyc);
        run<6>("FFMA2 x32", 32, w, out, cyc); run<7>("FFMA2 x32 (+16 int, not counted)", 32, w, out, cyc);
        run<8>("FFMA2 x32 (+16 LDS.64, not counted)", 32, w, out, cyc); run<9>("DCT item loop: per f-step", 8, w, out, cyc);
        run<10>("STS.128 distinct", 8, w, out, cyc); run<11>("STS.64 distinct", 8, w, out, cyc);
        run<12>("FFT xchg read: LDS.128 (16 B rec)", 8, w, out, cyc); run<13>("FFT xchg read: 2 LDS.64 (planes)", 8, w, out, cyc);
        run<14>("FFT xchg write: STS.128", 8, w, out, cyc); run<15>("FFT xchg write: 2 STS.64 (planes)", 8, w, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
