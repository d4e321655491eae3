// tc_probe_bf16.cu -- validates, on a B200, the building blocks of the tensor-core mel projection (srfe_fbank_tc.cuh)
// (dev tool; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tc_probe_bf16 tc_probe_bf16.cu):
//   * kind::f16 MMA, bf16 x bf16 -> f32, A operand (the mel weights, M = 128 filters) in TMEM as packed bf16 pairs
//     (column c of lane m = K elements 2c | 2c+1 in the low | high half), B operand (power spectra, N = frames) in shared
//     memory, K-major, in two layouts: no swizzle with a PADDED K stride -- element (n, k) at (n/8) SBO + (k/8) LBO +
//     (n%8) 16 + (k%8) 2 bytes, LBO = 144 (so that 16 lanes storing 16-byte chunks of 16 different K chunks hit distinct
//     banks) -- and SWIZZLE_128B (rows of 64 K elements, 16-byte chunk index XOR row % 8, 8-row groups 1024 B apart)
//   * the round-to-nearest hi + mid split of both operands with 3 products (hi hi + hi mid + mid hi): relative error
//     of an all-positive contraction against float64
//   * cycles per chain of 48 MMAs (K = 3 x 256) as a function of N (issue-bound: the single issuing thread)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, KP = 256, NMAX = 128;       // D[128 x N] = W[128 x 256] . P[N x 256]^T, each operand as hi + mid
constexpr int CH = KP / 8;                          // 16-byte chunks per part and row
constexpr int LBO = 144, SBO = 2 * CH * LBO;        // hi chunks 0 .. CH-1, mid chunks CH .. 2 CH - 1

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}
// 128-byte swizzle, K-major: rows of 64 bf16 (128 B), 8-row groups 1024 B apart, 16-byte chunk index XOR (row % 8)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void split_bf16(float x, unsigned short& h, unsigned short& m) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(x);
    const __nv_bfloat16 mb = __float2bfloat16_rn(x - __bfloat162float(hb));
    h = __bfloat16_as_ushort(hb);
    m = __bfloat16_as_ushort(mb);
}

__global__ void __launch_bounds__(128, 1) probe(const float* W, const float* P, float* D, int N, int reps, long long* cyc, int sw) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (NMAX / 8) * SBO / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    for (int i = tid; i < N * KP; i += 128) {
        const int n = i / KP, k = i % KP;
        unsigned short h, m;
        split_bf16(P[i], h, m);
        if (sw) {      // [part][K atom of 64][row][128 B], chunk ^ (row % 8)
            const int off = ((k / 64) * N + n) * 128 + ((((k % 64) / 8) ^ (n % 8)) * 16) + (k % 8) * 2;
            *reinterpret_cast<unsigned short*>(smem + off) = h;
            *reinterpret_cast<unsigned short*>(smem + off + 4 * N * 128) = m;
        } else {
            const int off = (n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2;
            *reinterpret_cast<unsigned short*>(smem + off) = h;
            *reinterpret_cast<unsigned short*>(smem + off + CH * LBO) = m;
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
    // A: thread = filter row; columns [0, KP/2) hi pairs, [KP/2, KP) mid pairs
    for (int c = 0; c < KP / 2; c += 8) {
        uint32_t h[8], l[8];
        for (int j = 0; j < 8; ++j) {
            unsigned short h0, m0, h1, m1;
            split_bf16(W[tid * KP + 2 * (c + j)], h0, m0);
            split_bf16(W[tid * KP + 2 * (c + j) + 1], h1, m1);
            h[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            l[j] = (uint32_t)m0 | ((uint32_t)m1 << 16);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(lane_base + c), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(lane_base + KP / 2 + c), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t d_col = tb + KP;                              // D columns [272, 272 + N)
    if (tid == 0) {
        // instruction descriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3, M >> 4
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            uint32_t acc = 0;
            for (int pr = 0; pr < 3; ++pr) {                     // W_hi P_hi + W_hi P_mid + W_mid P_hi
                const uint32_t a_col = tb + (pr == 2 ? KP / 2 : 0);
                const uint32_t b_base = smem_u32(smem) + (pr == 1 ? CH * LBO : 0);
                for (int ks = 0; ks < KP / 16; ++ks) {
                    const uint64_t bdesc = sw ? make_desc_sw128(smem_u32(smem) + ((pr == 1 ? 4 : 0) + ks / 4) * N * 128 + (ks % 4) * 32)
                                              : make_desc(b_base + ks * 2 * LBO);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 :: "r"(d_col), "r"(a_col + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc));
                    acc = 1;
                }
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)));
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
        }
        const long long t2 = clock64();
        cyc[0] = t1 - t0;
        cyc[1] = t2 - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(lane_base + KP + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) D[tid * NMAX + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}

int main() {
    std::vector<float> W(M * KP), P(NMAX * KP), D(M * NMAX);
    float *dW, *dP, *dD;
    long long* dC;
    CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dC, 16));
    const int smem = (NMAX / 8) * SBO;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    srand(4321);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < KP; ++k) W[m * KP + k] = (abs(k - 2 * m) < 6) ? 0.02f * (float)rand() / RAND_MAX : 0.f;   // banded, positive
    for (auto& p : P) p = expf(30.f * (float)rand() / RAND_MAX);                                                     // 13 decades
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
    int fails = 0;
    const int Ns[5] = {32, 48, 64, 96, 128};
    for (int t = 0; t < 5; ++t) {
        const int N = Ns[t];
        for (int sw = 0; sw < 2; ++sw)
        for (int reps = 1; reps <= 4; reps *= 4) {
            CK(cudaMemset(dD, 0xff, D.size() * 4));
            probe<<<1, 128, smem>>>(dW, dP, dD, N, reps, dC, sw);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            long long c[2];
            CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double rel = 0;
            int bad = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double s = 0;
                    for (int k = 0; k < KP; ++k) s += (double)W[m * KP + k] * P[n * KP + k];
                    const double d = D[m * NMAX + n];
                    if (s == 0) { if (d != 0) ++bad; continue; }
                    rel = fmax(rel, fabs(d - s) / s);
                }
            printf("%s N = %3d reps = %d: max rel err %.3e (exact-zero rows wrong: %d)   issue %lld cycles, issue -> complete %lld cycles (%.0f per MMA)\n",
                   sw ? "sw128     " : "interleave", N, reps, rel, bad, c[0], c[1], (double)c[1] / (48.0 * reps));
            if (reps == 1 && (rel > 1e-4 || bad)) { printf("  FAIL\n"); ++fails; }
        }
    }
    printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
    return fails;
}
