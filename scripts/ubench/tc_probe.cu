// tc_probe.cu -- validates, on a B200, every tcgen05 / TMEM building block the MFCC epilogue kernel relies on
// (dev tool; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu):
//   * TMEM allocation, tcgen05.st / tcgen05.ld in the 32x32b shape (thread = TMEM lane, register j = column j)
//   * kind::tf32 MMA with the A operand in TMEM (row = lane, K index = column) and B in shared memory,
//     K-major, no swizzle: element (n, k) at (n/8) SBO + (k/4) LBO + (n%8) 16 + (k%4) 4 bytes
//   * tcgen05.commit -> mbarrier, fences, D read-back: D[row][col] = sum_k A[row][k] B[col][k]
//   * the 3-product split (hi hi + lo hi + hi lo) that gives fp32-grade accuracy, and whether kind::tf32 truncates
//     or rounds the 13 low mantissa bits of its fp32 inputs
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, K = 64, N = 48;       // D[128 x 48] = A[128 x 64] . B[48 x 64]^T
constexpr int LBO = 128, SBO = (K / 4) * 128;   // core matrix = 8 rows x 16 B = 128 B; K-direction core matrices adjacent

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)(LBO >> 4) << 16;
    d |= (uint64_t)(SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    return d;                                  // base offset 0, no swizzle
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, int split) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* sBh = reinterpret_cast<float*>(smem);                 // B hi, UMMA layout
    float* sBl = sBh + N * K;                                    // B lo
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float b = B[i];
        const float bh = split ? __uint_as_float(__float_as_uint(b) & 0xffffe000u) : b;
        const int off = (n / 8) * SBO + (k / 4) * LBO + (n % 8) * 16 + (k % 4) * 4;
        sBh[off / 4] = bh;
        sBl[off / 4] = b - bh;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256));
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
    // A: thread = row, columns [0, K) hi, [K, 2K) lo
    for (int c = 0; c < K; c += 8) {
        uint32_t h[8], l[8];
        for (int j = 0; j < 8; ++j) {
            const float a = A[tid * K + c + j];
            const float ah = split ? __uint_as_float(__float_as_uint(a) & 0xffffe000u) : a;
            h[j] = __float_as_uint(ah);
            l[j] = __float_as_uint(a - ah);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(lane_base + c), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(lane_base + K + c), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("fence.proxy.async.shared::cta;");              // generic-proxy smem writes -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t d_col = tb + 2 * K;                           // D columns [2K, 2K + N)
    if (tid == 0) {
        // instruction descriptor: F32 accumulate, TF32 x TF32, K-major both, N >> 3, M >> 4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const int nprod = split ? 3 : 1;
        int first = 1;
        for (int pr = 0; pr < nprod; ++pr) {
            const uint32_t a_col = tb + (pr == 1 ? K : 0);       // 0: Ah Bh, 1: Al Bh, 2: Ah Bl
            const float* bsrc = (pr == 2) ? sBl : sBh;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t bdesc = make_desc(smem_u32(bsrc) + ks * 2 * LBO);
                const uint32_t acc = first ? 0u : 1u;
                first = 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             :: "r"(d_col), "r"(a_col + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)));
    }
    {   // everyone waits for the MMAs (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(lane_base + 2 * K + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(256));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    const int smem = 2 * N * K * 4;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int fails = 0;
    for (int test = 0; test < 3; ++test) {
        srand(1234 + test);
        for (auto& a : A) a = test == 0 ? (float)((rand() % 33) - 16) / 8.f : 80.f * ((float)rand() / RAND_MAX - 0.5f);
        for (auto& b : B) b = test == 0 ? (float)((rand() % 17) - 8) / 4.f : 0.4f * ((float)rand() / RAND_MAX - 0.5f);
        const int split = test == 2;
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, D.size() * 4));
        probe<<<1, 128, smem>>>(dA, dB, dD, split);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double e_exact = 0, e_trunc = 0, e_rna = 0, ref_max = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0, st = 0, sr = 0;
                for (int k = 0; k < K; ++k) {
                    s += (double)A[m * K + k] * B[n * K + k];
                    st += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[n * K + k]);
                    sr += (double)tf32_rna(A[m * K + k]) * tf32_rna(B[n * K + k]);
                }
                const double d = D[m * N + n];
                e_exact = fmax(e_exact, fabs(d - s)); e_trunc = fmax(e_trunc, fabs(d - st)); e_rna = fmax(e_rna, fabs(d - sr));
                ref_max = fmax(ref_max, fabs(s));
            }
        const char* names[3] = {"exact-in-tf32 inputs, 1 product", "random fp32 inputs, 1 product", "random fp32 inputs, 3-product split"};
        printf("test %d (%s): max|D - fp64| = %.3e  vs truncated-input model %.3e  vs rounded-input model %.3e  (max |ref| %.3f)\n",
               test, names[test], e_exact, e_trunc, e_rna, ref_max);
        if (test == 0 && e_exact > 1e-4) { printf("  FAIL: layout / descriptor mismatch\n"); ++fails; }
        if (test == 2 && e_exact > 2e-4) { printf("  FAIL: split accuracy\n"); ++fails; }
    }
    printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
    return fails;
}
