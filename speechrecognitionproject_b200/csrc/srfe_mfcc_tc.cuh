// srfe_mfcc_tc.cuh -- MFCC kernel with the DCT-II on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// One persistent 16-warp CTA per SM, warp-specialised, NO CTA-wide barrier in steady state:
//
//   frame warps (0 .. FW-1)      claim frame PAIRS from a shared-memory counter (a warp takes two, one per half-warp) and
//                                run the packed half-warp FFT -> power -> sparse Slaney sums -> log2 of srfe_kernels.cuh;
//                                the (frame A, frame B) mel rows go into a RING of pair rows in shared memory
//                                (capacity ~2 clips), followed by an atomic max (clip-wide top_db reference) and an
//                                atomic "pairs done" count per clip.  A CTA's clips form one continuous pair stream,
//                                so there are no partly filled rounds at clip boundaries any more.
//   epilogue warps (FW .. FW+3)  warp w owns TMEM quadrant w % 4 (a warp can only address its own 32 lanes) = frames
//                                32 q .. 32 q + 31 (28 q - 2 .. with two deltas).  Per clip, thread = frame:
//        stage 1  row of 128 log2-mel values from the ring (conflict-free: one LDS per value), top_db clamp against
//                 the clip maximum, re-centre on the frame mean, DCT-II symmetry fold s = x[f] + x[n-1-f],
//                 d = x[f] - x[n-1-f], split into TF32 hi + lo, tcgen05.st into the A-operand columns of TMEM;
//                 the last of the four warps to finish (shared-memory counter) issues the MMAs and releases the ring rows
//        MMA      one thread: [128 frames x n/2] . [n/2 x N] for even and odd coefficients, 3 products each
//                 (hi hi + lo hi + hi lo = fp32-grade), kind::tf32, A from TMEM, B (folded DCT-II rows, hi / lo) from
//                 shared memory (K-major, no swizzle), accumulators in TMEM; tcgen05.commit -> mbarrier
//        stage 2  (of the PREVIOUS clip, under this clip's MMAs: the accumulators exist twice) tcgen05.ld of the thread's
//                 coefficients, + c_t sqrt(n) on c0, np.gradient deltas through warp shuffles (each quadrant carries a
//                 halo of n_deltas frames, so time neighbours are always lanes of the same warp), straight to global
//                 memory: [k][t] rows are coalesced across the warp's frames.
//
// Parity: same arithmetic as the classic kernel up to the contraction (identical frame / mel code); the contraction's
// 3xTF32 error is ~2^-22 relative per product on re-centred values (|x - c_t| < 30 log2 units).
#pragma once

#include <cstdint>
#include <cstdio>

#include "srfe_kernels.cuh"

namespace srfe {

constexpr int kTcThreads = 512;               // 12 frame warps + 4 epilogue warps at 128 registers (16 + 4 at 96 measured: no faster)
constexpr int kTcEpiWarps = 4;
constexpr int kTcSlots = 4;                  // clips whose counters can be live at once (ring < 2 clips => 3 suffice)

// control block in shared memory (ints)
enum TcCtrl { TC_NEXT = 0, TC_CONSUMED = 1, TC_DONE = 2, TC_GMAX = 6, TC_S1 = 10, TC_TMEM = 14, TC_BAR = 16 /* two 8-byte mbarriers */, TC_WORDS = 20 };

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int tc_fkey(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }   // order-preserving
__device__ __forceinline__ float tc_funkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ void tc_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tc_ld8_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t tc_ld1_issue(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
// shared-memory matrix descriptor: K-major, no swizzle, core matrix = 8 rows x 16 B contiguous;
// LBO = byte step between core matrices along K, SBO = along the row (N) dimension
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool tc_elect_one() {        // one lane of the (converged) warp; ptxas then issues single-thread code without a per-instruction election loop
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred) :: "memory");
    return pred != 0;
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

template <int NFFT, int JLO, int JHI, int NG, unsigned CODE, typename SAMP, int MB = 1>
__global__ void __launch_bounds__(kTcThreads, 1) srfe_mfcc_tc_kernel(const KParams p) {
    constexpr int FAM = FAM_MFCC;
    typedef FftGeom<NFFT> G;
    extern __shared__ __align__(128) unsigned char smem[];
    const SAMP* pcm = reinterpret_cast<const SAMP*>(p.pcm);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, l = tid & 15;
    const int FW = (nthr >> 5) - kTcEpiWarps;               // frame warps

    {   // tables -> shared memory, once per (persistent) CTA: the common tables, then the DCT B operand (hi / lo, UMMA layout)
        const int4* src = reinterpret_cast<const int4*>(p.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = tid; i < p.blob_bytes / 16; i += nthr) dst[i] = __ldg(src + i);
        const int4* vs = reinterpret_cast<const int4*>(p.blob + p.tc_b_src);
        int4* vd = reinterpret_cast<int4*>(smem + p.tc_off_b);
        for (int i = tid; i < p.tc_b_bytes / 16; i += nthr) vd[i] = __ldg(vs + i);
        int4* z = reinterpret_cast<int4*>(smem + p.sm_scratch);          // see srfe_kernel: padded mel runs read a few unwritten slots
        for (int i = tid; i < (p.sm_tile - p.sm_scratch) / 16; i += nthr) z[i] = make_int4(0, 0, 0, 0);
    }
    volatile int* ctrl = reinterpret_cast<volatile int*>(smem + p.tc_off_ctrl);
    const uint32_t bar = tc_smem(smem + p.tc_off_ctrl + 4 * TC_BAR);
    if (tid == 0) {
        ctrl[TC_NEXT] = 0; ctrl[TC_CONSUMED] = 0;
        for (int s = 0; s < kTcSlots; ++s) { ctrl[TC_DONE + s] = 0; ctrl[TC_GMAX + s] = (int)0x80000000; ctrl[TC_S1 + s] = 0; }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FW) {                                       // this warp owns the TMEM allocation (and frees it at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tc_smem(smem + p.tc_off_ctrl + 4 * TC_TMEM)), "r"(p.tc_tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // B operand: generic-proxy writes -> the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = (uint32_t)ctrl[TC_TMEM];

    const float* s_win = reinterpret_cast<const float*>(smem + p.off_win);
    FftTables T;
    T.tw1 = reinterpret_cast<const cpx*>(smem + p.off_tw1);
    T.twu = reinterpret_cast<const cpx*>(smem + p.off_twu);
    T.tw16 = reinterpret_cast<const cpx*>(smem + p.off_tw16);
    const int2* g_meta = reinterpret_cast<const int2*>(smem + p.off_gm);
    const int* f_start = reinterpret_cast<const int*>(smem + p.off_fs);
    const float2* f_w2 = reinterpret_cast<const float2*>(smem + p.off_fw4);
    P2* tileP = reinterpret_cast<P2*>(smem + p.sm_tile);                 // ring of pair rows (+1 dummy row), stride p.tile_stride
    P2* fmeanP = reinterpret_cast<P2*>(smem + p.tc_off_fmean);           // per pair row: the two frame means
    const int TSP = p.tile_stride, RING = p.tc_ring;

    const int nc = (p.n_clips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // this CTA's clips: b, b + grid, ...
    const int P = (p.T + 1) >> 1;                                                        // frame pairs per clip
    const int total = nc * P;

    // The four epilogue warps (FW .. FW+3, one per TMEM quadrant) put every COMPLETE clip through the epilogue and sleep in
    // between.  Letting them claim frame pairs while idle was built and measured (B200, clock64 build): they then run about
    // one pair iteration per clip each, yet C-MFCC drops from 15.0 to 14.1 M clips/s -- twelve frame warps already hold the
    // shared-memory pipe at ~75 % of its peak, extra frame warps only add contention (and delay the epilogues they owe).
    const bool is_epi = warp >= FW;
    const int qd = warp & 3;                                // TMEM quadrant = warp % 4: the four epilogue warps are consecutive, so all four occur
    // frame of (quadrant, lane): with deltas every quadrant carries a halo of n_deltas frames on either side (TMEM rows of
    // two quadrants overlap in time), so that np.gradient's neighbours are always lanes of the same warp
    const int halo = p.n_deltas, own = 32 - 2 * halo;
    const int t = own * qd + lane - halo;                   // this thread's frame (may lie outside [0, T): clamped duplicates)
    const int tr = min(max(t, 0), p.T - 1);
    const bool towned = lane >= halo && lane < 32 - halo && t < p.T;
    const int n = p.n_filt, half = n >> 1;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * qd) << 16);
    // TMEM columns: A operand s_hi | s_lo | d_hi | d_lo (n/2 each), then the accumulators (even | odd coefficients), TWO sets
    // alternating with the clip: the MMAs of clip c run while stage 2 reads clip c-1's set
    const uint32_t col_sh = 0, col_sl = half, col_dh = 2 * half, col_dl = 3 * half, col_acc = 4 * half, acc_cols = p.tc_ne + p.tc_no;
    const int R = (1 + p.n_deltas) * p.n_mfcc;
    auto bar_of = [&](int clip) { return bar + 8u * (uint32_t)(clip & 1); };   // MMA-complete barriers, one per accumulator set
    float c_prev = 0.f;                                     // centre of this thread's frame in the clip whose stage 2 is pending

    // ---- stage 2 of clip cj (its MMAs are complete): coefficients of this thread's frame, deltas through warp shuffles,
    //      straight to global memory ----
    auto stage2 = [&](int cj, float cc) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* oc = p.out + (long long)((int)blockIdx.x + cj * (int)gridDim.x) * R * p.T;
        // element (row r = d n_mfcc + k, frame t): FT -> oc[r T + t] (coalesced across the warp's frames), TF -> oc[t R + r]
        const int st_r = p.layout == SRFE_LAYOUT_FT ? p.T : 1;                       // stride between rows
        float* ot = oc + (p.layout == SRFE_LAYOUT_FT ? t : t * R);
        // A deliberately ROLLED loop, one even and one odd coefficient per trip: the epilogue's code is cold every time
        // it runs (the frame warps stream ~100 KB of unrolled FFT code through the instruction caches in between),
        // and an unrolled version of this stage measured 5-8x slower on instruction fetch alone.
        const int st_d = p.n_mfcc * st_r;
        const uint32_t de = lane_base + col_acc + (cj & 1) * acc_cols, dox = de + p.tc_ne;
        const float c0fix = cc * p.dct_row0_sum;                                     // put the frame's centre back on c0
#pragma unroll 1
        for (int j = 0; 2 * j < p.n_mfcc; ++j) {          // (4 + 4 coefficients per trip, and issuing trip j + 1's loads under trip j,
            const uint32_t e0 = tc_ld1_issue(de + j), o0 = tc_ld1_issue(dox + j);     //  both measured no faster: the cost is code fetch)
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float ve = __uint_as_float(e0), vo = __uint_as_float(o0);
            if (j == 0) ve += c0fix;
            const bool odd_ok = 2 * j + 1 < p.n_mfcc;
            float* o = ot + 2 * j * st_r;                                           // rows k = 2 j and 2 j + 1
            if (towned) { o[0] = ve; if (odd_ok) o[st_r] = vo; }
#pragma unroll 1
            for (int d = 1; d <= p.n_deltas; ++d) {
                // np.gradient along time (unit spacing, edge_order 1): neighbours are lanes of this warp
                const float ue = __shfl_down_sync(0xffffffffu, ve, 1), le = __shfl_up_sync(0xffffffffu, ve, 1);
                const float uo = __shfl_down_sync(0xffffffffu, vo, 1), lo = __shfl_up_sync(0xffffffffu, vo, 1);
                ve = t <= 0 ? ue - ve : (t >= p.T - 1 ? ve - le : 0.5f * (ue - le));
                vo = t <= 0 ? uo - vo : (t >= p.T - 1 ? vo - lo : 0.5f * (uo - lo));
                o += st_d;
                if (towned) { o[0] = ve; if (odd_ok) o[st_r] = vo; }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    int ci = 0;                                             // epilogue warps: the next clip to finish
#ifdef SRFE_DEV
    long long tk_s1 = 0, tk_sync = 0, tk_mma = 0, tk_s2 = 0, tk_idle = 0, tk_mark = clock64(), tk_t0 = tk_mark;
#define SRFE_TICK(acc) do { const long long now__ = clock64(); acc += now__ - tk_mark; tk_mark = now__; } while (0)
#else
#define SRFE_TICK(acc) do { } while (0)
#endif
    P2* xb = reinterpret_cast<P2*>(smem + p.sm_scratch) + (tid >> 4) * G::SCRATCH_P2;
    int g_next = 0;
    if (!is_epi && p.tc_early) {
        if (lane == 0) g_next = atomicAdd(const_cast<int*>(&ctrl[TC_NEXT]), 2);
        g_next = __shfl_sync(0xffffffffu, g_next, 0);
    }
    for (;;) {
        int g0 = 0;
        if (is_epi) {
            bool quit = false;
            for (;;) {
                if (ci == nc) {                           // all clips through stage 1: the last one's stage 2 is still pending
                    tc_mbar_wait(bar_of(nc - 1), (uint32_t)(((nc - 1) >> 1) & 1));
                    stage2(nc - 1, c_prev);
                    quit = true;
                    break;
                }
                const int slot = ci & (kTcSlots - 1);
                // ONE warp watches the clip's pair count; the other three wait in a hardware barrier.  With all four polling
                // (a __nanosleep(1000) returns after ~110 cycles here: 680 polls per clip) the polling loop was 13 % of all
                // issued instructions and 5 % of the shared-memory wavefronts of the kernel.
                if (warp == FW) while (ctrl[TC_DONE + slot] < P) __nanosleep(1000);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                {
                    SRFE_TICK(tk_idle);
                    __threadfence_block();                                        // every pair of the clip is in the ring
                    // the A operand is single-buffered: the previous clip's MMAs must have read it (they have had a whole
                    // clip period; this wait is normally free) -- which also makes that clip's accumulators final
                    if (ci > 0) tc_mbar_wait(bar_of(ci - 1), (uint32_t)(((ci - 1) >> 1) & 1));
                    SRFE_TICK(tk_mma);
                    const float gmax = tc_funkey(ctrl[TC_GMAX + slot]);
                    const float thr = p.top_db >= 0.f ? gmax - p.top_db : -CUDART_INF_F;     // power_to_db(top_db), log2 units
                    // ---- stage 1: ring row -> clamp, re-centre, fold, split -> TMEM A operand ----
                    const int gp = ci * P + (tr >> 1);
                    const int row = gp - (int)__umulhi((unsigned)gp, p.tc_ring_magic) * RING;
                    const float* xrow = reinterpret_cast<const float*>(tileP + row * TSP) + (tr & 1);     // element m at xrow[2 m]
                    const float cm = reinterpret_cast<const float*>(fmeanP + row)[tr & 1];
                    const float c = fmaxf(cm, thr);                 // centre of the accumulation (classic kernel: same choice)
                    for (int f0 = 0; f0 < half; f0 += 8) {
                        float sh[8], sl[8], dh[8], dl[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int f = f0 + j;
                            const float a0 = fmaxf(xrow[2 * f], thr) - c, a1 = fmaxf(xrow[2 * (n - 1 - f)], thr) - c;
                            const float s = a0 + a1, d = a0 - a1;
                            sh[j] = __uint_as_float(__float_as_uint(s) & 0xffffe000u);       // TF32 part (kind::tf32 ignores the 13 low bits)
                            sl[j] = s - sh[j];
                            dh[j] = __uint_as_float(__float_as_uint(d) & 0xffffe000u);
                            dl[j] = d - dh[j];
                        }
                        tc_st8(lane_base + col_sh + f0, sh);
                        tc_st8(lane_base + col_sl + f0, sl);
                        tc_st8(lane_base + col_dh + f0, dh);
                        tc_st8(lane_base + col_dl + f0, dl);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    SRFE_TICK(tk_s1);
                    int last = 0;
                    if (lane == 0) {
                        __threadfence_block();
                        last = atomicAdd(const_cast<int*>(&ctrl[TC_S1 + slot]), 1) == kTcEpiWarps - 1;
                    }
                    last = __shfl_sync(0xffffffffu, last, 0);
                    if (last) {                                     // all four quadrants are in TMEM: this warp issues the contraction
                        if (lane == 0) {
                            __threadfence_block();
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t sb = tc_smem(smem + p.tc_off_b);
                            const uint32_t lbo = 128, sbo = (uint32_t)half * 32;             // (half / 4) core matrices of 128 B per 8-row group
                            const uint32_t be_bytes = (uint32_t)p.tc_ne * half * 4, bo_bytes = (uint32_t)p.tc_no * half * 4;
                            // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
                            const uint32_t id_e = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.tc_ne >> 3) << 17) | ((128u >> 4) << 24);
                            const uint32_t id_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.tc_no >> 3) << 17) | ((128u >> 4) << 24);
                            const uint32_t dset = tmem + col_acc + (ci & 1) * acc_cols;
#pragma unroll
                            for (int par = 0; par < 2; ++par) {
                                const uint32_t b_hi = sb + (par ? 2 * be_bytes : 0), b_lo = b_hi + (par ? bo_bytes : be_bytes);
                                const uint32_t a_hi = tmem + (par ? col_dh : col_sh), a_lo = tmem + (par ? col_dl : col_sl);
                                const uint32_t dcol = dset + (par ? p.tc_ne : 0), idesc = par ? id_o : id_e;
                                uint32_t acc = 0;
#pragma unroll
                                for (int pr = 0; pr < 3; ++pr) {        // hi hi + lo hi + hi lo, one accumulator
                                    const uint32_t a = pr == 1 ? a_lo : a_hi, b = pr == 2 ? b_lo : b_hi;
                                    for (int ks = 0; ks < half / 8; ++ks) {
                                        tc_mma_tf32(dcol, a + 8 * ks, tc_desc(b + ks * 2 * lbo, lbo, sbo), idesc, acc);
                                        acc = 1;
                                    }
                                }
                            }
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_of(ci)) : "memory");
                            // the clip's ring rows and counters are free again (nobody reads them after stage 1)
                            ctrl[TC_S1 + slot] = 0; ctrl[TC_DONE + slot] = 0; ctrl[TC_GMAX + slot] = (int)0x80000000;
                            __threadfence_block();
                            ctrl[TC_CONSUMED] = ci + 1;
                        }
                        __syncwarp();
                    }
                    SRFE_TICK(tk_sync);
                    // ---- stage 2 of the PREVIOUS clip, while this clip's MMAs run ----
                    if (ci > 0) stage2(ci - 1, c_prev);
                    c_prev = c;
                    SRFE_TICK(tk_s2);
                    ++ci;
                    continue;
                }
            }
            SRFE_TICK(tk_idle);
#ifdef SRFE_DEV
            if (blockIdx.x == 0 && lane == 0)
                printf("epi warp %d: clips %d cycles/clip: total %lld idle %lld stage1 %lld sync+issue %lld a-operand wait %lld stage2 %lld\n",
                       qd, nc, (clock64() - tk_t0) / nc, tk_idle / nc, tk_s1 / nc, tk_sync / nc, tk_mma / nc, tk_s2 / nc);
#endif
            (void)quit;
            break;                                          // the epilogue warps never run frame pairs (see header comment)
        } else if (p.tc_early) {
            // The pair in hand was claimed during the previous iteration; claim the one after it now, so that the atomic's
            // round trip is off the path to the next fetch and its samples can be prefetched (together +1.7 % on the headline
            // shape; the claim alone +0.4 %, a prefetch of a GUESSED next pair nothing).  Only when the ring can hold the
            // 4 FW pairs then in flight plus a clip (host: plan_tc) -- with the 51-frame clips' 52-row ring the frame warps
            // would wait for the read-out of the clip before last all the time (measured: 17.0 -> 14.2 M clips/s).
            g0 = g_next;
            if (g0 >= total) break;
            if (lane == 0) g_next = atomicAdd(const_cast<int*>(&ctrl[TC_NEXT]), 2);
            g_next = __shfl_sync(0xffffffffu, g_next, 0);
            if (g_next < total) {                           // ... and ask L2 for that pair's samples: one 128-byte line per lane
                const int gq = min(g_next + (lane >> 4), total - 1);
                const int pcn = (int)__umulhi((unsigned)gq, p.tc_p_magic), qn = gq - pcn * P;
                const SAMP* cn = pcm + (long long)((int)blockIdx.x + pcn * (int)gridDim.x) * p.clip_stride;
                const int s0 = max(p.start0 + 2 * qn * p.hop, 0) + l * (128 / (int)sizeof(SAMP));
                if (s0 < min(p.start0 + 2 * qn * p.hop + p.hop + NFFT, p.n_samples))     // two frames: hop + n_fft samples
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(cn + s0));
            }
        } else {
            if (lane == 0) g0 = atomicAdd(const_cast<int*>(&ctrl[TC_NEXT]), 2);
            g0 = __shfl_sync(0xffffffffu, g0, 0);
            if (g0 >= total) break;
        }
        const int g = g0 + (lane >> 4);
        const bool valid = g < total;                   // an odd stream ends with a half-warp that recomputes the last pair, no side effects
        const int gg = valid ? g : total - 1;
        const int pc = (int)__umulhi((unsigned)gg, p.tc_p_magic);     // the pair's clip: gg / P
        const int q = gg - pc * P;
        const SAMP* clip = pcm + (long long)((int)blockIdx.x + pc * (int)gridDim.x) * p.clip_stride;
        const int fA = 2 * q, fB = min(2 * q + 1, p.T - 1);         // odd T: the last pair's second frame repeats the last frame
        P2 pa[G::M / 32], pb[G::M / 32];
        const P2 pmid = pair_power<NFFT, FAM, JLO, JHI, SAMP>(p, clip, p.start0 + fA * p.hop, clip, p.start0 + fB * p.hop,
                                                              l, lane, s_win, T, xb, pa, pb);
        P2* pbuf = xb;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < G::M / 32; ++r) {
            const int k = NFFT == 512 ? l + 16 * r : bin640(l, r);
            pbuf[k] = pa[r];
            pbuf[G::M - k] = pb[r];
        }
        if (l == 0) pbuf[G::M / 2] = pmid;
        __syncwarp();
        // the ring row of this pair: free once the clip that last used it has been through stage 1
        const int row = valid ? gg - (int)__umulhi((unsigned)gg, p.tc_ring_magic) * RING : RING;       // RING = dummy row
        if (valid && gg >= RING) {
            const int need = (int)__umulhi((unsigned)(gg - RING), p.tc_p_magic);                     // clip of pair gg - RING
            while (ctrl[TC_CONSUMED] <= need) __nanosleep(256);
            __threadfence_block();
        }
        P2* prow = tileP + row * TSP;
        P2 fsum = bc(0.f);
        float run_max = -CUDART_INF_F;
        auto emit = [&](int m, const P2& acc, bool guard) {
            if (guard && m >= p.n_filt) return;
            // log2 units: 10 log10(2) lives in the DCT operand and in top_db (host, double precision)
            const float da = lg2_ftz(fmaxf(acc.lo, p.amin));
            const float db = lg2_ftz(fmaxf(acc.hi, p.amin));
            prow[m] = mkp(da, db);
            run_max = fmaxf(run_max, fmaxf(da, db));
            fsum = padd(fsum, mkp(da, db));
        };
        mel_project<NG, CODE, MB>(pbuf, g_meta, p.n_fgroups, f_start, f_w2, l, emit);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            fsum.lo += __shfl_xor_sync(0xffffffffu, fsum.lo, o);
            fsum.hi += __shfl_xor_sync(0xffffffffu, fsum.hi, o);
            run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
        }
        if (l == 0) fmeanP[row] = pmul(fsum, bc(1.f / (float)p.n_filt));
        __syncwarp();                                   // the half-warp's row is complete (and pbuf free for the next pair)
        if (valid && l == 0) {
            atomicMax(const_cast<int*>(&ctrl[TC_GMAX + (pc & (kTcSlots - 1))]), tc_fkey(run_max));
            __threadfence_block();                      // row, mean and max before the count
            atomicAdd(const_cast<int*>(&ctrl[TC_DONE + (pc & (kTcSlots - 1))]), 1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == FW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(p.tc_tmem_cols) : "memory");
    }
}

}  // namespace srfe
