// srfe_fbank_tc.cuh -- log filter-bank kernel with the triangular-filter projection on the 5th-generation tensor cores
// (tcgen05 / TMEM), sm_100a.  n_fft = 512, <= 128 filters.
//
// The projection  E[filter][frame] = sum_bin W[filter][bin] P[frame][bin]  is run TRANSPOSED, so that the frames are the
// N dimension of the MMA and a tile of N = 48 frames is a complete operand as soon as its 24 frame pairs are done:
//
//   A operand  = the filter weights, M = 128 filters (rows past n_filt are zero) x K, resident in TMEM for the whole
//                kernel as bf16 hi + bf16 mid (round-to-nearest split of the fp32 weights the classic kernel uses)
//   B operand  = the power spectra of 48 frames in shared memory, K-major, no swizzle: [frame][bins] is exactly what the
//                half-warp FFT produces.  The K order is free (the host permutes W the same way), so every lane's 16 bins
//                of a frame are two contiguous 16-byte chunks: chunk l = bins l + 16 r, chunk 16 + l = bins 256 - (l + 16 r);
//                bin 128 rides in the slot of an edge bin that no filter weighs (srfe_abi.cu), so K = 256 = 16 MMA steps.
//                Each value is split into bf16 hi + mid (cvt.rn: zero-mean remainders); chunks are 144 B apart along K
//                (LBO), which puts the 16 lanes' 16-byte stores on distinct banks.
//   products   = W_hi P_hi + W_hi P_mid + W_mid P_hi, kind::f16 (bf16 x bf16 -> f32): 48 MMAs of K = 16 per tile.  All
//                terms are non-negative, so the relative error of a band sum is that of its terms: <= 2.4e-5 measured
//                (scripts/ubench/tc_probe_bf16.cu), i.e. 2e-4 on the 20 log10 scale; exact zeros stay exact zeros.
//   D          = [128 filters (TMEM lanes) x 48 frames (columns)] f32, two accumulators alternating with the tile.
//
// One persistent 16-warp CTA per SM, warp-specialised, no CTA-wide barrier in steady state:
//   frame warps (12)    claim frame PAIRS from a shared-memory counter (the CTA's clips are one continuous pair stream;
//                       tiles ignore clip boundaries), run the packed half-warp FFT of srfe_kernels.cuh (pair_power),
//                       split the power values and store them into the tile buffer (three buffers of 48 frames), then
//                       fence.proxy.async + an atomic "pairs done" count per tile.  The warp whose count completes a tile
//                       issues the tile's MMAs (one elected thread, a rolled loop) + tcgen05.commit onto the mbarrier of
//                       the tile's accumulator.  No power buffer, no sparse sums, no logs on these warps any more.
//   read-out warps (4)  wait for the tile's mbarrier (one warp polls, the others sit in a hardware barrier), read their
//                       TMEM quadrant (thread = filter, columns = frames), 0 -> eps, 20 log10 via lg2.approx, and store
//                       out[clip][frame][filter]: the 32 lanes of a warp write 128 contiguous bytes per frame.
//
// Measured on the way (B200, 16,384 clips, clock64 build; profiles/r2_notes.md): a single warp runs ~8 cycles per
// instruction here whatever it does, so serial work per tile is what counts -- with one read-out warp also issuing the
// MMAs (2.5 k cycles per tile, ~53 per MMA whatever N) and an unrolled read-out (4.5 k) the read-out warps, not the frame
// warps, set the pace (16.3 M clips/s against 23.2 M for the frame warps alone and 18.4 M for the classic kernel).
#pragma once

#include "srfe_mfcc_tc.cuh"

namespace srfe {

constexpr int kFtThreads = 512;                           // 12 frame warps + 4 epilogue warps (128 registers)
constexpr int kFtTile = 48;                               // frames per MMA tile (the MMA's N, a multiple of 16)
constexpr int kFtBufs = 3;                                // tile buffers: 24 pairs are in flight on the frame warps and a tile's MMAs
                                                          // take ~2.5 k cycles, so two buffers of 64 frames stalled the frame warps
constexpr int kFtPairs = kFtTile / 2;
constexpr int kFtKP = 256, kFtCH = kFtKP / 8;             // K per part: 256 of the 257 bins (see below), 16-byte chunks per part
constexpr int kFtLBO = 144;                               // byte step between K chunks (core matrices are 128 B: 16 B of padding)
constexpr int kFtSBO = 2 * kFtCH * kFtLBO;                // byte step between 8-frame groups: hi chunks, then mid chunks
constexpr int kFtTileBytes = (kFtTile / 8) * kFtSBO;      // 55,296
constexpr int kFtTmemCols = 512;                          // 256 (weights) + 2 x 48 (accumulators) -> next power of two
constexpr int kFtSlots = 4;

enum FtCtrl { FT_NEXT = 0, FT_CONSUMED = 1, FT_FULL = 2 /* kFtSlots */, FT_TMEM = 6, FT_READ = 7, FT_BAR = 8 /* two 8-byte mbarriers */, FT_WORDS = 12 };

__device__ __forceinline__ uint32_t ft_pack_bf16(float lo_elem, float hi_elem) {      // {hi_elem : lo_elem}, round to nearest even
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
    return d;
}
__device__ __forceinline__ void ft_mma_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// one 16-byte chunk (8 K elements = this lane's slots r = 0 .. 7 of one frame) as bf16 hi and bf16 mid
template <bool HI_HALF>
__device__ __forceinline__ void ft_split8(const P2* v, uint4& h, uint4& m) {
    uint32_t hw[4], mw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x0 = HI_HALF ? v[2 * i].hi : v[2 * i].lo, x1 = HI_HALF ? v[2 * i + 1].hi : v[2 * i + 1].lo;
        hw[i] = ft_pack_bf16(x0, x1);
        const float r0 = x0 - __uint_as_float(hw[i] << 16), r1 = x1 - __uint_as_float(hw[i] & 0xffff0000u);   // exact
        mw[i] = ft_pack_bf16(r0, r1);
    }
    h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    m = make_uint4(mw[0], mw[1], mw[2], mw[3]);
}

template <int JLO, int JHI, typename SAMP>
__global__ void __launch_bounds__(kFtThreads, 1) srfe_fbank_tc_kernel(const KParams p) {
    constexpr int NFFT = 512, FAM = FAM_FBANK;
    typedef FftGeom<NFFT> G;
    extern __shared__ __align__(128) unsigned char smem[];
    const SAMP* pcm = reinterpret_cast<const SAMP*>(p.pcm);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, l = tid & 15;
    const int FW = (nthr >> 5) - kTcEpiWarps;               // frame warps

    {   // FFT tables -> shared memory (the sparse mel tables of the classic kernel are not needed); tile buffers zeroed once
        // (rows past the end of the last, partly filled tile are never written: their columns of D are never read either)
        const int4* src = reinterpret_cast<const int4*>(p.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = tid; i < p.blob_bytes / 16; i += nthr) dst[i] = __ldg(src + i);
        int4* z = reinterpret_cast<int4*>(smem + p.tc_off_b);
        for (int i = tid; i < kFtBufs * kFtTileBytes / 16; i += nthr) z[i] = make_int4(0, 0, 0, 0);
    }
    volatile int* ctrl = reinterpret_cast<volatile int*>(smem + p.tc_off_ctrl);
    const uint32_t bar = tc_smem(smem + p.tc_off_ctrl + 4 * FT_BAR);
    if (tid == 0) {
        ctrl[FT_NEXT] = 0; ctrl[FT_CONSUMED] = 0; ctrl[FT_READ] = 0;
        for (int s = 0; s < kFtSlots; ++s) ctrl[FT_FULL + s] = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FW) {                                       // this warp owns the TMEM allocation (and frees it at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(tc_smem(smem + p.tc_off_ctrl + 4 * FT_TMEM)), "r"(kFtTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // the zeroed tiles: generic-proxy writes -> the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = (uint32_t)ctrl[FT_TMEM];
    const uint32_t col_acc = kFtKP;                         // weights: columns [0, 128) hi pairs, [128, 256) mid pairs; then 2 accumulators

    const int nc = (p.n_clips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // this CTA's clips: b, b + grid, ...
    const int P = (p.T + 1) >> 1;                                                        // frame pairs per clip
    const int total = nc * P;

    if (warp >= FW) {
        // ------------------------------------------------------------------ read-out warps ----
        const int qd = warp & 3;                            // TMEM quadrant = warp % 4 (four consecutive warps: all four occur)
        const int m = 32 * qd + lane;                       // this thread's filter
        const uint32_t lane_base = tmem + ((uint32_t)(32 * qd) << 16);
        {   // the thread's weight row -> TMEM (global table: [128][256] packed bf16 pairs, K-permuted by the host)
            const uint4* wsrc = reinterpret_cast<const uint4*>(p.blob + p.tc_b_src) + (size_t)m * (kFtKP / 4);
#pragma unroll 1
            for (int c = 0; c < kFtKP; c += 8) {
                const uint4 a = __ldg(wsrc + c / 4), b = __ldg(wsrc + c / 4 + 1);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             :: "r"(lane_base + c), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == FW && lane == 0) { __threadfence_block(); ctrl[FT_READ] = 1; }      // weights in place: MMAs may be issued
        }
        const int ntiles = (total + kFtPairs - 1) / kFtPairs;
        const bool m_ok = m < p.n_filt;
        const bool q_ok = 32 * qd < p.n_filt;               // quadrants without a valid filter skip the read-out
        const int TP = 2 * P;                               // stream slots per clip (odd T: the last slot repeats frame T - 1)
#pragma unroll 1
        for (int ti = 0; ti < ntiles; ++ti) {
            if (warp == FW) {
                // only this warp watches the mbarrier (with sleeps); the other three wait in the hardware barrier below --
                // warps spinning on mbarrier.try_wait take issue slots from the frame warps
                const uint32_t parity = (uint32_t)((ti >> 1) & 1), b = bar + 8u * (uint32_t)(ti & 1);
                for (;;) {
                    uint32_t done;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(b), "r"(parity) : "memory");
                    if (done) break;
                    __nanosleep(200);
                }
                if (lane == 0) {                            // the tile buffer has been read: frame warps may refill it
                    __threadfence_block();
                    ctrl[FT_CONSUMED] = ti + 1;
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");    // the tile's accumulator is complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (q_ok) {
                // column j of tile ti = stream slot s = kFtTile ti + j = (clip s / TP, slot u = s % TP); frame u when u < T.
                // A ROLLED loop, eight columns per trip, the next trip's TMEM load issued under this trip's arithmetic: this
                // code is cold every time it runs (the frame warps stream ~100 KB of unrolled FFT code through the
                // instruction caches in between) and a warp gets an instruction issued every ~8 cycles here
                const int s_end = min(kFtTile * (ti + 1), 2 * total);
                const uint32_t acc0 = lane_base + col_acc + (uint32_t)(ti & 1) * kFtTile;
                int s = kFtTile * ti;
                int pc = (int)__umulhi((unsigned)(s >> 1), p.tc_p_magic);
                int u = s - pc * TP;
                float* orow = p.out + ((long long)((int)blockIdx.x + pc * (int)gridDim.x) * p.T + u) * p.n_filt + m;
                uint32_t r[8], rn[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                tc_ld8_issue(acc0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll 1
                for (int c0 = 0; c0 < kFtTile && s < s_end; c0 += 8) {
                    if (c0 + 8 < kFtTile) tc_ld8_issue(acc0 + c0 + 8, rn);
                    if (u >= TP) {                          // the previous trip ended on a clip boundary
                        u = 0; ++pc;
                        orow = p.out + (long long)((int)blockIdx.x + pc * (int)gridDim.x) * p.T * p.n_filt + m;
                    }
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float a = __uint_as_float(r[j]);
                        // model_fbanks_cnn.py:61: exact zeros -> eps (same treatment as the classic kernel's emit)
                        a = (a == 0.f) ? 2.220446049250313e-16f : fmaxf(a, 1.17549435e-38f);
                        y[j] = 6.020599913279624f * lg2_ftz(a);                          // 20 log10
                    }
                    if (u + 8 <= p.T && s + 8 <= s_end) {
                        if (m_ok) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) orow[j * p.n_filt] = y[j];
                        }
                        s += 8; u += 8;
                        orow += 8 * p.n_filt;
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 8; ++j) {
                            float yj = y[0];
#pragma unroll
                            for (int jj = 1; jj < 8; ++jj) yj = (j == jj) ? y[jj] : yj;
                            if (m_ok && u < p.T && s < s_end) *orow = yj;
                            ++s; ++u;
                            orow += p.n_filt;
                            if (u >= TP) {
                                u = 0; ++pc;
                                orow = p.out + (long long)((int)blockIdx.x + pc * (int)gridDim.x) * p.T * p.n_filt + m;
                            }
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = rn[j];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");    // the accumulator is drained: tile ti + 2 may overwrite it
            if (warp == FW && lane == 0) { __threadfence_block(); ctrl[FT_READ] = ti + 2; }
        }
    } else {
        // ------------------------------------------------------------------ frame warps ----
        const float* s_win = reinterpret_cast<const float*>(smem + p.off_win);
        FftTables T;
        T.tw1 = reinterpret_cast<const cpx*>(smem + p.off_tw1);
        T.twu = reinterpret_cast<const cpx*>(smem + p.off_twu);
        T.tw16 = reinterpret_cast<const cpx*>(smem + p.off_tw16);
        P2* xb = reinterpret_cast<P2*>(smem + p.sm_scratch) + (tid >> 4) * G::SCRATCH_P2;
        unsigned char* tiles = smem + p.tc_off_b;
        for (;;) {
            int g0 = 0;
            if (lane == 0) g0 = atomicAdd(const_cast<int*>(&ctrl[FT_NEXT]), 2);
            g0 = __shfl_sync(0xffffffffu, g0, 0);
            if (g0 >= total) break;
            const int g = g0 + (lane >> 4);
            const bool valid = g < total;                   // an odd stream ends with a half-warp that recomputes the last pair, no side effects
            const int gg = valid ? g : total - 1;
            const int pc = (int)__umulhi((unsigned)gg, p.tc_p_magic);     // the pair's clip: gg / P
            const int q = gg - pc * P;
            const SAMP* clip = pcm + (long long)((int)blockIdx.x + pc * (int)gridDim.x) * p.clip_stride;
            const int fA = 2 * q, fB = min(2 * q + 1, p.T - 1);           // odd T: the last pair's second frame repeats the last frame
            P2 pa[G::M / 32], pb[G::M / 32];
            P2 pmid = pair_power<NFFT, FAM, JLO, JHI, SAMP>(p, clip, p.start0 + fA * p.hop, clip, p.start0 + fB * p.hop,
                                                            l, lane, s_win, T, xb, pa, pb);
            // the pair's two rows of its tile buffer (g0 and the pairs per tile are even: both half-warps, same tile)
            const int ti = (int)((unsigned)g0 / (unsigned)kFtPairs), slot = gg - ti * kFtPairs;
            if (ti >= kFtBufs) {                            // the buffer's previous tile must have been through its MMAs
                while (ctrl[FT_CONSUMED] < ti - kFtBufs + 1) __nanosleep(128);
                __threadfence_block();
            }
            if (valid) {
                unsigned char* rowA = tiles + (ti % kFtBufs) * kFtTileBytes + (slot >> 2) * kFtSBO + ((2 * slot) & 7) * 16 + l * kFtLBO;
                // bin 128 (held by lane 0) takes the K slot of the weightless edge bin: bin 0 = lane 0's pa[0], bin 256 = its pb[0]
                if (l == 0) { if (p.tc_ne == 0) pa[0] = pmid; else pb[0] = pmid; }
                uint4 h, md;
                ft_split8<false>(pa, h, md);                                            // frame A, bins l + 16 r
                *reinterpret_cast<uint4*>(rowA) = h;
                *reinterpret_cast<uint4*>(rowA + kFtCH * kFtLBO) = md;
                ft_split8<true>(pa, h, md);                                             // frame B
                *reinterpret_cast<uint4*>(rowA + 16) = h;
                *reinterpret_cast<uint4*>(rowA + 16 + kFtCH * kFtLBO) = md;
                ft_split8<false>(pb, h, md);                                            // frame A, bins 256 - (l + 16 r)
                *reinterpret_cast<uint4*>(rowA + 16 * kFtLBO) = h;
                *reinterpret_cast<uint4*>(rowA + (16 + kFtCH) * kFtLBO) = md;
                ft_split8<true>(pb, h, md);
                *reinterpret_cast<uint4*>(rowA + 16 + 16 * kFtLBO) = h;
                *reinterpret_cast<uint4*>(rowA + 16 + (16 + kFtCH) * kFtLBO) = md;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // these rows are read by the MMA (async proxy)
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                const int cnt = g0 + 1 < total ? 2 : 1, expect = min(kFtPairs, total - kFtPairs * ti);
                __threadfence_block();
                last = atomicAdd(const_cast<int*>(&ctrl[FT_FULL + (ti & (kFtSlots - 1))]), cnt) + cnt == expect;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                // this warp completed tile ti: it issues the tile's MMAs.  The accumulator (ti & 1) must have been read out
                // (tile ti - 2; normally long done) -- FT_READ also covers the weights' arrival in TMEM before tile 0
                if (lane == 0) ctrl[FT_FULL + (ti & (kFtSlots - 1))] = 0;
                while (ctrl[FT_READ] < max(ti, 1)) __nanosleep(64);
                __threadfence_block();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tc_elect_one()) {
                    // instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
                    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kFtTile >> 3) << 17) | ((128u >> 4) << 24);
                    const uint32_t tb = tc_smem(tiles) + (uint32_t)(ti % kFtBufs) * kFtTileBytes;
                    const uint32_t dcol = tmem + col_acc + (uint32_t)(ti & 1) * kFtTile;
                    uint32_t acc = 0;
#pragma unroll 1
                    for (int pr = 0; pr < 3; ++pr) {        // W_hi P_hi + W_hi P_mid + W_mid P_hi (rolled: small code in the frame loop)
                        uint32_t a_col = tmem + (pr == 2 ? kFtKP / 2 : 0);
                        uint64_t bd = tc_desc(tb + (pr == 1 ? kFtCH * kFtLBO : 0), kFtLBO, kFtSBO);
#pragma unroll 1
                        for (int ks = 0; ks < kFtKP / 16; ++ks) {
                            ft_mma_bf16(dcol, a_col, bd, idesc, acc);
                            a_col += 8;
                            bd += (2 * kFtLBO) >> 4;        // two K chunks per MMA (start-address field, 16-byte units)
                            acc = 1;
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 :: "r"(bar + 8u * (uint32_t)(ti & 1)) : "memory");
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == FW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kFtTmemCols) : "memory");
    }
}

}  // namespace srfe
