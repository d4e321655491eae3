// srfe_kernels.cuh -- fused feature kernels for sm_100a (packed two-frames-per-lane version).
//
// Persistent CTAs: grid = min(#clip groups, SMs x CTAs/SM); a CTA loops over groups of
// `cpc` consecutive clips (1 for MFCC, 1..8 for SPEC/FBANK).  Inside a group the
// flattened frames are taken two at a time: half-warp h of the CTA's HW = blockDim/16
// half-warps processes the frame PAIRS q = h, h+HW, ... (frames 2q and 2q+1) with every
// value held as an f32x2 register pair (srfe_fft.cuh).  HW and cpc are picked on the host
// so that cpc*T fills whole rounds of 2*HW frames.
//
//   global PCM --LDG.64, coalesced (16 lanes x 8 B = one 128 B line per load; the
//                2-2.5x frame overlap is served by L1/L2, HBM sees each sample once)
//     -> pre-emphasis (fbank, exact fp32 rounding) -> window (shared table)
//     -> packed half-warp FFT (registers + one/two shared-memory exchanges, real and imaginary
//        planes one after the other through the same 8-byte slots); the window extent
//        [32 JLO, 32 JHI) is a template parameter, so the zero inputs of the 400-in-512
//        frames are constant-folded out of the first pass
//     -> untangle (mirror bins by warp shuffle) -> power (both frames at once)
//     -> SPEC : scale, ln(. + eps)              -> global (TF) or a CTA-wide [bin][frame] tile (FT)
//        FBANK: sparse triangle sums (uniform-trip ELL), 20 log10 -> global [T][nfilt]
//        MFCC : sparse Slaney sums, 10 log10    -> per-clip dB tile in shared memory
//   MFCC epilogue (CTA barrier): clip max -> top_db clamp, re-centre, fold -> DCT-II in packed
//   FFMA2, one thread per (coefficient block, frame pairs) item -> np.gradient deltas ->
//   coalesced store.  Only final features reach HBM.  Two MFCC CTAs share an SM so that one's
//   epilogue overlaps the other's frame phase.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/srfe.h"
#include "srfe_fft.cuh"

// Developer switches exist only in -DSRFE_DEV builds; the shipped kernels carry no way to skip a phase.
#ifdef SRFE_DEV
#define SRFE_DBG(p, bit) ((p).debug & (bit))
#else
#define SRFE_DBG(p, bit) 0
#endif

namespace srfe {

enum Family { FAM_SPEC = 0, FAM_FBANK = 1, FAM_MFCC = 2 };

constexpr int kMaxThreads = 512;       // <= 16 warps per CTA at 128 registers (640 threads / 96 regs measured slower)

struct KParams {
    const void* pcm;                   // float32 or int16 samples (kernel template parameter SAMP)
    float* out;
    long long clip_stride;
    int n_clips, n_samples;
    int T, hop, start0;
    int cpc, n_groups;                 // clips per group, number of groups
    unsigned t_magic;                  // ceil(2^32 / T): f / T == umulhi(f, t_magic)
    int debug;                         // SRFE_DEV builds only (phase-skipping switches for timing experiments); ignored otherwise
    int sm_ctile;                      // coefficient tile offset (aliases the FFT scratch)
    const unsigned char* blob;         // tables, copied to shared memory once per CTA
    int blob_bytes;                    // multiple of 16
    int off_win, off_tw1, off_twu, off_tw16;
    int off_gm, off_fs, off_fw4;       // mel ELL: int2 {w4 offset, n4} per 16-filter group; start bin per filter; planar float2 weights
    int n_filt, n_fgroups;
    float scale, log_eps;
    int take_log, layout;
    float preemph;
    int n_mfcc, n_deltas;
    float top_db, amin, dct_row0_sum;  // top_db in log2 units (dB / 3.0103); the DCT tables carry the 3.0103
    const float* dct_kf;               // global [n_mfcc][n_mels] DCT-II rows (generic path, odd n_mels)
    int dct_fold, off_dfold;           // folded DCT table: shared-memory offset (right after the common tables)
    int dct_cb, dct_pq, dct_nbe, dct_nbo;   //   coefficients per block, frame pairs per thread, even / odd block counts
    int dct_src, dct_bytes;            //   where the chosen block-size variant sits in the global blob
    int sm_scratch, sm_tile;           // byte offsets into dynamic shared memory
    int tile_stride;                   // MFCC dB tile row stride (P2 units, odd)
    int w_lo, w_hi;                    // non-zero extent of the window (host: picks the JLO / JHI instantiation)
    int sm_stage, sm_stage_mb;         // staged spectrogram: per-half-warp 2-frame buffers and mbarriers (shared-memory offsets)
    unsigned stage_phase0;             //   (always 0: the barriers are initialised by the kernel)
    // tcgen05 MFCC kernel (srfe_mfcc_tc.cuh)
    int tc_b_src, tc_b_bytes, tc_off_b;        // DCT B operand (hi / lo, UMMA layout): offset in the global blob, bytes, shared-memory offset
    int tc_ne, tc_no, tc_tmem_cols;            // accumulator widths (even / odd coefficients, multiples of 16), TMEM columns to allocate
    int tc_ring, tc_off_fmean, tc_off_ctrl;    // ring capacity in pair rows, per-row frame means, control block
    unsigned tc_ring_magic, tc_p_magic;        // ceil(2^32 / RING), ceil(2^32 / pairs per clip)
    int tc_early;                              // frame warps claim their next pair one iteration ahead (ring permitting)
};

// --------------------------------------------------------------------------------
// frame fetch: raw samples for n = 2 l + 32 j (+1), j in [JLO, JHI)
// --------------------------------------------------------------------------------
// sample loads: float32 as the dataset hands it over (dataset.py:117) or the wav's native int16 (dataset.py:103)
template <typename S> __device__ __forceinline__ float ld1(const S* q);
template <> __device__ __forceinline__ float ld1<float>(const float* q) { return __ldg(q); }
template <> __device__ __forceinline__ float ld1<short>(const short* q) { return (float)__ldg(q); }
template <typename S> __device__ __forceinline__ float2 ld2(const S* q);
template <> __device__ __forceinline__ float2 ld2<float>(const float* q) { return __ldg(reinterpret_cast<const float2*>(q)); }
template <> __device__ __forceinline__ float2 ld2<short>(const short* q) {
    const short2 v = __ldg(reinterpret_cast<const short2*>(q));
    return make_float2((float)v.x, (float)v.y);
}

// staged frames (shared memory, filled by cp.async.bulk): plain loads
template <typename S> __device__ __forceinline__ float2 lds2(const S* q);
template <> __device__ __forceinline__ float2 lds2<float>(const float* q) { return *reinterpret_cast<const float2*>(q); }
template <> __device__ __forceinline__ float2 lds2<short>(const short* q) {
    const short2 v = *reinterpret_cast<const short2*>(q);
    return make_float2((float)v.x, (float)v.y);
}

template <int FAM, typename S>
__device__ __forceinline__ float edge_sample(const KParams& p, const S* __restrict__ x, int idx) {
    if (FAM == FAM_MFCC) {                                  // np.pad(mode='reflect')
        if (idx < 0) idx = -idx;
        if (idx >= p.n_samples) idx = 2 * (p.n_samples - 1) - idx;
        return ld1<S>(x + idx);
    }
    if (FAM == FAM_FBANK) {                                 // zero padding past the clip; e[0] = x[0]
        if (idx >= p.n_samples) return 0.f;
        const float prev = idx > 0 ? ld1<S>(x + idx - 1) : 0.f;
        return __fsub_rn(ld1<S>(x + idx), __fmul_rn(p.preemph, prev));   // model_fbanks_cnn.py:20 (float32)
    }
    return (idx >= 0 && idx < p.n_samples) ? ld1<S>(x + idx) : 0.f;
}

template <int FAM, int NJ>
struct RawFrame {
    float2 s[NJ];
    float prev[FAM == FAM_FBANK ? NJ : 1];
    bool final_;                       // samples already carry the pre-emphasis (edge path)
};

template <int FAM, int JLO, int JHI, typename S, bool SMEM = false>
__device__ __forceinline__ void fetch_frame(const KParams& p, const S* __restrict__ x, int base, int l,
                                            RawFrame<FAM, JHI - JLO>& r) {
    if (SMEM) {                                             // x = the frame's staged copy, base = 0, always interior
        r.final_ = false;
#pragma unroll
        for (int j = JLO; j < JHI; ++j) r.s[j - JLO] = lds2<S>(x + 2 * l + 32 * j);
        return;
    }
    const bool interior = (base + 32 * JLO >= 0) && (base + 32 * JHI <= p.n_samples);
    r.final_ = !interior;
    if (interior) {
        const S* xs = x + base + 2 * l;
#pragma unroll
        for (int j = JLO; j < JHI; ++j) {
            r.s[j - JLO] = ld2<S>(xs + 32 * j);
            if (FAM == FAM_FBANK) r.prev[j - JLO] = (base + 2 * l + 32 * j > 0) ? ld1<S>(xs + 32 * j - 1) : 0.f;
        }
    } else {
#pragma unroll
        for (int j = JLO; j < JHI; ++j) {
            const int i = base + 2 * l + 32 * j;
            r.s[j - JLO] = make_float2(edge_sample<FAM, S>(p, x, i), edge_sample<FAM, S>(p, x, i + 1));
            if (FAM == FAM_FBANK) r.prev[j - JLO] = 0.f;
        }
    }
}

// Frame B of a pair when it starts HS * 32 samples after frame A in the same clip (hop = 160 = 5 x 32 for the filter-bank
// presets): lane l's sample pair (l, j) of B is the pair (l, j + HS) of A, so only B's last HS pairs (and, for FBANK, the
// samples before them) are loaded: 36 instead of 52 loads per lane and frame pair.
template <int FAM, int JLO, int JHI, int HS, typename S>
__device__ __forceinline__ void fetch_frame_after(const KParams& p, const S* __restrict__ x, int base, int l,
                                                  const RawFrame<FAM, JHI - JLO>& ra, RawFrame<FAM, JHI - JLO>& rb) {
    constexpr int NJ = JHI - JLO;
    rb.final_ = false;
    const S* xs = x + base + 2 * l;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        if (i + HS < NJ) {
            rb.s[i] = ra.s[i + HS];
            if (FAM == FAM_FBANK) rb.prev[i] = ra.prev[i + HS];
        } else {
            rb.s[i] = ld2<S>(xs + 32 * (i + JLO));
            if (FAM == FAM_FBANK) rb.prev[i] = ld1<S>(xs + 32 * (i + JLO) - 1);       // base > 0 here: never the clip's first sample
        }
    }
}

template <int FAM, int NJ>
__device__ __forceinline__ void emphasise(const KParams& p, const RawFrame<FAM, NJ>& r, int j, float& s0, float& s1) {
    s0 = r.s[j].x;
    s1 = r.s[j].y;
    if (FAM == FAM_FBANK && !r.final_) {                    // float32, no FMA contraction
        s1 = __fsub_rn(s1, __fmul_rn(p.preemph, s0));
        s0 = __fsub_rn(s0, __fmul_rn(p.preemph, r.prev[j]));
    }
}

// v[j] = window[n] * (sA[n], sB[n]); zero outside [JLO, JHI) (compile-time, folds into the DFT)
template <int NFFT, int FAM, int JLO, int JHI>
__device__ __forceinline__ void window_pair(const KParams& p, const RawFrame<FAM, JHI - JLO>& ra,
                                            const RawFrame<FAM, JHI - JLO>& rb, int l, const float* s_win, C2* v) {
    typedef FftGeom<NFFT> G;
#pragma unroll
    for (int j = 0; j < G::V; ++j) {
        if (j >= JLO && j < JHI) {
            const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * l + 32 * j);
            float a0, a1, b0, b1;
            emphasise<FAM>(p, ra, j - JLO, a0, a1);
            emphasise<FAM>(p, rb, j - JLO, b0, b1);
            v[j].re = mkp(a0 * w.x, b0 * w.x);           // scalar FMULs land directly in the pair registers
            v[j].im = mkp(a1 * w.y, b1 * w.y);           // (a packed multiply would need 2 MOVs to form its operand)
        } else {
            v[j].re = bc(0.f);
            v[j].im = bc(0.f);
        }
    }
}

// lg2.approx with flush-to-zero: no denormal guard code (arguments are bounded below by eps / amin)
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- N = 512 untangle without the shared-memory round trip ----------------------------------------------
// After pass 2 lane l holds v[k2] = Z[l + 16 k2].  The mirror of k = l + 16 r (r < 8) is M - k = (16 - l) + 16 (15 - r):
// register 15 - r of lane (16 - l) & 15 of the same half-warp -- 4 shuffles per packed value, 32 per lane, instead of a
// 16 x STS.128 store of Z plus 33 x LDS.128 reads (24.5 KB of shared-memory traffic per 4-frame pass).  Lane 0 mirrors
// onto itself with a different register index (Z[256 - 16 r] = v[16 - r]; r = 0: Z[0]); lane 8 mirrors onto itself exactly.
__device__ __forceinline__ P2 fft_untangle_512_shfl(int l, int lane, const C2* v, const FftTables& T, P2* pa, P2* pb) {
    const int src = (lane & 16) | ((16 - l) & 15);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        C2 zm;
        zm.re.lo = __shfl_sync(0xffffffffu, v[15 - r].re.lo, src);
        zm.re.hi = __shfl_sync(0xffffffffu, v[15 - r].re.hi, src);
        zm.im.lo = __shfl_sync(0xffffffffu, v[15 - r].im.lo, src);
        zm.im.hi = __shfl_sync(0xffffffffu, v[15 - r].im.hi, src);
        if (l == 0) zm = (r == 0) ? v[0] : v[(16 - r) & 15];
        const cpx w = T.twu[l + 16 * r];
        untangle_pair(v[r], zm, w.x, w.y, pa[r], pb[r]);
    }
    return pmul(bc(4.f), pfma(v[8].re, v[8].re, pmul(v[8].im, v[8].im)));      // bin M/2 (meaningful on lane 0)
}

// ---- N = 640 untangle through shuffles --------------------------------------------------------------------
// After pass 3 lane (c = l & 3, b = l >> 2) holds v[4 i + kb] = Z[k1 + 20 k2], k1 = b + 4 i, k2 = c + 4 kb.  Each lane
// takes its 10 bins with kb < 2 (k < 160) and needs the mirror M - k = (20 - k1) % 20 + 20 (15 - k2) [k1 != 0] or
// 20 ((16 - k2) % 16) [k1 == 0]:
//   b != 0          : lane (3 - c, 4 - b), register 19 - j                      (j = 4 i + kb)
//   b == 0, i >= 1  : lane (3 - c, 0),     register 23 - j
//   b == 0, i == 0  : lane ((4 - c) & 3, 0), register (c == 0 ? (4 - kb) & 3 : 3 - kb)
// A b == 0 lane is only ever read by b == 0 lanes, so the sender picks the register by its own (b, c).
// bin of slot r = 2 i + kb:  k = b + 4 i + 20 c + 80 kb.
__device__ __forceinline__ int bin640(int l, int r) { return (l >> 2) + 4 * (r >> 1) + 20 * (l & 3) + 80 * (r & 1); }

__device__ __forceinline__ P2 fft_untangle_640_shfl(int l, int lane, const C2* v, const FftTables& T, P2* pa, P2* pb) {
    const int c = l & 3, b = l >> 2, hb = lane & 16;
    const int src_hi = hb | (b != 0 ? (3 - c) + 4 * (4 - b) : (3 - c));      // slots with i >= 1
    const int src_lo = hb | (b != 0 ? (3 - c) + 4 * (4 - b) : ((4 - c) & 3));    // slots with i == 0
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const int j = 4 * i + kb, r = 2 * i + kb;
            C2 snd;
            if (i >= 1) snd = (b != 0) ? v[19 - j] : v[23 - j];
            else        snd = (b != 0) ? v[19 - j] : ((c == 0) ? v[(4 - kb) & 3] : v[3 - kb]);
            const int src = (i >= 1) ? src_hi : src_lo;
            C2 zm;
            zm.re.lo = __shfl_sync(0xffffffffu, snd.re.lo, src);
            zm.re.hi = __shfl_sync(0xffffffffu, snd.re.hi, src);
            zm.im.lo = __shfl_sync(0xffffffffu, snd.im.lo, src);
            zm.im.hi = __shfl_sync(0xffffffffu, snd.im.hi, src);
            const cpx w = T.twu[bin640(l, r)];
            untangle_pair(v[j], zm, w.x, w.y, pa[r], pb[r]);
        }
    }
    return pmul(bc(4.f), pfma(v[2].re, v[2].re, pmul(v[2].im, v[2].im)));      // bin 160 = lane 0, register (i = 0, kb = 2)
}

struct FramePos { int c, t; bool ok; };
// flattened frame -> (clip in group, frame); division by T through a host-computed magic multiplier
// (exact for f * T < 2^32); single-clip groups (MFCC) skip it altogether
__device__ __forceinline__ FramePos frame_pos(int f, int nf, int T, unsigned magic, int cpc) {
    FramePos r;
    r.ok = f < nf;
    const int fc = r.ok ? f : (nf - 1);                     // invalid frames alias the last valid one (outputs suppressed)
    r.c = (cpc == 1) ? 0 : (T == 1 ? fc : (int)__umulhi((unsigned)fc, magic));   // T == 1: the magic would be 2^32
    r.t = fc - r.c * T;
    return r;
}

// --------------------------------------------------------------------------------
// one frame PAIR of a half-warp: fetch -> pre-emphasis -> window -> packed FFT -> untangle.
//   pa[r] / pb[r] = 4|X[k]|^2 / 4|X[M-k]|^2 of both frames for this lane's bins (k = l + 16 r for N = 512, bin640(l, r)
//   for N = 640); returns 4|X[M/2]|^2 (meaningful on lane 0).  xb = the half-warp's exchange scratch.
// --------------------------------------------------------------------------------
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// SMEM: the two frames were staged in shared memory (clipA / clipB point at them, bases 0); `after_fetch` runs once the
// samples sit in registers (the staging buffer may be refilled from there on)
template <int NFFT, int FAM, int JLO, int JHI, typename SAMP, bool SMEM = false, typename Hook = NoHook>
__device__ __forceinline__ P2 pair_power(const KParams& p, const SAMP* __restrict__ clipA, int baseA, const SAMP* __restrict__ clipB,
                                         int baseB, int l, int lane, const float* s_win, const FftTables& T, P2* xb, P2* pa, P2* pb,
                                         Hook after_fetch = Hook()) {
    typedef FftGeom<NFFT> G;
    constexpr int NJ = JHI - JLO;
    // (holding the NEXT pair in registers was measured three ways -- fetched before the output stage, fetched
    //  after the power values went to shared memory so that the loads fly under the mel loop, and at 168
    //  registers / 12 warps: it spills or gains nothing; prefetch.global.L1 has no effect either.  The stall on
    //  the first use of the samples stays at ~6 % of warp time.)
    RawFrame<FAM, NJ> rawA, rawB;
    fetch_frame<FAM, JLO, JHI, SAMP, SMEM>(p, clipA, baseA, l, rawA);
    // FBANK only (its frames cost two loads per sample pair: the pair and the sample before it).  Measured on the B200 with the
    // sharing enabled for every family (16,384 clips, M clips/s): R-FBANK 18.35 -> 19.43, C-FBANK 18.41 -> 19.17, R-SPEC TF 26.8 ->
    // 27.3, but C-SPEC TF 35.2 -> 33.7, C-MFCC 15.24 -> 14.57, R-MFCC 17.2 -> 15.8: there the copies and the branch cost more
    // than the loads they replace (which hit L1 anyway).
    const bool after = FAM == FAM_FBANK && !SMEM && clipA == clipB && !rawA.final_ && baseB - baseA == 160 && 5 < NJ &&
                       baseB + 32 * JHI <= p.n_samples;
    if (after) fetch_frame_after<FAM, JLO, JHI, 5, SAMP>(p, clipB, baseB, l, rawA, rawB);
    else fetch_frame<FAM, JLO, JHI, SAMP, SMEM>(p, clipB, baseB, l, rawB);
    C2 v[G::V];
    window_pair<NFFT, FAM, JLO, JHI>(p, rawA, rawB, l, s_win, v);
    after_fetch();
    // The two planes of every exchange pass through the same slots: put / sync / get, twice.
    fft_pass1<NFFT>(v, l, T);
    C2 w[G::V];
    __syncwarp();                               // the previous pair's readers are done with the tile
    xs_put<NFFT, 0>(v, l, xb);
    __syncwarp();
    if (NFFT == 512) xs_get_512<0>(l, xb, w); else xs_get2_640<0>(l, xb, w);
    __syncwarp();
    xs_put<NFFT, 1>(v, l, xb);
    __syncwarp();
    if (NFFT == 512) {
        xs_get_512<1>(l, xb, w);
        dft16(w);                               // w[k2] = Z[l + 16 k2]
        return fft_untangle_512_shfl(l, lane, w, T, pa, pb);
    } else {
        xs_get2_640<1>(l, xb, w);
        fft_pass2_640(l, w, T);
        __syncwarp();
        xs_put3_640<0>(l, w, xb);
        __syncwarp();
        xs_get3_640<0>(l, xb, v);
        __syncwarp();
        xs_put3_640<1>(l, w, xb);
        __syncwarp();
        xs_get3_640<1>(l, xb, v);
        fft_pass3_640(v);
        return fft_untangle_640_shfl(l, lane, v, T, pa, pb);
    }
}

// --------------------------------------------------------------------------------
// sparse triangular band sums over the packed power buffer: one filter per lane and 16-filter group, uniform trip
// count per group (ELL); emit(m, acc, guard) receives filter m's (frame A, frame B) sums.
//   NG / CODE: compile-time bank shape (NG groups, 2 bits per group = float4 steps - 1), 0 = runtime metadata
// --------------------------------------------------------------------------------
template <int NG, unsigned CODE, int BATCH = 1, typename Emit>
__device__ __forceinline__ void mel_project(const P2* pbuf, const int2* g_meta, int n_fgroups, const int* f_start,
                                            const float2* f_w2, int l, Emit emit) {
    if (NG > 0) {
        // BATCH groups at a time: first their band sums (independent accumulator chains the scheduler can interleave), then
        // their emits.  With emit() inside the group loop (BATCH = 1) every group's power reads wait for the previous group's
        // shared-memory store (the compiler must assume they alias): eight load -> 16-deep FFMA2 chain -> log -> store
        // sequences in a row per frame pair.  Same arithmetic in the same order either way (bit-identical results).
        // Measured (B200, tcgen05 MFCC kernel, BATCH 1 / 2 / 4 / 8): C-MFCC 15.24 / 15.42 / 15.57 / 15.63 M clips/s at 262,144
        // clips -- but C-MFCC-D2 11.79 / 11.20 / 11.24 / 11.11 and R-MFCC 17.1 / 16.6 / 16.6 / 16.8: where the read-out warps
        // set the pace (deltas), burstier frame warps take issue slots from them.  The dispatcher picks per launch.
        constexpr int HALF = NG > BATCH ? BATCH : (NG > 0 ? NG : 1);
        int off4 = 0;                       // all of this folds at compile time
#pragma unroll
        for (int h0 = 0; h0 < NG; h0 += HALF) {
            P2 accs[HALF];
#pragma unroll
            for (int i = h0; i < h0 + HALF && i < NG; ++i) {
                const int n4 = (int)((CODE >> (2 * i)) & 3u) + 1;
                const int m = 16 * i + l;
                const P2* pq = pbuf + f_start[m];
                const float2* wq = f_w2 + off4 * 32 + l;
                P2 acc = bc(0.f);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    if (q4 < n4) {
                        const float2 wa = wq[q4 * 32], wb = wq[q4 * 32 + 16];
                        acc = pfma(pq[4 * q4 + 0], bc(wa.x), acc);
                        acc = pfma(pq[4 * q4 + 1], bc(wa.y), acc);
                        acc = pfma(pq[4 * q4 + 2], bc(wb.x), acc);
                        acc = pfma(pq[4 * q4 + 3], bc(wb.y), acc);
                    }
                }
                off4 += n4;
                accs[i - h0] = acc;
            }
#pragma unroll
            for (int i = h0; i < h0 + HALF && i < NG; ++i) emit(16 * i + l, accs[i - h0], i == NG - 1);
        }
    } else {
        for (int i = 0; i < n_fgroups; ++i) {
            const int2 gm = g_meta[i];
            const int m = 16 * i + l;
            const P2* pq = pbuf + f_start[m];
            const float2* wq = f_w2 + gm.x * 32 + l;
            P2 acc = bc(0.f);
            for (int q4 = 0; q4 < gm.y; ++q4) {
                const float2 wa = wq[q4 * 32], wb = wq[q4 * 32 + 16];
                acc = pfma(pq[4 * q4 + 0], bc(wa.x), acc);
                acc = pfma(pq[4 * q4 + 1], bc(wa.y), acc);
                acc = pfma(pq[4 * q4 + 2], bc(wb.x), acc);
                acc = pfma(pq[4 * q4 + 3], bc(wb.y), acc);
            }
            emit(m, acc, true);
        }
    }
}

// --------------------------------------------------------------------------------
// folded DCT-II on the FP32 pipe (see the epilogue comment in the kernel)
//   work item = (block of CB same-parity coefficients, PQ frame pairs); one THREAD per item, items dealt over the
//   whole CTA, nothing is combined across lanes.  Per f the thread reads its pairs' folded values (PQ LDS.64) and its
//   block's CB coefficients (LDS.128 [+ LDS], the same address for all lanes of a block) and issues PQ CB FFMA2.
//   Table layout: a float4 plane [block][f < n/2] with coefficients 0..3 of the block, then a plane of 1 / 2 / 4 floats
//   per entry with coefficients 4.. (CB > 4); even-k blocks first.
//   The phase is bound by shared-memory wavefronts or by FFMA2 issue depending on the tile; the host picks (CB, PQ)
//   from a small cost model (srfe_abi.cu: smem_plan).
//   Measured alternatives, all slower: warp tasks of 4 coefficients x 64 pairs with the f range split over half-warps
//   (10 tasks for 13-14 warps, ~550 instructions of per-task set-up / shuffle combine / stores against 512 FFMA2);
//   5-coefficient blocks with shared-memory atomics; the table read through L1 from global memory.
// --------------------------------------------------------------------------------
template <int CB, int PQ>
__device__ __forceinline__ void dct_items(const KParams& p, const P2* tileP, const float* dtab, const P2* fmeanP,
                                          float* ctile, float thr, int npairs, int TC, int tid, int nthr) {
    constexpr int RB = CB <= 4 ? 0 : CB == 5 ? 1 : CB == 6 ? 2 : 4;    // floats per entry in the second plane
    const int half = p.n_filt >> 1, TSP = p.tile_stride;
    const int nq = (npairs + PQ - 1) / PQ;                              // a thread's pairs: q, q + nq, ... (row stride TSP
    const int nqp = (nq + 15) & ~15;                                    //  between lanes keeps the LDS.64 conflict-free);
    const int nitems = (p.dct_nbe + p.dct_nbo) * nqp;                   // a half-warp never straddles two blocks
    for (int it = tid; it < nitems; it += nthr) {
        const int b = it / nqp, q0 = it - b * nqp;
        if (q0 >= nq) continue;
        const int par = b >= p.dct_nbe ? 1 : 0;
        const P2* x[PQ];
#pragma unroll
        for (int i = 0; i < PQ; ++i) x[i] = tileP + min(q0 + i * nq, npairs - 1) * TSP + par * half;   // s[0..half) | d[0..half)
        const float4* d = reinterpret_cast<const float4*>(dtab) + b * half;
        const float* e = dtab + 4 * (p.dct_nbe + p.dct_nbo) * half + RB * b * half;
        P2 acc[PQ][CB];
#pragma unroll
        for (int i = 0; i < PQ; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) acc[i][j] = bc(0.f);
#pragma unroll 4
        for (int f = 0; f < half; ++f) {
            float c[8];
            const float4 d0 = d[f];
            c[0] = d0.x; c[1] = d0.y; c[2] = d0.z; c[3] = d0.w;
            if (RB == 1) c[4] = e[f];
            if (RB == 2) { const float2 d1 = *reinterpret_cast<const float2*>(e + 2 * f); c[4] = d1.x; c[5] = d1.y; }
            if (RB == 4) { const float4 d1 = *reinterpret_cast<const float4*>(e + 4 * f); c[4] = d1.x; c[5] = d1.y; c[6] = d1.z; c[7] = d1.w; }
#pragma unroll
            for (int i = 0; i < PQ; ++i) {
                const P2 xv = x[i][f];
#pragma unroll
                for (int j = 0; j < CB; ++j) acc[i][j] = pfma(xv, bc(c[j]), acc[i][j]);
            }
        }
        const int k0 = 2 * CB * (b - par * p.dct_nbe) + par;
#pragma unroll
        for (int i = 0; i < PQ; ++i) {
            const int q = q0 + i * nq;
            if (q < npairs) {
                if (k0 == 0) {                                          // put the frame's centre back on c0
                    const P2 cm = fmeanP[q];
                    acc[i][0] = pfma(mkp(fmaxf(cm.lo, thr), fmaxf(cm.hi, thr)), bc(p.dct_row0_sum), acc[i][0]);
                }
                float* cq = ctile + 2 * q;
                const bool two = 2 * q + 1 < p.T;
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    const int k = k0 + 2 * j;
                    if (k < p.n_mfcc) {
                        cq[k * TC] = acc[i][j].lo;
                        if (two) cq[k * TC + 1] = acc[i][j].hi;
                    }
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------
// the fused kernel
// --------------------------------------------------------------------------------
// NG / CODE: compile-time shape of the mel ELL bank (NG 16-filter groups, 2 bits per group = float4 steps - 1)
// for the known presets, so the projection unrolls into straight-line code; NG = 0 -> runtime metadata.
// tables -> shared memory, once per (persistent) CTA; the caller synchronises the CTA afterwards
template <int FAM>
__device__ __forceinline__ void srfe_load_tables(const KParams& p, unsigned char* smem) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    {
        const int4* src = reinterpret_cast<const int4*>(p.blob);
        int4* dst = reinterpret_cast<int4*>(smem);
        for (int i = tid; i < p.blob_bytes / 16; i += nthr) dst[i] = __ldg(src + i);
        if (FAM == FAM_MFCC && p.dct_fold) {                // the chosen block-size variant of the folded DCT table
            const int4* vs = reinterpret_cast<const int4*>(p.blob + p.dct_src);
            int4* vd = reinterpret_cast<int4*>(smem + p.off_dfold);
            for (int i = tid; i < p.dct_bytes / 16; i += nthr) vd[i] = __ldg(vs + i);
        }
    }
    {   // The padded mel runs read (with zero weights) a few power-buffer slots that neither the exchange nor the power
        // store ever writes: give them a finite value once, 0 x stale NaN bits would poison a band sum.
        int4* z = reinterpret_cast<int4*>(smem + p.sm_scratch);
        for (int i = tid; i < (p.sm_tile - p.sm_scratch) / 16; i += nthr) z[i] = make_int4(0, 0, 0, 0);
    }
}

// the clip groups grp_begin, grp_begin + grp_step, ... < grp_end of one feature family; `smem` = this family's region
// ([tables][scratch][tile], offsets in p).  srfe_kernel runs one family over the CTA's share of all groups; the fused
// kernel (below) alternates two families group by group so that the second one finds the PCM in L1 / L2.
// GRID = true: the CTA's grid-stride share of all groups (bounds read from blockIdx / gridDim / p, not held in registers)
template <int NFFT, int FAM, int JLO, int JHI, int NG, unsigned CODE, typename SAMP, bool GRID, bool STAGED = false>
__device__ __forceinline__ void srfe_groups(const KParams& p, unsigned char* smem_region, int grp_begin, int grp_end, int grp_step) {
    extern __shared__ __align__(128) unsigned char smem_base[];
    unsigned char* smem = GRID ? smem_base : smem_region;   // single-family kernel: addresses fold to constants, no base register
    const SAMP* pcm = reinterpret_cast<const SAMP*>(p.pcm);
    typedef FftGeom<NFFT> G;
    constexpr int F = G::M + 1;
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int HW = nthr >> 4;

    const float* s_win = reinterpret_cast<const float*>(smem + p.off_win);
    FftTables T;
    T.tw1 = reinterpret_cast<const cpx*>(smem + p.off_tw1);
    T.twu = reinterpret_cast<const cpx*>(smem + p.off_twu);
    T.tw16 = reinterpret_cast<const cpx*>(smem + p.off_tw16);
    const int2* g_meta = reinterpret_cast<const int2*>(smem + p.off_gm);
    const int* f_start = reinterpret_cast<const int*>(smem + p.off_fs);
    const float2* f_w2 = reinterpret_cast<const float2*>(smem + p.off_fw4);     // [step][plane][lane]: two LDS.64 per step
    P2* scratch_all = reinterpret_cast<P2*>(smem + p.sm_scratch);
    float* tile = reinterpret_cast<float*>(smem + p.sm_tile);
    // MFCC only: the per-clip dB tile holds one row per frame PAIR, each element an (A, B) pair -- exactly what the
    // mel stage produces and what the packed DCT consumes; row stride p.tile_stride (P2 units, odd); +1 dummy row
    P2* tileP = reinterpret_cast<P2*>(tile);
    P2* fmeanP = tileP + (((p.T + 1) >> 1) + 1) * p.tile_stride;      // per-pair frame means, right after the tile
    const float* s_dfold = reinterpret_cast<const float*>(smem + p.off_dfold);

    const int hw = tid >> 4, l = tid & 15, lane = tid & 31;
    P2* xb = scratch_all + hw * G::SCRATCH_P2;           // this half-warp's exchange tile / packed power buffer
    // STAGED (spectrogram): each half-warp owns a 2-frame staging buffer in shared memory and an mbarrier; one lane issues
    // two cp.async.bulk (TMA, 1-D) copies for the half-warp's NEXT frame pair as soon as the current pair's samples are in
    // registers, so the copy flies under the FFT and the next pair's samples come from shared memory, not from L1 / L2
    SAMP* stg = reinterpret_cast<SAMP*>(smem + p.sm_stage) + hw * 2 * NFFT;
    const unsigned stg_mb = (unsigned)__cvta_generic_to_shared(smem + p.sm_stage_mb + 8 * hw);
    unsigned stg_phase = p.stage_phase0;                 // (the wrapper kernel initialises the barriers)
    auto stage_issue = [&](const SAMP* fa, const SAMP* fb) {
        constexpr unsigned kBytes = NFFT * sizeof(SAMP);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stg);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(stg_mb), "r"(2 * kBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(fa), "r"(kBytes), "r"(stg_mb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst + kBytes), "l"(fb), "r"(kBytes), "r"(stg_mb) : "memory");
    };

    for (int grp = GRID ? (int)blockIdx.x : grp_begin; grp < (GRID ? p.n_groups : grp_end); grp += GRID ? (int)gridDim.x : grp_step) {
        const int clip0 = grp * p.cpc;
        const int ncl = min(p.cpc, p.n_clips - clip0);
        const int nf = ncl * p.T;                           // flattened frames of this group
        const int npairs = (nf + 1) >> 1;
        const int rounds = (npairs + HW - 1) / HW;
        float run_max = -CUDART_INF_F;

        bool staged_ahead = false;                          // STAGED: the copy for this iteration's pair is already in flight
        for (int it = 0; it < rounds; ++it) {
            const int q = it * HW + hw;
            P2 pa[G::M / 32], pb[G::M / 32], pmid;
            const bool active = (it * HW + (hw & ~1)) < npairs;
            if (active) {                                   // warp-uniform: both half-warps of a warp run together
                const FramePos cA = frame_pos(2 * q, nf, p.T, p.t_magic, p.cpc), cB = frame_pos(2 * q + 1, nf, p.T, p.t_magic, p.cpc);
                if (STAGED) {
                    auto frame_ptr = [&](const FramePos& c) { return pcm + (long long)(clip0 + c.c) * p.clip_stride + p.start0 + c.t * p.hop; };
                    if (!staged_ahead && l == 0) stage_issue(frame_ptr(cA), frame_ptr(cB));
                    {   // the pair's samples have landed
                        unsigned done = 0;
                        while (!done)
                            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                                         : "=r"(done) : "r"(stg_mb), "r"(stg_phase) : "memory");
                        stg_phase ^= 1u;
                    }
                    const int qn = q + HW;                  // this half-warp's pair of the next round (same group)
                    const bool more = (it + 1 < rounds) && ((it + 1) * HW + (hw & ~1)) < npairs;
                    staged_ahead = more;
                    auto refill = [&]() {
                        __syncwarp();                       // every lane of the half-warp holds its samples in registers
                        if (more && l == 0) {
                            const FramePos nA = frame_pos(2 * qn, nf, p.T, p.t_magic, p.cpc), nB = frame_pos(2 * qn + 1, nf, p.T, p.t_magic, p.cpc);
                            stage_issue(frame_ptr(nA), frame_ptr(nB));
                        }
                    };
                    pmid = pair_power<NFFT, FAM, JLO, JHI, SAMP, true>(p, stg, 0, stg + NFFT, 0, l, lane, s_win, T, xb, pa, pb, refill);
                } else {
                pmid = pair_power<NFFT, FAM, JLO, JHI, SAMP>(p, pcm + (long long)(clip0 + cA.c) * p.clip_stride, p.start0 + cA.t * p.hop,
                                                              pcm + (long long)(clip0 + cB.c) * p.clip_stride, p.start0 + cB.t * p.hop,
                                                              l, lane, s_win, T, xb, pa, pb);
                }

                // bin held in slot r of this lane (and its mirror M - k): natural stride-16 order for N = 512,
                // the radix-4 order of bin640() for N = 640
                auto kbin = [&](int r) { return NFFT == 512 ? l + 16 * r : bin640(l, r); };

                // Stores to global rows are NOT guarded by cA.ok / cB.ok: a frame slot past the end of the group aliases the
                // group's last valid frame (frame_pos), so it recomputes that frame bit for bit and rewrites the same row with
                // the same values -- cheaper than a branch around every store.
                if (FAM == FAM_SPEC) {
                    // density scaling, one-sided doubling (not DC / Nyquist), optional ln(. + eps)
                    const float s2 = 2.f * p.scale;
                    if (p.layout == SRFE_LAYOUT_TF) {
                        float* rowA = p.out + ((long long)(clip0 + cA.c) * p.T + cA.t) * F;
                        float* rowB = p.out + ((long long)(clip0 + cB.c) * p.T + cB.t) * F;
                        P2* pbuf = xb;                              // N = 640 only: re-order through shared memory
#pragma unroll
                        for (int r = 0; r < G::M / 32; ++r) {
                            const int k = kbin(r);
                            const float sc = (r == 0 && l == 0) ? p.scale : s2;
                            float a0, a1, b0, b1;
                            if (p.take_log) {               // ln(S + eps) = ln2 * lg2(fma(P, scale, eps))
                                a0 = 0.6931471805599453f * lg2_ftz(fmaf(pa[r].lo, sc, p.log_eps));
                                a1 = 0.6931471805599453f * lg2_ftz(fmaf(pa[r].hi, sc, p.log_eps));
                                b0 = 0.6931471805599453f * lg2_ftz(fmaf(pb[r].lo, sc, p.log_eps));
                                b1 = 0.6931471805599453f * lg2_ftz(fmaf(pb[r].hi, sc, p.log_eps));
                            } else {
                                a0 = pa[r].lo * sc; a1 = pa[r].hi * sc; b0 = pb[r].lo * sc; b1 = pb[r].hi * sc;
                            }
                            if (NFFT == 512) {              // lanes hold consecutive bins: store straight from registers
                                rowA[k] = a0; rowA[G::M - k] = b0;
                                rowB[k] = a1; rowB[G::M - k] = b1;
                            } else {
                                pbuf[k] = mkp(a0, a1);
                                pbuf[G::M - k] = mkp(b0, b1);
                            }
                        }
                        float c0 = pmid.lo * s2, c1 = pmid.hi * s2;
                        if (p.take_log) { c0 = 0.6931471805599453f * lg2_ftz(c0 + p.log_eps); c1 = 0.6931471805599453f * lg2_ftz(c1 + p.log_eps); }
                        if (NFFT == 512) {
                            if (l == 0) {
                                rowA[G::M / 2] = c0;
                                rowB[G::M / 2] = c1;
                            }
                        } else {
                            if (l == 0) pbuf[G::M / 2] = mkp(c0, c1);
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < G::M / 16; ++r) {       // natural order: 64-byte runs per half-warp
                                const P2 q2 = pbuf[l + 16 * r];
                                rowA[l + 16 * r] = q2.lo;
                                rowB[l + 16 * r] = q2.hi;
                            }
                            if (l == 0) {
                                const P2 q2 = pbuf[G::M];
                                rowA[G::M] = q2.lo;
                                rowB[G::M] = q2.hi;
                            }
                            __syncwarp();
                        }
                    } else {
                        // FT: keep the scaled / logged values in registers; they go through a CTA-wide
                        // [bin][frames of this round] tile after the round's barrier (below)
                        const float s2b = 2.f * p.scale;
#pragma unroll
                        for (int r = 0; r < G::M / 32; ++r) {
                            const float sc = (r == 0 && l == 0) ? p.scale : s2b;
                            if (p.take_log) {
                                pa[r] = pfma(pa[r], bc(sc), bc(p.log_eps));
                                pb[r] = pfma(pb[r], bc(sc), bc(p.log_eps));
                                pa[r] = mkp(0.6931471805599453f * lg2_ftz(pa[r].lo), 0.6931471805599453f * lg2_ftz(pa[r].hi));
                                pb[r] = mkp(0.6931471805599453f * lg2_ftz(pb[r].lo), 0.6931471805599453f * lg2_ftz(pb[r].hi));
                            } else {
                                pa[r] = pmul(pa[r], bc(sc));
                                pb[r] = pmul(pb[r], bc(sc));
                            }
                        }
                        pmid = pmul(pmid, bc(s2b));
                        if (p.take_log) pmid = mkp(0.6931471805599453f * lg2_ftz(pmid.lo + p.log_eps), 0.6931471805599453f * lg2_ftz(pmid.hi + p.log_eps));
                    }
                } else {
                    // power -> shared as (A, B) pairs (aliases the FFT scratch), then sparse triangular sums:
                    // 16-filter groups with a uniform trip count (ELL), weights as float4 runs
                    P2* pbuf = xb;
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < G::M / 32; ++r) {
                        const int k = kbin(r);
                        pbuf[k] = pa[r];
                        pbuf[G::M - k] = pb[r];
                    }
                    if (l == 0) pbuf[G::M / 2] = pmid;
                    __syncwarp();
                    float* orowA = p.out + ((long long)(clip0 + cA.c) * p.T + cA.t) * p.n_filt;     // FBANK rows
                    float* orowB = p.out + ((long long)(clip0 + cB.c) * p.T + cB.t) * p.n_filt;
                    P2* prow = tileP + min(q, npairs) * p.tile_stride;                            // MFCC pair row (dummy: npairs)
                    P2 fsum = bc(0.f);
                    // one filter per lane and group; emit() turns the band sum into dB and stores it
                    auto emit = [&](int m, const P2& acc, bool guard) {
                        if (guard && m >= p.n_filt) return;
                        if (FAM == FAM_FBANK) {
                            float a = acc.lo, b = acc.hi;
                            // model_fbanks_cnn.py:61: exact zeros -> eps; denormal sums are lifted to FLT_MIN so the
                            // flush-to-zero log never returns -inf (far below any level the front end can resolve)
                            a = (a == 0.f) ? 2.220446049250313e-16f : fmaxf(a, 1.17549435e-38f);
                            b = (b == 0.f) ? 2.220446049250313e-16f : fmaxf(b, 1.17549435e-38f);
                            orowA[m] = 6.020599913279624f * lg2_ftz(a);                 // 20 log10
                            orowB[m] = 6.020599913279624f * lg2_ftz(b);
                        } else {
                            // log2 units: the factor 10 log10(2) of power_to_db lives in the DCT tables and in top_db
                            // (host, double precision) -- the DCT is linear and the clamp only compares
                            const float da = lg2_ftz(fmaxf(acc.lo, p.amin));
                            const float db = lg2_ftz(fmaxf(acc.hi, p.amin));
                            prow[m] = mkp(da, db);              // one STS.64 per filter and frame pair
                            // (frames past the end alias the clip's last valid frame -- frame_pos -- so their values are
                            //  values the maximum has seen anyway: no masking needed)
                            run_max = fmaxf(run_max, fmaxf(da, db));
                            fsum = padd(fsum, mkp(da, db));
                        }
                    };
                    mel_project<NG, CODE>(pbuf, g_meta, p.n_fgroups, f_start, f_w2, l, emit);
                    if (FAM == FAM_MFCC) {                   // per-frame mean dB: centre of the DCT accumulation
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) {
                            fsum.lo += __shfl_xor_sync(0xffffffffu, fsum.lo, o);
                            fsum.hi += __shfl_xor_sync(0xffffffffu, fsum.hi, o);
                        }
                        if (l == 0) fmeanP[min(q, npairs)] = pmul(fsum, bc(1.f / (float)p.n_filt));
                    }
                    __syncwarp();
                }
            }
            if (FAM == FAM_SPEC && p.layout == SRFE_LAYOUT_FT) {
                // CTA-wide tile [bin][2 HW frames] aliased onto the (now idle) FFT scratch; rows of up to
                // 2 HW consecutive frames leave as long runs along time
                float* ft = reinterpret_cast<float*>(scratch_all);
                const int RS = 2 * HW + 2;                  // even: the (A, B) pair of a bin goes out as one STS.64; 2 k (mod 32)
                                                            // puts the 16 bins of a half-warp on 16 distinct bank pairs
                __syncthreads();                            // every half-warp is done with its Z buffer
                if (active) {
                    float* col = ft + 2 * hw;
#pragma unroll
                    for (int r = 0; r < G::M / 32; ++r) {
                        const int k = NFFT == 512 ? l + 16 * r : bin640(l, r);
                        *reinterpret_cast<P2*>(col + k * RS) = pa[r];
                        *reinterpret_cast<P2*>(col + (G::M - k) * RS) = pb[r];
                    }
                    if (l == 0) *reinterpret_cast<P2*>(col + (G::M / 2) * RS) = pmid;
                }
                __syncthreads();
                const int f0 = 2 * it * HW;
                for (int cb = 0; cb < 2 * HW; cb += 32) {
                    const FramePos po = frame_pos(f0 + cb + lane, nf, p.T, p.t_magic, p.cpc);
                    if (cb + lane < 2 * HW && po.ok) {
                        float* oc = p.out + (long long)(clip0 + po.c) * F * p.T + po.t;
                        const float* src = ft + cb + lane;
                        for (int k = tid >> 5; k < F; k += (nthr >> 5)) oc[(long long)k * p.T] = src[k * RS];
                    }
                }
                __syncthreads();                            // the tile is scratch again in the next round
            }
        }

        if (FAM == FAM_MFCC && !SRFE_DBG(p, 1)) {
            __shared__ float s_red[kMaxThreads / 32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) run_max = fmaxf(run_max, __shfl_xor_sync(0xffffffffu, run_max, o));
            if ((tid & 31) == 0) s_red[tid >> 5] = run_max;
            __syncthreads();
            float gmax = (lane < (nthr >> 5)) ? s_red[lane] : -CUDART_INF_F;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
            const float thr = p.top_db >= 0.f ? gmax - p.top_db : -CUDART_INF_F;   // power_to_db(top_db): max over the clip (log2 units)
            // DCT-II on values re-centred per frame: with c_t = max(mean_f dB[t][f], thr),
            //   C[k][t] = sum_f D[k][f] (x[t][f] - c_t) + c_t * sum_f D[k][f],   sum_f D[k][f] = sqrt(n_mels) [k == 0]
            // so the accumulated magnitudes are the within-frame spread, not c0 ~ 1e3.
            //
            // The contraction runs on the FP32 pipe in packed FFMA2 (both frames of a pair at once), NOT on tensor
            // cores: measured on B200 (scripts/ubench/mma_overlap.cu), the legacy mma.sync TF32 path retires one
            // HMMA.1688 per ~16 cycles per SMSP, i.e. ~65 fp32-grade MAC/clk/SM after the 3xTF32 split fp32 accuracy
            // needs, against 128 MAC/clk/SM for FFMA2.  It is a pipe of its own, and a version of this contraction on
            // it (overlapping the co-resident CTA's FFT) was built and measured at parity: no gain
            // (profiles/r1_notes.md).  tcgen05 would need split operand tiles that do not fit beside the dB tile.
            // The DCT-II symmetry D[k][n-1-f] = (-1)^k D[k][f] halves the MACs: even k see s = x[f] + x[n-1-f],
            // odd k see d = x[f] - x[n-1-f], f < n/2.
            const int TC = (p.layout == SRFE_LAYOUT_FT) ? p.T : p.T + 1 + (p.T & 1);
            const int warp = tid >> 5, nwarps = nthr >> 5;
            float* ctile = reinterpret_cast<float*>(smem + p.sm_ctile);
            const int TSP = p.tile_stride;
            if (p.dct_fold) {
                const int half = p.n_filt >> 1;
                // (1) clamp, re-centre, fold -- in place: s[f] -> [f], d[f] -> [n/2 + f].  One item folds f and
                //     n/2-1-f together: it reads and writes the same four slots, so no other thread's input is touched.
                const int hq = half >> 1;
                for (int idx = tid; idx < (SRFE_DBG(p, 16) ? 0 : npairs * hq); idx += nthr) {
                    const int q = idx / hq, f = idx - q * hq, g = half - 1 - f;
                    P2* row = tileP + q * TSP;
                    const P2 cm = fmeanP[q];
                    const P2 c = mkp(fmaxf(cm.lo, thr), fmaxf(cm.hi, thr));
                    const P2 a0 = row[f], a1 = row[p.n_filt - 1 - f], b0 = row[g], b1 = row[p.n_filt - 1 - g];
                    const P2 ya0 = psub(mkp(fmaxf(a0.lo, thr), fmaxf(a0.hi, thr)), c);
                    const P2 ya1 = psub(mkp(fmaxf(a1.lo, thr), fmaxf(a1.hi, thr)), c);
                    const P2 yb0 = psub(mkp(fmaxf(b0.lo, thr), fmaxf(b0.hi, thr)), c);
                    const P2 yb1 = psub(mkp(fmaxf(b1.lo, thr), fmaxf(b1.hi, thr)), c);
                    row[f] = padd(ya0, ya1);
                    row[g] = padd(yb0, yb1);
                    row[half + f] = psub(ya0, ya1);
                    row[half + g] = psub(yb0, yb1);
                }
                __syncthreads();
                // (2) the contraction
#define SRFE_DCT(CB_, PQ_) dct_items<CB_, PQ_>(p, tileP, s_dfold, fmeanP, ctile, thr, npairs, TC, tid, nthr)
                if (!SRFE_DBG(p, 2)) switch (p.dct_cb * 4 + p.dct_pq) {
                    case 2 * 4 + 1: SRFE_DCT(2, 1); break;   case 2 * 4 + 2: SRFE_DCT(2, 2); break;
                    case 3 * 4 + 1: SRFE_DCT(3, 1); break;   case 3 * 4 + 2: SRFE_DCT(3, 2); break;
                    case 4 * 4 + 1: SRFE_DCT(4, 1); break;   case 4 * 4 + 2: SRFE_DCT(4, 2); break;
                    case 5 * 4 + 1: SRFE_DCT(5, 1); break;   case 5 * 4 + 2: SRFE_DCT(5, 2); break;
                    case 6 * 4 + 1: SRFE_DCT(6, 1); break;   case 6 * 4 + 2: SRFE_DCT(6, 2); break;
                    case 8 * 4 + 1: SRFE_DCT(8, 1); break;   default:        SRFE_DCT(8, 2); break;
                }
#undef SRFE_DCT
            } else {
                // generic path (odd n_mels): plain packed dot products against the global DCT rows
                for (int task = tid; task < p.n_mfcc * npairs; task += nthr) {
                    const int k = task / npairs, q = task - k * npairs;
                    const P2* row = tileP + q * TSP;
                    const P2 cm = fmeanP[q];
                    const P2 c = mkp(fmaxf(cm.lo, thr), fmaxf(cm.hi, thr));
                    const float* dk = p.dct_kf + k * p.n_filt;
                    P2 acc = bc(0.f);
                    for (int f = 0; f < p.n_filt; ++f) {
                        const P2 x = row[f];
                        acc = pfma(psub(mkp(fmaxf(x.lo, thr), fmaxf(x.hi, thr)), c), bc(__ldg(dk + f)), acc);
                    }
                    if (k == 0) acc = pfma(c, bc(p.dct_row0_sum), acc);
                    ctile[k * TC + 2 * q] = acc.lo;
                    if (2 * q + 1 < p.T) ctile[k * TC + 2 * q + 1] = acc.hi;
                }
            }
            __syncthreads();
            // np.gradient along time (unit spacing, edge_order 1), applied n_deltas times
            for (int d = 1; d <= p.n_deltas; ++d) {
                const float* src = ctile + (d - 1) * p.n_mfcc * TC;
                float* dst = ctile + d * p.n_mfcc * TC;
                for (int k = warp; k < p.n_mfcc; k += nwarps) {
                    const float* sr = src + k * TC;
                    for (int t = lane; t < p.T; t += 32) {
                        float gr;
                        if (t == 0) gr = sr[1] - sr[0];
                        else if (t == p.T - 1) gr = sr[t] - sr[t - 1];
                        else gr = 0.5f * (sr[t + 1] - sr[t - 1]);
                        dst[k * TC + t] = gr;
                    }
                }
                __syncthreads();
            }
            const int R = (1 + p.n_deltas) * p.n_mfcc;
            float* oc = p.out + (long long)clip0 * R * p.T;
            if (p.layout == SRFE_LAYOUT_FT) {
                const int n = SRFE_DBG(p, 8) ? 0 : R * p.T;   // the clip's features are one contiguous block
                if (((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(oc) & 15) == 0)) {
                    const float4* c4 = reinterpret_cast<const float4*>(ctile);
                    float4* o4 = reinterpret_cast<float4*>(oc);
                    for (int i = tid; i < (n >> 2); i += nthr) o4[i] = c4[i];
                } else {
                    for (int i = tid; i < n; i += nthr) oc[i] = ctile[i];
                }
            } else {
                for (int t = warp; t < p.T; t += nwarps) {
                    float* orow = oc + t * R;
                    const float* ccol = ctile + t;
                    for (int r = lane; r < R; r += 32) orow[r] = ccol[r * TC];
                }
            }
            __syncthreads();                                // scratch / tile are reused by the next clip
        }
    }
}

template <int NFFT, int FAM, int JLO, int JHI, int NG, unsigned CODE, typename SAMP>
__global__ void __launch_bounds__(kMaxThreads, 1) srfe_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    srfe_load_tables<FAM>(p, smem);
    __syncthreads();
    srfe_groups<NFFT, FAM, JLO, JHI, NG, CODE, SAMP, true>(p, smem, 0, 0, 0);
}

// spectrogram with TMA-staged frames (see srfe_groups: STAGED)
template <int NFFT, typename SAMP>
__global__ void __launch_bounds__(kMaxThreads, 1) srfe_spec_staged_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    srfe_load_tables<FAM_SPEC>(p, smem);
    if (threadIdx.x < (blockDim.x >> 4)) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(smem + p.sm_stage_mb + 8 * threadIdx.x);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    srfe_groups<NFFT, FAM_SPEC, 0, NFFT / 32, 0, 0u, SAMP, true, true>(p, smem, 0, 0, 0);
}

// --------------------------------------------------------------------------------
// spectrogram + log-fbank of the SAME clips in one launch: the reference's ensemble feeds one batch to
// model_spec_* and model_fbanks_cnn in turn (analyst_training.py:91-94, predictions.py:58-60), each recomputing its
// features from the PCM.  Here a CTA takes a group of clips through the spectrogram frames and then, straight away,
// through the fbank frames (different framing: 640 / 320 vs 400 / 160, so the FFTs are not shared): the second pass finds
// the group's samples in L1 / L2, HBM delivers each sample once for both outputs.  Both families keep their own
// [tables][scratch] region of shared memory; pa.cpc == pb.cpc and pa.n_groups == pb.n_groups (host).
// --------------------------------------------------------------------------------
template <int NA, int NB, int JLO_B, int JHI_B, int NG_B, unsigned CODE_B, typename SAMP>
__global__ void __launch_bounds__(kMaxThreads, 1) srfe_spec_fbank_kernel(const KParams pa, const KParams pb, int off_b) {
    extern __shared__ __align__(128) unsigned char smem[];
    srfe_load_tables<FAM_SPEC>(pa, smem);
    srfe_load_tables<FAM_FBANK>(pb, smem + off_b);
    __syncthreads();
    for (int grp = blockIdx.x; grp < pa.n_groups; grp += gridDim.x) {
        srfe_groups<NA, FAM_SPEC, 0, NA / 32, 0, 0u, SAMP, false>(pa, smem, grp, grp + 1, 1);
        srfe_groups<NB, FAM_FBANK, JLO_B, JHI_B, NG_B, CODE_B, SAMP, false>(pb, smem + off_b, grp, grp + 1, 1);
    }
}

}  // namespace srfe
