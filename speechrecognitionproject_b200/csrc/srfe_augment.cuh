// srfe_augment.cuh -- on-device training-time PCM augmentation and silence synthesis.
//
// Restates, with a counter-based random stream, what the reference's Dataset does on the host one clip at a time
// (/root/reference/dataset.py): generate_silence_sample :148-161, add_noise_snr :163-183, add_noise_uniform :185-191,
// time_stretching :193-202 and the band selection of __getitem__ :107-116.  One CTA per clip; the clip stays on the
// device and feeds the fused front end on the same stream.  Every draw is Philox4x32-10 keyed by the seed and counted
// by (global clip index, block), the layout documented in oracle/augment.py, so the numpy oracle reproduces the output
// bit for bit; sample arithmetic follows the reference's dtypes (float64 multiply, then add, np.int16 truncation with
// wrap, float32 result).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/srfe.h"

namespace srfe {

struct AugParams {
    const short* pcm;                  // [n_clips][clip_stride]
    float* out;                        // [n_clips][n_samples]
    const signed char* kind;           // per clip: 0 clip, 1 silence (zeros), 2 silence (noise slice); may be null
    signed char* op_out;               // may be null
    const short* bank;                 // background-noise samples, files back to back
    const long long* bank_off;         // [n_files + 1] (device)
    long long clip_stride, first_index;
    int n_clips, n_samples, n_files;
    unsigned key0, key1;
    float shift_lo, shift_hi, noise_lo, noise_hi, snr_lo, snr_hi, pitch_lo, pitch_hi, speed_lo, speed_hi;
    int shift_range;
    float noise_upper;
    double snr_div[4];                 // 10^(snr/10) for -5, 0, 5, 10 dB (host, double)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float aug_u01(unsigned r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }   // exact
__device__ __forceinline__ int aug_scaled(unsigned r, unsigned span) { return (int)(((unsigned long long)r * span) >> 32); }
__device__ __forceinline__ float aug_to_i16(double v) { return (float)(short)(int)v; }        // trunc toward zero, low 16 bits

__global__ void __launch_bounds__(256) srfe_augment_kernel(const AugParams p) {
    const int i = blockIdx.x, tid = threadIdx.x, N = p.n_samples;
    const long long g = p.first_index + i;
    const short* s = p.pcm + (long long)i * p.clip_stride;
    float* o = p.out + (long long)i * N;
    const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), 0u, 0u), p.key0, p.key1);
    const int kind = p.kind ? p.kind[i] : 0;
    // the noise slice the draws select (used by three of the ops)
    const short* nz = nullptr;
    if (p.n_files > 0) {
        const int f = aug_scaled(r.y, (unsigned)p.n_files);
        const long long b0 = p.bank_off[f], len = p.bank_off[f + 1] - b0;
        if (len >= N) nz = p.bank + b0 + aug_scaled(r.z, (unsigned)(len - N + 1));      // (a file shorter than a clip violates the
    }                                                                                   //  contract: its ops degrade to a copy)
    const float u = aug_u01(r.x);
    int op = SRFE_AUG_NONE;
    if (nz == nullptr && (kind == 2 || (u >= p.noise_lo && u < p.noise_hi) || (u >= p.snr_lo && u < p.snr_hi))) {
        for (int j = tid; j < N; j += blockDim.x) o[j] = kind == 2 ? 0.f : (float)s[j];
        op = kind == 2 ? SRFE_AUG_SILENCE_ZERO : SRFE_AUG_NONE;
    } else if (kind == 1) {
        op = SRFE_AUG_SILENCE_ZERO;
        for (int j = tid; j < N; j += blockDim.x) o[j] = 0.f;
    } else if (kind == 2) {
        op = SRFE_AUG_SILENCE_NOISE;
        const double sc = (double)aug_u01(r.w);
        for (int j = tid; j < N; j += blockDim.x) o[j] = (float)__dmul_rn((double)nz[j], sc);
    } else if (u >= p.shift_lo && u < p.shift_hi) {
        op = SRFE_AUG_SHIFT;
        const int shift = -p.shift_range + aug_scaled(r.y, 2u * (unsigned)p.shift_range + 1u);
        const int nf = shift >= 0 ? shift : -shift;
        for (int j = tid; j < N; j += blockDim.x) {
            // shift >= 0: [s[shift:], fill]; shift < 0: [fill, s[:shift]]
            const int fj = shift >= 0 ? j - (N - nf) : j;                 // index into the fill, if inside it
            float v;
            if (fj >= 0 && fj < nf) {
                const uint4 w = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), 2u + (unsigned)(fj >> 2), 0u), p.key0, p.key1);
                const unsigned ww = (fj & 3) == 0 ? w.x : (fj & 3) == 1 ? w.y : (fj & 3) == 2 ? w.z : w.w;
                v = (float)((int)(ww & 63u) - 32);
            } else {
                v = (float)s[shift >= 0 ? j + shift : j - nf];
            }
            o[j] = v;
        }
    } else if (u >= p.noise_lo && u < p.noise_hi) {
        op = SRFE_AUG_NOISE_UNIFORM;
        const double f = __dmul_rn((double)aug_u01(r.w), (double)p.noise_upper);
        for (int j = tid; j < N; j += blockDim.x) o[j] = aug_to_i16(__dadd_rn((double)s[j], __dmul_rn(f, (double)nz[j])));
    } else if (u >= p.snr_lo && u < p.snr_hi) {
        op = SRFE_AUG_NOISE_SNR;
        const int level = aug_scaled(philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), 1u, 0u), p.key0, p.key1).x, 5u);
        // exact integer sums of squares (|x| <= 2^15, N <= 2^24: < 2^54)
        long long ps = 0, pn = 0;
        for (int j = tid; j < N; j += blockDim.x) { const long long a = s[j], b = nz[j]; ps += a * a; pn += b * b; }
        __shared__ long long red[2][8];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, d); pn += __shfl_xor_sync(0xffffffffu, pn, d); }
        if ((tid & 31) == 0) { red[0][tid >> 5] = ps; red[1][tid >> 5] = pn; }
        __syncthreads();
        ps = 0; pn = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ps += red[0][w]; pn += red[1][w]; }
        if (level == 4 || pn == 0) {
            for (int j = tid; j < N; j += blockDim.x) o[j] = (float)s[j];
        } else {
            const double sp = __ddiv_rn(__ddiv_rn((double)ps, 1073741824.0), (double)N);
            const double np_ = __ddiv_rn(__ddiv_rn((double)pn, 1073741824.0), (double)N);
            const double f = __dsqrt_rn(__ddiv_rn(__ddiv_rn(sp, np_), p.snr_div[level]));
            for (int j = tid; j < N; j += blockDim.x) o[j] = aug_to_i16(__dadd_rn((double)s[j], __dmul_rn(f, (double)nz[j])));
        }
    } else {
        op = (u >= p.pitch_lo && u < p.pitch_hi) ? SRFE_AUG_HOST_PITCH : (u >= p.speed_lo && u < p.speed_hi) ? SRFE_AUG_HOST_SPEED : SRFE_AUG_NONE;
        for (int j = tid; j < N; j += blockDim.x) o[j] = (float)s[j];      // dataset.py:117
    }
    if (tid == 0 && p.op_out) p.op_out[i] = (signed char)op;
}

}  // namespace srfe
