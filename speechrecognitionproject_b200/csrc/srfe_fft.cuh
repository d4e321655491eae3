// srfe_fft.cuh -- register DFT codelets and the half-warp real-FFT pipeline,
// TWO FRAMES PER LANE in packed f32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2).
//
// A half-warp (16 lanes) transforms a PAIR of analysis frames (A, B) at once: every
// value is a P2 = (frame A, frame B) register pair and every butterfly is one packed
// instruction, so FP issue slots, shared-memory instructions and address arithmetic per
// frame are halved (the FP32 pipe itself still retires 128 lane-ops/clk/SM -- measured,
// scripts/ubench/fp32x2.cu -- which is what finally bounds this path).
//
//   z[m] = xw[2m] + i xw[2m+1]           M = N/2 complex points, V = M/16 per lane
//   pass 1   lane l: DFT-V over m = l + 16 j        (registers)
//   twiddle  W_M^{l k1}                              (scalar table in shared memory, broadcast operand)
//   exchange through a padded shared-memory tile     (real parts, then imaginary parts, through the same 8-byte slots)
//   pass 2   M=256: one DFT-16 per lane              (registers)
//            M=320: DFT-16 = two radix-4 passes with a second (skewed) exchange
//   untangle lane takes the pairs (k, M-k): one twiddle per pair, yields 4|X[k]|^2 and 4|X[M-k]|^2 (the 1/4 is
//            folded downstream); on the device the partner Z[M-k] comes by warp shuffle (srfe_kernels.cuh), the
//            shared-memory form below serves tests/emu
//
// Everything is __host__ __device__ and free of warp intrinsics so tests/emu runs the
// identical index arithmetic lane by lane on the CPU (phases are separated by warp
// syncs on the device, by loops on the host).
#pragma once

#include <vector_types.h>

#if defined(__CUDACC__)
#define SRFE_HD __host__ __device__ __forceinline__
#else
#define SRFE_HD inline
#endif

namespace srfe {

// ---- packed pair (frame A, frame B) ------------------------------------------------
struct __align__(8) P2 { float lo, hi; };

#if defined(__CUDA_ARCH__)
#define SRFE_U64(p) (*reinterpret_cast<unsigned long long*>(&(p)))
#define SRFE_CU64(p) (*reinterpret_cast<const unsigned long long*>(&(p)))
SRFE_HD P2 padd(const P2& a, const P2& b) { P2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(SRFE_U64(d)) : "l"(SRFE_CU64(a)), "l"(SRFE_CU64(b))); return d; }
SRFE_HD P2 psub(const P2& a, const P2& b) { P2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(SRFE_U64(d)) : "l"(SRFE_CU64(a)), "l"(SRFE_CU64(b))); return d; }
SRFE_HD P2 pmul(const P2& a, const P2& b) { P2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(SRFE_U64(d)) : "l"(SRFE_CU64(a)), "l"(SRFE_CU64(b))); return d; }
SRFE_HD P2 pfma(const P2& a, const P2& b, const P2& c) { P2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(SRFE_U64(d)) : "l"(SRFE_CU64(a)), "l"(SRFE_CU64(b)), "l"(SRFE_CU64(c))); return d; }
#else
SRFE_HD P2 padd(const P2& a, const P2& b) { P2 d; d.lo = a.lo + b.lo; d.hi = a.hi + b.hi; return d; }
SRFE_HD P2 psub(const P2& a, const P2& b) { P2 d; d.lo = a.lo - b.lo; d.hi = a.hi - b.hi; return d; }
SRFE_HD P2 pmul(const P2& a, const P2& b) { P2 d; d.lo = a.lo * b.lo; d.hi = a.hi * b.hi; return d; }
SRFE_HD P2 pfma(const P2& a, const P2& b, const P2& c) { P2 d; d.lo = a.lo * b.lo + c.lo; d.hi = a.hi * b.hi + c.hi; return d; }
#endif
SRFE_HD P2 bc(float s) { P2 d; d.lo = s; d.hi = s; return d; }                 // scalar -> both frames (FFMA2 .F32 operand)
SRFE_HD P2 mkp(float a, float b) { P2 d; d.lo = a; d.hi = b; return d; }

// ---- packed complex: 16 bytes, the unit of every shared-memory exchange ----------------
struct __align__(16) C2 { P2 re, im; };
typedef float2 cpx;                                                             // scalar complex (twiddle tables)

SRFE_HD C2 cadd(const C2& a, const C2& b) { C2 d; d.re = padd(a.re, b.re); d.im = padd(a.im, b.im); return d; }
SRFE_HD C2 csub(const C2& a, const C2& b) { C2 d; d.re = psub(a.re, b.re); d.im = psub(a.im, b.im); return d; }
// a * (wr + i wi), scalar twiddle shared by both frames: 2 FMUL2 + 2 FFMA2
SRFE_HD C2 cmuls(const C2& a, float wr, float wi) {
    const P2 WR = bc(wr), WI = bc(wi), NWI = bc(-wi);
    C2 d;
    d.re = pfma(a.im, NWI, pmul(a.re, WR));
    d.im = pfma(a.re, WI, pmul(a.im, WR));
    return d;
}
#define SRFE_SQRT1_2 0.70710678118654752440f
SRFE_HD C2 mul_w8_1(const C2& a) { C2 d; const P2 c = bc(SRFE_SQRT1_2); d.re = pmul(padd(a.re, a.im), c); d.im = pmul(psub(a.im, a.re), c); return d; }   // a * W8^1
SRFE_HD C2 mul_w8_3(const C2& a) { C2 d; d.re = pmul(psub(a.im, a.re), bc(SRFE_SQRT1_2)); d.im = pmul(padd(a.re, a.im), bc(-SRFE_SQRT1_2)); return d; }   // a * W8^3
SRFE_HD C2 mul_mi(const C2& a) { C2 d; d.re = a.im; d.im = psub(bc(0.f), a.re); return d; }                                                                // a * (-i)

// ---- forward DFT codelets (e^{-2 pi i nk/N}), in place, natural order out --------------
SRFE_HD void dft4(C2& a0, C2& a1, C2& a2, C2& a3) {
    const C2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
    C2 t3;                                               // (a1 - a3) * (-i), no explicit negation
    t3.re = psub(a1.im, a3.im);
    t3.im = psub(a3.re, a1.re);
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

SRFE_HD void dft5(C2& a0, C2& a1, C2& a2, C2& a3, C2& a4) {
    const P2 c1 = bc(0.30901699437494742410f), c2 = bc(-0.80901699437494742410f);
    const P2 s1 = bc(0.95105651629515357212f), s2 = bc(0.58778525229247312917f), ns1 = bc(-0.95105651629515357212f);
    const C2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
    C2 m1, m2, n1, n2;
    m1.re = pfma(c2, t2.re, pfma(c1, t1.re, a0.re)); m1.im = pfma(c2, t2.im, pfma(c1, t1.im, a0.im));
    m2.re = pfma(c1, t2.re, pfma(c2, t1.re, a0.re)); m2.im = pfma(c1, t2.im, pfma(c2, t1.im, a0.im));
    n1.re = pfma(s2, t4.re, pmul(s1, t3.re));        n1.im = pfma(s2, t4.im, pmul(s1, t3.im));
    n2.re = pfma(ns1, t4.re, pmul(s2, t3.re));       n2.im = pfma(ns1, t4.im, pmul(s2, t3.im));
    a0.re = padd(a0.re, padd(t1.re, t2.re)); a0.im = padd(a0.im, padd(t1.im, t2.im));
    a1.re = padd(m1.re, n1.im); a1.im = psub(m1.im, n1.re);        // m1 - i n1
    a4.re = psub(m1.re, n1.im); a4.im = padd(m1.im, n1.re);        // m1 + i n1
    a2.re = padd(m2.re, n2.im); a2.im = psub(m2.im, n2.re);
    a3.re = psub(m2.re, n2.im); a3.im = padd(m2.im, n2.re);
}

// 16-point: n = n1 + 4 n2, k = ka + 4 kb; constant twiddles W16^{n1 ka}
SRFE_HD void dft16(C2* v) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    // v[n1 + 4 ka] holds T[n1][ka]
    v[1 + 4] = cmuls(v[1 + 4], c1, -s1);  v[1 + 8] = mul_w8_1(v[1 + 8]);   v[1 + 12] = cmuls(v[1 + 12], s1, -c1);
    v[2 + 4] = mul_w8_1(v[2 + 4]);        v[2 + 8] = mul_mi(v[2 + 8]);     v[2 + 12] = mul_w8_3(v[2 + 12]);
    v[3 + 4] = cmuls(v[3 + 4], s1, -c1);  v[3 + 8] = mul_w8_3(v[3 + 8]);   v[3 + 12] = cmuls(v[3 + 12], -c1, s1);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) dft4(v[4 * ka], v[4 * ka + 1], v[4 * ka + 2], v[4 * ka + 3]);
    // v[kb + 4 ka] holds X[ka + 4 kb]  -> transpose to natural order (register renaming)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { const C2 t = v[b + 4 * a]; v[b + 4 * a] = v[a + 4 * b]; v[a + 4 * b] = t; }
}

// 20-point, Good-Thomas 4 x 5 (no internal twiddles), in place:
//   n = (5 n1 + 4 n2) mod 20 ; X[(5 k1 + 16 k2) mod 20] ends up at position (5 k1 + 4 k2) mod 20
SRFE_HD void dft20(C2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2)
        dft4(v[(4 * n2) % 20], v[(5 + 4 * n2) % 20], v[(10 + 4 * n2) % 20], v[(15 + 4 * n2) % 20]);       // over n1 -> k1
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
        dft5(v[(5 * k1) % 20], v[(5 * k1 + 4) % 20], v[(5 * k1 + 8) % 20], v[(5 * k1 + 12) % 20], v[(5 * k1 + 16) % 20]);
    C2 u[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) u[i] = v[i];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) v[(5 * k1 + 16 * k2) % 20] = u[(5 * k1 + 4 * k2) % 20];
}

// ---- geometry of the half-warp FFT ---------------------------------------------------
// The exchange tile holds ONE 8-byte plane (real parts, then imaginary parts, of both frames) at a time: the two
// planes go through the same SCRATCH_P2 slots one after the other, so a half-warp needs 2.2 KB (N = 512) / 3.2 KB
// (N = 640) of shared memory instead of twice that -- which is what lets two MFCC CTAs (each with its per-clip dB
// tile) share an SM -- and every access is an 8-byte one (LDS.128 delivers half the bandwidth of LDS.64 on B200,
// scripts/ubench/lds_issue.cu).
template <int NFFT> struct FftGeom;
template <> struct FftGeom<512> {
    static constexpr int N = 512, M = 256, L = 16, V = 16;
    static constexpr int XS = 17;                 // exchange row stride (points): odd -> the 16 lanes hit 16 bank pairs
    static constexpr int SCRATCH_P2 = 16 * 17;    // >= M + 1 + mel padding (packed power buffer)
};
template <> struct FftGeom<640> {
    static constexpr int N = 640, M = 320, L = 16, V = 20;
    static constexpr int XS = 20;                 // = 4 (mod 16): both radix-4 gathers conflict-free
    static constexpr int SCRATCH_P2 = 20 * 20;
};

// Twiddle tables a CTA keeps in shared memory (built on the host in double):
//   tw1[k1*16 + l] = W_M^{l k1}         (M entries)
//   twu[k]         = W_N^{k}, k<=M/2    (M/2+1 entries)
//   tw16[a*4 + c]  = W_16^{a c}         (16 entries, M=320 only)
struct FftTables {
    const cpx* tw1;
    const cpx* twu;
    const cpx* tw16;
};

template <int PLANE> SRFE_HD const P2& plane_of(const C2& c) { return PLANE == 0 ? c.re : c.im; }
template <int PLANE> SRFE_HD P2& plane_of(C2& c) { return PLANE == 0 ? c.re : c.im; }

// ---- pass 1: DFT-V over m = l + 16 j and the twiddles W_M^{l k1}, all in registers ---------
template <int NFFT>
SRFE_HD void fft_pass1(C2* v, int l, const FftTables& T) {
    typedef FftGeom<NFFT> G;
    if (G::V == 16) dft16(v); else dft20(v);
#pragma unroll
    for (int k1 = 1; k1 < G::V; ++k1) {
        const cpx w = T.tw1[k1 * 16 + l];
        v[k1] = cmuls(v[k1], w.x, w.y);
    }
}
// ---- first exchange: row k1, column l.  One plane per call; callers separate put / get with a warp sync ----
template <int NFFT, int PLANE>
SRFE_HD void xs_put(const C2* v, int l, P2* buf) {
    typedef FftGeom<NFFT> G;
#pragma unroll
    for (int k1 = 0; k1 < G::V; ++k1) buf[k1 * G::XS + l] = plane_of<PLANE>(v[k1]);
}
// N = 512: lane l gathers row l (k1 = l), 16 columns -> the inputs of its DFT-16 (output v[k2] = Z[l + 16 k2])
template <int PLANE>
SRFE_HD void xs_get_512(int l, const P2* buf, C2* w) {
    typedef FftGeom<512> G;
#pragma unroll
    for (int i = 0; i < 16; ++i) plane_of<PLANE>(w[i]) = buf[l * G::XS + i];
}

// ---- N = 640: DFT-16 over l = l1 + 4 l2 as two radix-4 passes ----------------------------------
// lane = a + 4 b.  pass 2: l1 = a, k1 = b + 4 i (i<5), DFT-4 over l2 -> k2a, then the W_16^{a k2a} twiddles.
template <int PLANE>
SRFE_HD void xs_get2_640(int lane, const P2* buf, C2* w) {
    typedef FftGeom<640> G;
    const int a = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int l2 = 0; l2 < 4; ++l2) plane_of<PLANE>(w[4 * i + l2]) = buf[(b + 4 * i) * G::XS + a + 4 * l2];
}
SRFE_HD void fft_pass2_640(int lane, C2* v, const FftTables& T) {
    const int a = lane & 3;
#pragma unroll
    for (int i = 0; i < 5; ++i) dft4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);      // index = k2a
    const cpx t1 = T.tw16[a * 4 + 1], t2 = T.tw16[a * 4 + 2], t3 = T.tw16[a * 4 + 3];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        v[4 * i + 1] = cmuls(v[4 * i + 1], t1.x, t1.y);
        v[4 * i + 2] = cmuls(v[4 * i + 2], t2.x, t2.y);
        v[4 * i + 3] = cmuls(v[4 * i + 3], t3.x, t3.y);
    }
}
// second exchange: element (l1 = a, k2a = c) of row k1 lives at column 4 c + ((a + c) & 3)
// (skewed so that the 16 lanes of both the scatter and the gather hit 16 distinct bank pairs)
template <int PLANE>
SRFE_HD void xs_put3_640(int lane, const C2* v, P2* buf) {
    typedef FftGeom<640> G;
    const int a = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) buf[(b + 4 * i) * G::XS + 4 * c + ((a + c) & 3)] = plane_of<PLANE>(v[4 * i + c]);
}
// pass 3: lane = c + 4 b handles k2a = c, k1 = b + 4 i; DFT-4 over l1 -> k2b;
// v[4 i + k2b] = Z[k1 + 20 (c + 4 k2b)]
template <int PLANE>
SRFE_HD void xs_get3_640(int lane, const P2* buf, C2* w) {
    typedef FftGeom<640> G;
    const int c = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int l1 = 0; l1 < 4; ++l1) plane_of<PLANE>(w[4 * i + l1]) = buf[(b + 4 * i) * G::XS + 4 * c + ((l1 + c) & 3)];
}
SRFE_HD void fft_pass3_640(C2* v) {
#pragma unroll
    for (int i = 0; i < 5; ++i) dft4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// ---- Z in natural order (tests/emu only: the kernels untangle straight from registers with shuffles) ----
SRFE_HD void fft_store_z_512(int l, const C2* v, C2* zbuf) {
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) zbuf[l + 16 * k2] = v[k2];
}
SRFE_HD void fft_store_z_640(int lane, const C2* v, C2* zbuf) {
    const int c = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) zbuf[(b + 4 * i) + 20 * (c + 4 * kb)] = v[4 * i + kb];
}

// ---- untangle: 4|X[k]|^2 and 4|X[M-k]|^2 from Z[k], Z[(M-k) mod M] ---------------------
// (k = 0 yields bins 0 and M; k = M/2 yields the same bin twice)
SRFE_HD void untangle_pair(const C2& zk, const C2& zm, float wr, float wi, P2& pk, P2& pm) {
    const P2 ax = padd(zk.re, zm.re), ay = psub(zk.im, zm.im);
    const P2 bx = padd(zk.im, zm.im), by = psub(zm.re, zk.re);      // (B.im, -B.re), B = zk - conj(zm)
    const P2 WR = bc(wr), WI = bc(wi), NWI = bc(-wi);
    const P2 cx = pfma(by, NWI, pmul(bx, WR)), cy = pfma(bx, WI, pmul(by, WR));
    const P2 px = padd(ax, cx), py = padd(ay, cy), qx = psub(ax, cx), qy = psub(ay, cy);
    pk = pfma(px, px, pmul(py, py));
    pm = pfma(qx, qx, pmul(qy, qy));
}

// lane l takes k = l + 16 r, r < M/32: pa[r] = 4|X[k]|^2, pb[r] = 4|X[M-k]|^2
// (lane 0, r = 0: pb[0] is the Nyquist bin M).  Returns 4|X[M/2]|^2 (valid on lane 0).
template <int NFFT>
SRFE_HD P2 fft_untangle(int l, const C2* zbuf, const FftTables& T, P2* pa, P2* pb) {
    typedef FftGeom<NFFT> G;
#pragma unroll
    for (int r = 0; r < G::M / 32; ++r) {
        const int k = l + 16 * r;
        const C2 zk = zbuf[k];
        const C2 zm = zbuf[(r == 0 && l == 0) ? 0 : G::M - k];
        const cpx w = T.twu[k];
        untangle_pair(zk, zm, w.x, w.y, pa[r], pb[r]);
    }
    P2 pmid = bc(0.f);
    if (l == 0) {
        const C2 z = zbuf[G::M / 2];
        pmid = pmul(bc(4.f), pfma(z.re, z.re, pmul(z.im, z.im)));
    }
    return pmid;
}

}  // namespace srfe
