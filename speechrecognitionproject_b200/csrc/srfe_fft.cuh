// srfe_fft.cuh -- register DFT codelets and the half-warp real-FFT pipeline.
//
// One analysis frame (real, N = 512 or 640 points after windowing / zero padding)
// is transformed by 16 lanes (half a warp; a warp carries two frames):
//
//   z[m] = xw[2m] + i xw[2m+1]           M = N/2 complex points, V = M/16 per lane
//   pass 1   lane l: DFT-V over m = l + 16 j        (registers)
//   twiddle  W_M^{l k1}                              (table in shared memory)
//   exchange through a padded shared-memory tile     (STS.64 / LDS.128)
//   pass 2   M=256: one DFT-16 per lane              (registers)
//            M=320: DFT-16 = two radix-4 passes with a second exchange
//   Z[k] -> shared, natural order
//   untangle lane takes the pairs (k, M-k), k = l + 16 r: one twiddle per pair,
//            yields |X[k]|^2 and |X[M-k]|^2 (x4, the 1/4 is folded downstream)
//
// Everything here is __host__ __device__ and free of warp intrinsics so that
// tests/emu can run the identical index arithmetic lane by lane on the CPU
// (phases are separated by warp syncs on the device, by loops on the host).
#pragma once

#include <vector_types.h>

#if defined(__CUDACC__)
#define SRFE_HD __host__ __device__ __forceinline__
#else
#define SRFE_HD inline
#endif

namespace srfe {

typedef float2 cpx;

SRFE_HD cpx mk(float x, float y) { cpx r; r.x = x; r.y = y; return r; }
SRFE_HD cpx cadd(cpx a, cpx b) { return mk(a.x + b.x, a.y + b.y); }
SRFE_HD cpx csub(cpx a, cpx b) { return mk(a.x - b.x, a.y - b.y); }
SRFE_HD cpx cmul(cpx a, cpx w) { return mk(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
SRFE_HD cpx mul_mi(cpx a) { return mk(a.y, -a.x); }                       // a * (-i)
#define SRFE_SQRT1_2 0.70710678118654752440f
SRFE_HD cpx mul_w8_1(cpx a) { return mk((a.x + a.y) * SRFE_SQRT1_2, (a.y - a.x) * SRFE_SQRT1_2); }   // a * W8^1
SRFE_HD cpx mul_w8_3(cpx a) { return mk((a.y - a.x) * SRFE_SQRT1_2, -(a.x + a.y) * SRFE_SQRT1_2); }  // a * W8^3

// ---- forward DFT codelets (e^{-2 pi i nk/N}), in place, natural order out ----
SRFE_HD void dft2(cpx& a, cpx& b) { cpx t = csub(a, b); a = cadd(a, b); b = t; }

SRFE_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
    cpx t0 = cadd(a0, a2), t1 = csub(a0, a2);
    cpx t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

SRFE_HD void dft5(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    cpx t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
    cpx m1 = mk(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
    cpx m2 = mk(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
    cpx n1 = mk(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    cpx n2 = mk(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    a0 = mk(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
    a1 = mk(m1.x + n1.y, m1.y - n1.x);        // m1 - i n1
    a4 = mk(m1.x - n1.y, m1.y + n1.x);        // m1 + i n1
    a2 = mk(m2.x + n2.y, m2.y - n2.x);
    a3 = mk(m2.x - n2.y, m2.y + n2.x);
}

// 16-point: n = n1 + 4 n2, k = ka + 4 kb; constant twiddles W16^{n1 ka}
SRFE_HD void dft16(cpx* v) {
    const cpx w1 = mk(0.92387953251128675613f, -0.38268343236508977173f);
    const cpx w3 = mk(0.38268343236508977173f, -0.92387953251128675613f);
    const cpx w9 = mk(-0.92387953251128675613f, 0.38268343236508977173f);
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    // v[n1 + 4 ka] holds T[n1][ka]
    v[1 + 4] = cmul(v[1 + 4], w1);      v[1 + 8] = mul_w8_1(v[1 + 8]);          v[1 + 12] = cmul(v[1 + 12], w3);
    v[2 + 4] = mul_w8_1(v[2 + 4]);      v[2 + 8] = mul_mi(v[2 + 8]);            v[2 + 12] = mul_w8_3(v[2 + 12]);
    v[3 + 4] = cmul(v[3 + 4], w3);      v[3 + 8] = mul_w8_3(v[3 + 8]);          v[3 + 12] = cmul(v[3 + 12], w9);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) dft4(v[4 * ka], v[4 * ka + 1], v[4 * ka + 2], v[4 * ka + 3]);
    // v[kb + 4 ka] holds X[ka + 4 kb]  -> transpose to natural order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { cpx t = v[b + 4 * a]; v[b + 4 * a] = v[a + 4 * b]; v[a + 4 * b] = t; }
}

// 20-point, Good-Thomas 4 x 5 (no internal twiddles):
//   n = (5 n1 + 4 n2) mod 20,  k = (5 k1 + 16 k2) mod 20
SRFE_HD void dft20(cpx* v) {
    cpx u[20];
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1)
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) u[n1 * 5 + n2] = v[(5 * n1 + 4 * n2) % 20];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) dft4(u[n2], u[5 + n2], u[10 + n2], u[15 + n2]);       // over n1 -> k1
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft5(u[5 * k1], u[5 * k1 + 1], u[5 * k1 + 2], u[5 * k1 + 3], u[5 * k1 + 4]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) v[(5 * k1 + 16 * k2) % 20] = u[k1 * 5 + k2];
}

// ---- geometry of the half-warp FFT -------------------------------------------
template <int NFFT> struct FftGeom;
template <> struct FftGeom<512> {
    static constexpr int N = 512, M = 256, L = 16, V = 16;
    static constexpr int XS = 18;                 // exchange row stride (cpx): LDS.128 conflict-free
    static constexpr int XROWS = 16;
    static constexpr int SCRATCH_CPX = 288;       // max(XROWS*XS, M)
};
template <> struct FftGeom<640> {
    static constexpr int N = 640, M = 320, L = 16, V = 20;
    static constexpr int XS = 18;
    static constexpr int XROWS = 20;
    static constexpr int SCRATCH_CPX = 360;
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

// ---- phase 1: DFT-V, twiddle, scatter into the exchange tile -------------------
template <int NFFT>
SRFE_HD void fft_phase1(cpx* v, int l, cpx* xbuf, const FftTables& T) {
    typedef FftGeom<NFFT> G;
    if (G::V == 16) dft16(v); else dft20(v);
#pragma unroll
    for (int k1 = 0; k1 < G::V; ++k1) {
        cpx a = (k1 == 0) ? v[0] : cmul(v[k1], T.tw1[k1 * 16 + l]);
        int col = (NFFT == 640) ? ((l & 3) * 4 + (l >> 2)) : l;     // M=320: l = l1 + 4 l2 -> 4 l1 + l2
        xbuf[k1 * G::XS + col] = a;
    }
}

// ---- phase 2 (N=512): one DFT-16 per lane, Z out in natural order ---------------
SRFE_HD void fft_phase2_512(int l, const cpx* xbuf, cpx* v) {
    typedef FftGeom<512> G;
    const float4* row = reinterpret_cast<const float4*>(xbuf + l * G::XS);     // 144 B rows: 16 B aligned
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 q = row[i]; v[2 * i] = mk(q.x, q.y); v[2 * i + 1] = mk(q.z, q.w); }
    dft16(v);                                   // v[k2] = Z[l + 16 k2]
}
SRFE_HD void fft_store_z_512(int l, const cpx* v, cpx* zbuf) {
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) zbuf[l + 16 * k2] = v[k2];
}

// ---- phases 2/3 (N=640): DFT-16 over l = l1 + 4 l2 as two radix-4 passes -------
// lane = a + 4 b.  pass 2: l1 = a, k1 = b + 4 i (i<5), DFT-4 over l2 -> k2a.
SRFE_HD void fft_phase2_640(int lane, const cpx* xbuf, cpx* v, const FftTables& T) {
    typedef FftGeom<640> G;
    const int a = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float4* row = reinterpret_cast<const float4*>(xbuf + (b + 4 * i) * G::XS + 4 * a);
        const float4 q0 = row[0], q1 = row[1];
        v[4 * i] = mk(q0.x, q0.y); v[4 * i + 1] = mk(q0.z, q0.w);
        v[4 * i + 2] = mk(q1.x, q1.y); v[4 * i + 3] = mk(q1.z, q1.w);
        dft4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);      // index = k2a
    }
    const cpx t1 = T.tw16[a * 4 + 1], t2 = T.tw16[a * 4 + 2], t3 = T.tw16[a * 4 + 3];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        v[4 * i + 1] = cmul(v[4 * i + 1], t1);
        v[4 * i + 2] = cmul(v[4 * i + 2], t2);
        v[4 * i + 3] = cmul(v[4 * i + 3], t3);
    }
}
// second exchange: row k1, column 4 k2a + a
SRFE_HD void fft_scatter2_640(int lane, const cpx* v, cpx* xbuf) {
    typedef FftGeom<640> G;
    const int a = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) xbuf[(b + 4 * i) * G::XS + 4 * c + a] = v[4 * i + c];
}
// pass 3: lane = c + 4 b handles k2a = c, k1 = b + 4 i; DFT-4 over l1 -> k2b;
// v[4 i + k2b] = Z[k1 + 20 (c + 4 k2b)]
SRFE_HD void fft_phase3_640(int lane, const cpx* xbuf, cpx* v) {
    typedef FftGeom<640> G;
    const int c = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float4* row = reinterpret_cast<const float4*>(xbuf + (b + 4 * i) * G::XS + 4 * c);
        const float4 q0 = row[0], q1 = row[1];
        v[4 * i] = mk(q0.x, q0.y); v[4 * i + 1] = mk(q0.z, q0.w);
        v[4 * i + 2] = mk(q1.x, q1.y); v[4 * i + 3] = mk(q1.z, q1.w);
        dft4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}
SRFE_HD void fft_store_z_640(int lane, const cpx* v, cpx* zbuf) {
    const int c = lane & 3, b = lane >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) zbuf[(b + 4 * i) + 20 * (c + 4 * kb)] = v[4 * i + kb];
}

// ---- untangle: 4|X[k]|^2 and 4|X[M-k]|^2 from Z[k], Z[(M-k) mod M] -------------
// (k = 0 yields bins 0 and M; k = M/2 yields the same bin twice)
SRFE_HD void untangle_pair(cpx zk, cpx zm, cpx w, float& pk, float& pm) {
    const float ax = zk.x + zm.x, ay = zk.y - zm.y;
    const float bx = zk.y + zm.y, by = zm.x - zk.x;      // (B.y, -B.x), B = zk - conj(zm)
    const float cx = bx * w.x - by * w.y, cy = bx * w.y + by * w.x;
    const float px = ax + cx, py = ay + cy, qx = ax - cx, qy = ay - cy;
    pk = px * px + py * py;
    pm = qx * qx + qy * qy;
}

// lane l takes k = l + 16 r, r < M/32: pa[r] = 4|X[k]|^2, pb[r] = 4|X[M-k]|^2
// (lane 0, r = 0: pb[0] is the Nyquist bin M).  Returns 4|X[M/2]|^2 (valid on lane 0).
template <int NFFT>
SRFE_HD float fft_untangle(int l, const cpx* zbuf, const FftTables& T, float* pa, float* pb) {
    typedef FftGeom<NFFT> G;
#pragma unroll
    for (int r = 0; r < G::M / 32; ++r) {
        const int k = l + 16 * r;
        const cpx zk = zbuf[k];
        const cpx zm = zbuf[(G::M - k) & (k == 0 ? 0 : 0x7fffffff)];
        untangle_pair(zk, zm, T.twu[k], pa[r], pb[r]);
    }
    float pmid = 0.f;
    if (l == 0) {
        const cpx z = zbuf[G::M / 2];
        pmid = 4.f * (z.x * z.x + z.y * z.y);
    }
    return pmid;
}

}  // namespace srfe
