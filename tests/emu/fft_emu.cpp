// fft_emu.cpp -- CPU lane-by-lane emulation of the packed half-warp FFT pipeline of
// speechrecognitionproject_b200/csrc/srfe_fft.cuh (TEST INFRASTRUCTURE).
//
// The device code separates its phases with warp syncs; here every phase is run for
// lanes 0..15 in a loop, against a plain array standing in for the shared-memory
// scratch.  Same templates, same index arithmetic, same tables (the f32x2 PTX ops have a
// plain C++ body on the host) -- so a wrong permutation or twiddle shows up in
// `pytest -m "not gpu"` instead of costing a GPU round trip.  Not part of the product;
// never used as a fallback.
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../speechrecognitionproject_b200/csrc/srfe_fft.cuh"
#include "../../speechrecognitionproject_b200/csrc/srfe_tables.h"

using namespace srfe;

template <int NFFT>
static void run(const float* xa, const float* xb, float* pa_out, float* pb_out) {
    typedef FftGeom<NFFT> G;
    std::vector<F2> tw1, twu, tw16;
    fft_twiddles(NFFT, tw1, twu, tw16);
    FftTables T;
    T.tw1 = reinterpret_cast<const cpx*>(tw1.data());
    T.twu = reinterpret_cast<const cpx*>(twu.data());
    T.tw16 = reinterpret_cast<const cpx*>(tw16.data());

    std::vector<P2> xs(G::SCRATCH_P2);            // the exchange tile: one plane at a time, as on the device
    std::vector<C2> zbuf(G::M);
    static C2 v[16][G::V], w[16][G::V];
    for (int l = 0; l < 16; ++l)
        for (int j = 0; j < G::V; ++j) {
            const int m = l + 16 * j;
            v[l][j].re = mkp(xa[2 * m], xb[2 * m]);
            v[l][j].im = mkp(xa[2 * m + 1], xb[2 * m + 1]);
        }
    for (int l = 0; l < 16; ++l) fft_pass1<NFFT>(v[l], l, T);
    if (NFFT == 512) {
        for (int l = 0; l < 16; ++l) xs_put<NFFT, 0>(v[l], l, xs.data());
        for (int l = 0; l < 16; ++l) xs_get_512<0>(l, xs.data(), w[l]);
        for (int l = 0; l < 16; ++l) xs_put<NFFT, 1>(v[l], l, xs.data());
        for (int l = 0; l < 16; ++l) xs_get_512<1>(l, xs.data(), w[l]);
        for (int l = 0; l < 16; ++l) { dft16(w[l]); fft_store_z_512(l, w[l], zbuf.data()); }
    } else {
        for (int l = 0; l < 16; ++l) xs_put<NFFT, 0>(v[l], l, xs.data());
        for (int l = 0; l < 16; ++l) xs_get2_640<0>(l, xs.data(), w[l]);
        for (int l = 0; l < 16; ++l) xs_put<NFFT, 1>(v[l], l, xs.data());
        for (int l = 0; l < 16; ++l) xs_get2_640<1>(l, xs.data(), w[l]);
        for (int l = 0; l < 16; ++l) fft_pass2_640(l, w[l], T);
        for (int l = 0; l < 16; ++l) xs_put3_640<0>(l, w[l], xs.data());
        for (int l = 0; l < 16; ++l) xs_get3_640<0>(l, xs.data(), v[l]);
        for (int l = 0; l < 16; ++l) xs_put3_640<1>(l, w[l], xs.data());
        for (int l = 0; l < 16; ++l) xs_get3_640<1>(l, xs.data(), v[l]);
        for (int l = 0; l < 16; ++l) { fft_pass3_640(v[l]); fft_store_z_640(l, v[l], zbuf.data()); }
    }
    for (int l = 0; l < 16; ++l) {
        P2 pa[G::M / 32], pb[G::M / 32];
        const P2 pmid = fft_untangle<NFFT>(l, zbuf.data(), T, pa, pb);
        for (int r = 0; r < G::M / 32; ++r) {
            const int k = l + 16 * r;
            pa_out[k] = 0.25f * pa[r].lo;           pb_out[k] = 0.25f * pa[r].hi;
            pa_out[G::M - k] = 0.25f * pb[r].lo;    pb_out[G::M - k] = 0.25f * pb[r].hi;
        }
        if (l == 0) { pa_out[G::M / 2] = 0.25f * pmid.lo; pb_out[G::M / 2] = 0.25f * pmid.hi; }
    }
}

// power spectra |rfft(x)|^2 of two frames processed as one packed pair
extern "C" int emu_power2(int nfft, const float* xa, const float* xb, float* pa, float* pb) {
    if (nfft == 512) { run<512>(xa, xb, pa, pb); return 0; }
    if (nfft == 640) { run<640>(xa, xb, pa, pb); return 0; }
    return -1;
}

// complex DFT codelets on two packed inputs (interleaved re/im arrays)
extern "C" int emu_dft2(int n, const float* a_ri, const float* b_ri, float* oa_ri, float* ob_ri) {
    C2 v[20];
    for (int i = 0; i < n; ++i) { v[i].re = mkp(a_ri[2 * i], b_ri[2 * i]); v[i].im = mkp(a_ri[2 * i + 1], b_ri[2 * i + 1]); }
    if (n == 4) dft4(v[0], v[1], v[2], v[3]);
    else if (n == 5) dft5(v[0], v[1], v[2], v[3], v[4]);
    else if (n == 16) dft16(v);
    else if (n == 20) dft20(v);
    else return -1;
    for (int i = 0; i < n; ++i) {
        oa_ri[2 * i] = v[i].re.lo; oa_ri[2 * i + 1] = v[i].im.lo;
        ob_ri[2 * i] = v[i].re.hi; ob_ri[2 * i + 1] = v[i].im.hi;
    }
    return 0;
}

// ---- mel / filterbank projection through the kernels' ELL tables (srfe_tables.cpp: to_sparse, to_ell) ----------
// Walks the tables exactly like the kernel's lane-per-filter loop (srfe_kernels.cuh: planar float2 weights, bank-skewed
// starts, uniform group run length) on one power spectrum.  `power` holds n_bins values; whatever the padded runs read
// beyond the filters' supports is filled with a large finite number, so a non-zero padding weight would show.
// Returns the worst number of lanes of a 16-filter group that start in the same shared-memory bank pair.
template <typename P>
static int ell_project(const P& p, bool mfcc, const float* power, double* out, int* max_reach) {
    std::vector<double> dense;
    SparseBank sb;
    EllBank ell;
    const int n_bins = p.n_fft / 2 + 1;
    int n_filt;
    if constexpr (std::is_same<P, srfe_mfcc_params>::value) { mfcc_filters(p, dense); n_filt = p.n_mels; }
    else { fbank_filters(p, dense); n_filt = p.nfilt; }
    (void)mfcc;
    to_sparse(dense, n_filt, n_bins, 1.0, sb);
    to_ell(sb, ell);
    std::vector<float> pbuf(std::max(ell.max_reach, n_bins) + 16, 1.0e30f);
    for (int k = 0; k < n_bins; ++k) pbuf[k] = power[k];
    int worst = 0;
    for (int g = 0; g < ell.groups; ++g) {
        const int off4 = ell.gmeta[2 * g], n4 = ell.gmeta[2 * g + 1];
        int hits[16] = {0};
        for (int l = 0; l < 16; ++l) {
            const int m = 16 * g + l;
            if (ell.start[m] < 0) return -1;
            worst = std::max(worst, ++hits[ell.start[m] & 15]);
            double acc = 0.0;
            for (int q4 = 0; q4 < n4; ++q4)
                for (int j = 0; j < 4; ++j) {
                    const float w = ell.w4[((((size_t)(off4 + q4) * 2 + (j >> 1)) * 16 + l) * 2) + (j & 1)];
                    if (w != 0.f) acc += (double)w * (double)pbuf[ell.start[m] + 4 * q4 + j];     // 0 * 1e30 stays 0 on the device too
                    else if (!(pbuf[ell.start[m] + 4 * q4 + j] == pbuf[ell.start[m] + 4 * q4 + j])) return -2;
                }
            if (m < n_filt) out[m] = acc;
        }
    }
    *max_reach = ell.max_reach;
    return worst;
}
extern "C" int emu_ell_project_mfcc(const srfe_mfcc_params* p, const float* power, double* out, int* max_reach) {
    return ell_project(*p, true, power, out, max_reach);
}
extern "C" int emu_ell_project_fbank(const srfe_fbank_params* p, const float* power, double* out, int* max_reach) {
    return ell_project(*p, false, power, out, max_reach);
}
