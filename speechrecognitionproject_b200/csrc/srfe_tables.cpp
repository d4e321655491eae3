// srfe_tables.cpp -- host-side table builders (double precision).
// Algorithms follow the reference call sites cited in include/srfe.h; parity with
// the oracle's numpy tables is asserted by tests/test_abi.py (no GPU needed).
#include "srfe_tables.h"

#include <algorithm>
#include <cmath>
#include <cstddef>

namespace srfe {

static const double kPi = 3.14159265358979323846264338327950288;

// numpy.linspace(start, stop, num) (endpoint=True), same operation order
static void linspace(double start, double stop, int num, std::vector<double>& y) {
    y.resize(num);
    if (num == 1) { y[0] = start; return; }
    const double step = (stop - start) / (double)(num - 1);
    for (int i = 0; i < num; ++i) y[i] = (double)i * step + start;
    y[num - 1] = stop;
}

int64_t spec_frames(const srfe_spec_params& p, int64_t n) {
    if (n < p.nperseg) return 0;
    return (n - p.noverlap) / (p.nperseg - p.noverlap);
}
int64_t fbank_frames(const srfe_fbank_params& p, int64_t n) {
    const int64_t d = n > p.frame_len ? n - p.frame_len : p.frame_len - n;
    return (d + p.frame_step - 1) / p.frame_step;           // ceil(|n - L| / step)
}
int64_t mfcc_frames(const srfe_mfcc_params& p, int64_t n) { return 1 + n / p.hop; }

static bool fft_ok(int n) { return n == 512 || n == 640; }

int validate(const srfe_spec_params& p, const char** why) {
    *why = "";
    if (!fft_ok(p.nperseg)) { *why = "spec: nperseg must be 512 or 640"; return SRFE_ERR_UNSUPPORTED; }
    if (p.noverlap < 0 || p.noverlap >= p.nperseg) { *why = "spec: need 0 <= noverlap < nperseg"; return SRFE_ERR_BAD_ARG; }
    if ((p.nperseg - p.noverlap) % 2) { *why = "spec: hop (nperseg - noverlap) must be even"; return SRFE_ERR_UNSUPPORTED; }
    if (p.sample_rate <= 0) { *why = "spec: sample_rate must be positive"; return SRFE_ERR_BAD_ARG; }
    if (p.layout != SRFE_LAYOUT_FT && p.layout != SRFE_LAYOUT_TF) { *why = "spec: bad layout"; return SRFE_ERR_BAD_ARG; }
    return SRFE_OK;
}
int validate(const srfe_fbank_params& p, const char** why) {
    *why = "";
    if (!fft_ok(p.n_fft)) { *why = "fbank: n_fft must be 512 or 640"; return SRFE_ERR_UNSUPPORTED; }
    if (p.frame_len < 2 || p.frame_len > p.n_fft) { *why = "fbank: need 2 <= frame_len <= n_fft"; return SRFE_ERR_BAD_ARG; }
    if (p.frame_step < 2 || p.frame_step % 2) { *why = "fbank: frame_step must be even and >= 2"; return SRFE_ERR_UNSUPPORTED; }
    if (p.nfilt < 1 || p.nfilt > 256) { *why = "fbank: need 1 <= nfilt <= 256"; return SRFE_ERR_UNSUPPORTED; }
    if (p.sample_rate <= 0) { *why = "fbank: sample_rate must be positive"; return SRFE_ERR_BAD_ARG; }
    if (p.vtlp_alpha != 0.f && !(p.vtlp_alpha >= 0.5f && p.vtlp_alpha <= 2.f)) { *why = "fbank: vtlp_alpha must be 0 (off) or in [0.5, 2]"; return SRFE_ERR_BAD_ARG; }
    return SRFE_OK;
}
int validate(const srfe_mfcc_params& p, const char** why) {
    *why = "";
    if (!fft_ok(p.n_fft)) { *why = "mfcc: n_fft must be 512 or 640"; return SRFE_ERR_UNSUPPORTED; }
    const int win = p.win_length > 0 ? p.win_length : p.n_fft;
    if (win > p.n_fft || win < 2) { *why = "mfcc: need 2 <= win_length <= n_fft"; return SRFE_ERR_BAD_ARG; }
    if (p.hop < 2 || p.hop % 2) { *why = "mfcc: hop must be even and >= 2"; return SRFE_ERR_UNSUPPORTED; }
    if (p.n_mels < 1 || p.n_mels > 256) { *why = "mfcc: need 1 <= n_mels <= 256"; return SRFE_ERR_UNSUPPORTED; }
    if (p.n_mfcc < 1 || p.n_mfcc > p.n_mels || p.n_mfcc > 64) { *why = "mfcc: need 1 <= n_mfcc <= min(n_mels, 64)"; return SRFE_ERR_UNSUPPORTED; }
    if (p.n_deltas < 0 || p.n_deltas > 2) { *why = "mfcc: n_deltas must be 0, 1 or 2"; return SRFE_ERR_UNSUPPORTED; }
    if (p.sample_rate <= 0) { *why = "mfcc: sample_rate must be positive"; return SRFE_ERR_BAD_ARG; }
    if (!(p.amin > 0.f)) { *why = "mfcc: amin must be positive"; return SRFE_ERR_BAD_ARG; }
    if (p.layout != SRFE_LAYOUT_FT && p.layout != SRFE_LAYOUT_TF) { *why = "mfcc: bad layout"; return SRFE_ERR_BAD_ARG; }
    return SRFE_OK;
}

// scipy.signal.windows.tukey(n, alpha, sym=False): build length n+1 symmetric, drop the last
void window_tukey_periodic(int n, double alpha, std::vector<double>& w) {
    const int m = n + 1;
    const int width = (int)std::floor(alpha * (m - 1) / 2.0);
    w.assign(n, 1.0);
    for (int i = 0; i <= width && i < n; ++i)
        w[i] = 0.5 * (1.0 + std::cos(kPi * (-1.0 + 2.0 * i / alpha / (m - 1))));
    for (int i = m - width - 1; i < n; ++i)
        w[i] = 0.5 * (1.0 + std::cos(kPi * (-2.0 / alpha + 1.0 + 2.0 * i / alpha / (m - 1))));
}

void spec_window(const srfe_spec_params& p, std::vector<double>& w) {
    window_tukey_periodic(p.nperseg, 0.25, w);
}

void fbank_window(const srfe_fbank_params& p, std::vector<double>& w) {
    w.assign(p.n_fft, 0.0);
    for (int i = 0; i < p.frame_len; ++i)                     // np.hamming(frame_len)
        w[i] = 0.54 - 0.46 * std::cos(2.0 * kPi * i / (double)(p.frame_len - 1));
}

void mfcc_window(const srfe_mfcc_params& p, std::vector<double>& w) {
    const int win = p.win_length > 0 ? p.win_length : p.n_fft;
    const int lpad = (p.n_fft - win) / 2;                     // librosa.util.pad_center
    w.assign(p.n_fft, 0.0);
    for (int i = 0; i < win; ++i) w[lpad + i] = 0.5 - 0.5 * std::cos(2.0 * kPi * i / (double)win);
}

void fbank_filters(const srfe_fbank_params& p, std::vector<double>& w) {
    const int nb = p.n_fft / 2 + 1;
    const double high = 2595.0 * std::log10(1.0 + (p.sample_rate / 2.0) / 700.0);
    std::vector<double> mel;
    linspace(0.0, high, p.nfilt + 2, mel);
    std::vector<double> bin(p.nfilt + 2);
    for (int i = 0; i < p.nfilt + 2; ++i) {
        double hz = 700.0 * (std::pow(10.0, mel[i] / 2595.0) - 1.0);
        if (p.vtlp_alpha != 0.f) {                           // legacy/model_8/dataset_top.py:251-252
            const double a = (double)p.vtlp_alpha, nyq = p.sample_rate / 2.0, m1 = std::min(a, 1.0);
            hz = hz < (4800.0 * m1 / a) ? hz * a : nyq - ((nyq - 4800.0 * m1) / (nyq - 4800.0 * (m1 / a))) * (nyq - hz);
        }
        bin[i] = std::floor((p.n_fft + 1) * hz / p.sample_rate);
    }
    w.assign((size_t)p.nfilt * nb, 0.0);
    for (int m = 1; m <= p.nfilt; ++m) {
        const int lo = (int)bin[m - 1], ce = (int)bin[m], hi = (int)bin[m + 1];
        for (int k = lo; k < ce && k < nb; ++k) w[(size_t)(m - 1) * nb + k] = (k - bin[m - 1]) / (bin[m] - bin[m - 1]);
        for (int k = ce; k < hi && k < nb; ++k) w[(size_t)(m - 1) * nb + k] = (bin[m + 1] - k) / (bin[m + 1] - bin[m]);
    }
}

static double hz_to_mel_slaney(double f) {
    const double logstep = std::log(6.4) / 27.0;
    return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) / logstep : f / (200.0 / 3.0);
}
static double mel_to_hz_slaney(double m) {
    const double logstep = std::log(6.4) / 27.0;
    return m >= 15.0 ? 1000.0 * std::exp(logstep * (m - 15.0)) : (200.0 / 3.0) * m;
}

void mfcc_filters(const srfe_mfcc_params& p, std::vector<double>& w) {
    const int nb = p.n_fft / 2 + 1;
    const double fmax = p.fmax > 0.f ? (double)p.fmax : p.sample_rate / 2.0;
    std::vector<double> freqs, mels;
    linspace(0.0, p.sample_rate / 2.0, nb, freqs);
    linspace(hz_to_mel_slaney((double)p.fmin), hz_to_mel_slaney(fmax), p.n_mels + 2, mels);
    std::vector<double> mf(p.n_mels + 2);
    for (int i = 0; i < p.n_mels + 2; ++i) mf[i] = mel_to_hz_slaney(mels[i]);
    w.assign((size_t)p.n_mels * nb, 0.0);
    for (int i = 0; i < p.n_mels; ++i) {
        const double d0 = mf[i + 1] - mf[i], d1 = mf[i + 2] - mf[i + 1];
        const double enorm = 2.0 / (mf[i + 2] - mf[i]);
        for (int k = 0; k < nb; ++k) {
            const double lower = -(mf[i] - freqs[k]) / d0;
            const double upper = (mf[i + 2] - freqs[k]) / d1;
            const double v = std::fmax(0.0, std::fmin(lower, upper));
            w[(size_t)i * nb + k] = v * enorm;
        }
    }
}

void mfcc_dct(const srfe_mfcc_params& p, std::vector<double>& d) {
    d.resize((size_t)p.n_mfcc * p.n_mels);
    for (int k = 0; k < p.n_mfcc; ++k)
        for (int n = 0; n < p.n_mels; ++n)
            d[(size_t)k * p.n_mels + n] =
                k == 0 ? 1.0 / std::sqrt((double)p.n_mels)
                       : std::cos(kPi * k * (2.0 * n + 1.0) / (2.0 * p.n_mels)) * std::sqrt(2.0 / p.n_mels);
}

void to_sparse(const std::vector<double>& dense, int n_filters, int n_bins, double scale, SparseBank& out) {
    out.start.assign(n_filters, 0);
    out.count.assign(n_filters, 0);
    out.offset.assign(n_filters, 0);
    out.weight.clear();
    out.max_count = 0;
    for (int m = 0; m < n_filters; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < n_bins; ++k)
            if (dense[(size_t)m * n_bins + k] != 0.0) { if (lo < 0) lo = k; hi = k; }
        out.offset[m] = (int32_t)out.weight.size();
        if (lo < 0) continue;                                 // structurally empty filter
        out.start[m] = lo;
        out.count[m] = hi - lo + 1;
        if (out.count[m] > out.max_count) out.max_count = out.count[m];
        for (int k = lo; k <= hi; ++k) out.weight.push_back((float)(dense[(size_t)m * n_bins + k] * scale));
    }
}

// Shared-memory bank skew of one 16-filter group.  Lane l reads the packed power values P[start_l + j] (8 bytes each,
// 16 bank pairs), so two lanes whose starts agree mod 16 serialise.  A filter may start up to `slack_l` bins early
// (leading zero weights) without lengthening the group's run: pick the shifts that minimise the worst multiplicity
// of a bank pair.  Capacitated bipartite matching lanes -> residues by augmenting paths, capacity 1, 2, ...
static void skew_group(const int* start, const int* slack, int* shift) {
    for (int cap = 1; cap <= 16; ++cap) {
        int owner[16][16], used[16] = {0}, res_of[16];
        for (int l = 0; l < 16; ++l) res_of[l] = -1;
        bool ok = true;
        for (int l0 = 0; l0 < 16 && ok; ++l0) {
            // BFS-free augmenting search (16 x 16: plain DFS with a visited mask per attempt)
            bool seen[16] = {false};
            struct Rec { static bool go(int l, const int* start, const int* slack, int cap, int owner[16][16], int* used,
                                        int* res_of, bool* seen) {
                for (int d = 0; d <= slack[l]; ++d) {
                    const int r = ((start[l] - d) % 16 + 16) % 16;
                    if (seen[r]) continue;
                    seen[r] = true;
                    if (used[r] < cap) { owner[r][used[r]++] = l; res_of[l] = r; return true; }
                    for (int i = 0; i < used[r]; ++i) {
                        const int o = owner[r][i];
                        if (go(o, start, slack, cap, owner, used, res_of, seen)) { owner[r][i] = l; res_of[l] = r; return true; }
                    }
                }
                return false;
            } };
            ok = Rec::go(l0, start, slack, cap, owner, used, res_of, seen);
        }
        if (!ok) continue;
        for (int l = 0; l < 16; ++l) {
            shift[l] = 0;
            for (int d = 0; d <= slack[l]; ++d)
                if (((start[l] - d) % 16 + 16) % 16 == res_of[l]) { shift[l] = d; break; }
        }
        return;
    }
    for (int l = 0; l < 16; ++l) shift[l] = 0;
}

void to_ell(const SparseBank& sb, EllBank& out) {
    const int n = (int)sb.start.size();
    out.groups = (n + 15) / 16;
    out.gmeta.assign(2 * out.groups, 0);
    out.start.assign(16 * out.groups, 0);
    out.w4.clear();
    out.max_reach = 0;
    int off4 = 0;
    for (int g = 0; g < out.groups; ++g) {
        int cmax = 0;
        for (int l = 0; l < 16; ++l) { const int m = 16 * g + l; if (m < n && sb.count[m] > cmax) cmax = sb.count[m]; }
        const int n4 = (cmax + 3) / 4;
        out.gmeta[2 * g] = off4;
        out.gmeta[2 * g + 1] = n4;
        out.w4.resize((size_t)(off4 + n4) * 16 * 4, 0.f);
        int st[16], slack[16], shift[16];
        for (int l = 0; l < 16; ++l) {
            const int m = 16 * g + l;
            const bool live = m < n && sb.count[m] > 0;
            // empty / padding lanes read (and ignore) 4 * n4 values from anywhere: any residue will do
            st[l] = live ? sb.start[m] : 15;
            slack[l] = live ? std::min(4 * n4 - sb.count[m], sb.start[m]) : 15;
        }
        skew_group(st, slack, shift);
        for (int l = 0; l < 16; ++l) {
            const int m = 16 * g + l;
            const int s0 = st[l] - shift[l];
            out.start[m] = s0;
            if (m >= n) continue;
            for (int q = 0; q < sb.count[m]; ++q) {
                const int qq = q + shift[l];
                // planar: [(run step)][plane = (qq & 3) / 2][lane] float2 -> two conflict-free LDS.64 per step
                out.w4[(((size_t)(off4 + qq / 4) * 2 + ((qq & 3) >> 1)) * 16 + l) * 2 + (qq & 1)] = sb.weight[sb.offset[m] + q];
            }
            if (s0 + 4 * n4 > out.max_reach) out.max_reach = s0 + 4 * n4;
        }
        off4 += n4;
    }
}

void fft_twiddles(int n_fft, std::vector<F2>& tw1, std::vector<F2>& twu, std::vector<F2>& tw16) {
    const int M = n_fft / 2, V = M / 16;
    tw1.resize(M);
    for (int k1 = 0; k1 < V; ++k1)
        for (int l = 0; l < 16; ++l) {
            const double a = -2.0 * kPi * (double)(l * k1) / (double)M;
            tw1[k1 * 16 + l] = F2{(float)std::cos(a), (float)std::sin(a)};
        }
    twu.resize(M / 2 + 1);
    for (int k = 0; k <= M / 2; ++k) {
        const double a = -2.0 * kPi * (double)k / (double)n_fft;
        twu[k] = F2{(float)std::cos(a), (float)std::sin(a)};
    }
    tw16.resize(16);
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c) {
            const double ang = -2.0 * kPi * (double)(a * c) / 16.0;
            tw16[a * 4 + c] = F2{(float)std::cos(ang), (float)std::sin(ang)};
        }
}

}  // namespace srfe
