// srfe_abi.cu -- the extern "C" boundary declared in include/srfe.h: parameter
// validation, the per-(device, parameter set) table cache, kernel launches and
// the host-buffer entry points (H2D -> kernel -> D2H, chunked over two streams).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

#include "../../include/srfe.h"
#include "srfe_kernels.cuh"
#include "srfe_mfcc_tc.cuh"
#include "srfe_fbank_tc.cuh"
#include "srfe_augment.cuh"
#include "srfe_tables.h"

namespace srfe {

static thread_local std::string g_err = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(SRFE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define SRFE_CUDA(call)                                              \
    do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)

// ------------------------------------------------------------------------------
// table cache
// ------------------------------------------------------------------------------
struct Entry {
    KParams kp;                 // everything except pcm/out/sizes
    int n_fft = 0;
    int family = 0;
    int blob_smem = 0;
    void* d_blob = nullptr;
    void* d_dct_kf = nullptr;
    struct DctVar { int cb, nbe, nbo, off, bytes; };
    std::vector<DctVar> dct_vars;   // folded DCT table, one copy per coefficient-block size (tail of the device blob)
    int blob_common = 0;        // bytes every CTA copies to shared memory; one DCT variant follows it there
    int mel_ng = 0;             // ELL shape: groups and 2-bit run-length code (0 groups = not encodable)
    unsigned mel_code = 0;
    int n_bins = 0;
    int tc_b_off = 0, tc_b_bytes = 0, tc_ne = 0, tc_no = 0;   // tcgen05 MFCC path: DCT B operand in the blob (0 bytes = not eligible)
    int ft_w_off = 0, ft_w_bytes = 0, ft_mid_slot = 0;         // tcgen05 FBANK path: filter weights as the MMA's A operand (0 bytes = not eligible)
};

struct Key {
    int device, family;
    std::string bytes;
    bool operator<(const Key& o) const {
        if (device != o.device) return device < o.device;
        if (family != o.family) return family < o.family;
        return bytes < o.bytes;
    }
};

static std::shared_mutex g_mu;                      // readers (every launch) share; table / configuration builds are exclusive
static std::map<Key, Entry*> g_cache;

static int align16(int x) { return (x + 15) & ~15; }

struct BlobBuilder {
    std::vector<unsigned char> data;
    int add(const void* src, size_t bytes) {
        const int off = (int)data.size();
        data.resize(align16(off + (int)bytes), 0);
        if (bytes) std::memcpy(data.data() + off, src, bytes);
        return off;
    }
};

static void window_range(const std::vector<double>& w, int& lo, int& hi) {
    int a = -1, b = -1;
    for (int i = 0; i < (int)w.size(); ++i) if (w[i] != 0.0) { if (a < 0) a = i; b = i; }
    if (a < 0) { lo = 0; hi = 0; return; }
    lo = a & ~1;
    hi = (b + 2) & ~1;                                   // exclusive, rounded up to even
}

static int upload(Entry* e, const BlobBuilder& bb, const std::vector<float>& dct_kf = std::vector<float>()) {
    if (!dct_kf.empty()) {
        SRFE_CUDA(cudaMalloc(&e->d_dct_kf, dct_kf.size() * sizeof(float)));
        SRFE_CUDA(cudaMemcpy(e->d_dct_kf, dct_kf.data(), dct_kf.size() * sizeof(float), cudaMemcpyHostToDevice));
        e->kp.dct_kf = (const float*)e->d_dct_kf;
    }
    SRFE_CUDA(cudaMalloc(&e->d_blob, bb.data.size()));
    SRFE_CUDA(cudaMemcpy(e->d_blob, bb.data.data(), bb.data.size(), cudaMemcpyHostToDevice));
    e->kp.blob = (const unsigned char*)e->d_blob;
    e->kp.blob_bytes = (int)bb.data.size();
    e->blob_smem = (int)bb.data.size();
    return SRFE_OK;
}

static void add_fft_tables(BlobBuilder& bb, KParams& kp, int n_fft, const std::vector<double>& win) {
    std::vector<float> wf(win.begin(), win.end());
    std::vector<F2> tw1, twu, tw16;
    fft_twiddles(n_fft, tw1, twu, tw16);
    kp.off_win = bb.add(wf.data(), wf.size() * 4);
    kp.off_tw1 = bb.add(tw1.data(), tw1.size() * 8);
    kp.off_twu = bb.add(twu.data(), twu.size() * 8);
    kp.off_tw16 = bb.add(tw16.data(), tw16.size() * 8);
    window_range(win, kp.w_lo, kp.w_hi);
}

static int add_bank(BlobBuilder& bb, KParams& kp, const SparseBank& sb, int n_fft, Entry* e) {
    EllBank ell;
    to_ell(sb, ell);
    e->mel_ng = ell.groups <= 16 ? ell.groups : 0;
    e->mel_code = 0;
    for (int g = 0; g < ell.groups && e->mel_ng; ++g) {
        const int n4 = ell.gmeta[2 * g + 1];
        if (n4 < 1 || n4 > 4) { e->mel_ng = 0; break; }
        e->mel_code |= (unsigned)(n4 - 1) << (2 * g);
    }
    // the packed power buffer aliases the FFT scratch: the padded runs must stay inside it
    const int cap = n_fft == 512 ? FftGeom<512>::SCRATCH_P2 : FftGeom<640>::SCRATCH_P2;          // P2 slots
    if (ell.max_reach > cap) return fail(SRFE_ERR_UNSUPPORTED, "filterbank too wide for the shared-memory power buffer");
    kp.off_gm = bb.add(ell.gmeta.data(), ell.gmeta.size() * 4);
    kp.off_fs = bb.add(ell.start.data(), ell.start.size() * 4);
    kp.off_fw4 = bb.add(ell.w4.data(), ell.w4.size() * 4);
    kp.n_filt = (int)sb.start.size();
    kp.n_fgroups = ell.groups;
    return SRFE_OK;
}

static int build_entry(const srfe_spec_params& p, Entry* e);
static int build_entry(const srfe_fbank_params& p, Entry* e);
static int build_entry(const srfe_mfcc_params& p, Entry* e);

template <typename P>
static int get_entry(int family, const P& p, Entry** out) {
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return fail(SRFE_ERR_NO_DEVICE, std::string("cudaGetDevice: ") + cudaGetErrorString(ce));
    Key key{dev, family, std::string((const char*)&p, sizeof(P))};
    {
        std::shared_lock<std::shared_mutex> rl(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *out = it->second; return SRFE_OK; }
    }
    std::unique_lock<std::shared_mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second; return SRFE_OK; }
    Entry* e = new Entry();
    std::memset(&e->kp, 0, sizeof(KParams));
    e->family = family;
    int rc = build_entry(p, e);
    if (rc != SRFE_OK) { delete e; return rc; }
    g_cache[key] = e;
    *out = e;
    return SRFE_OK;
}

static int build_entry(const srfe_spec_params& p, Entry* e) {
    std::vector<double> win;
    spec_window(p, win);
    double sw2 = 0.0;
    for (double w : win) sw2 += w * w;
    BlobBuilder bb;
    add_fft_tables(bb, e->kp, p.nperseg, win);
    e->n_fft = p.nperseg;
    e->n_bins = p.nperseg / 2 + 1;
    e->kp.hop = p.nperseg - p.noverlap;
    e->kp.start0 = 0;
    e->kp.scale = (float)(0.25 / ((double)p.sample_rate * sw2));   // density scaling; 1/4 from the untangle
    e->kp.log_eps = p.log_eps;
    e->kp.take_log = p.take_log;
    e->kp.layout = p.layout;
    return upload(e, bb);
}

static int build_entry(const srfe_fbank_params& p, Entry* e) {
    std::vector<double> win, dense;
    fbank_window(p, win);
    fbank_filters(p, dense);
    SparseBank sb;
    to_sparse(dense, p.nfilt, p.n_fft / 2 + 1, 0.25 / (double)p.n_fft, sb);   // |X|^2 / NFFT (model_fbanks_cnn.py:43)
    BlobBuilder bb;
    add_fft_tables(bb, e->kp, p.n_fft, win);
    { int rc = add_bank(bb, e->kp, sb, p.n_fft, e); if (rc != SRFE_OK) return rc; }
    e->n_fft = p.n_fft;
    e->n_bins = p.n_fft / 2 + 1;
    e->kp.hop = p.frame_step;
    e->kp.start0 = 0;
    e->kp.preemph = p.preemph;
    e->kp.layout = SRFE_LAYOUT_TF;
    const int classic_bytes = (int)bb.data.size();
    // tcgen05 path (srfe_fbank_tc.cuh): the same fp32 weights as a dense [128 filters x 256] A operand, K-permuted the way
    // the frame warps lay a frame's bins out (chunk c < 16: bins c + 16 e; 16 <= c < 32: bins 256 - (c - 16 + 16 e)), each
    // weight split into bf16 hi + bf16 mid (round to nearest), two K elements per 32-bit TMEM column (even element in the
    // low half): row = 128 hi words, then 128 mid words.  K = 256 = 16 MMA steps holds 256 of the 257 bins: bin 128 takes
    // the slot of an edge bin that no filter weighs (bin 0 whenever the bank starts at 0 Hz -- a triangle is zero at its
    // left foot -- else bin 256); a bank that weighs both edge bins stays on the classic kernel.
    if (p.n_fft == 512 && p.nfilt <= 128) {
        const int n_bins = p.n_fft / 2 + 1;
        auto weightless = [&](int bin) {
            for (int m = 0; m < p.nfilt; ++m) if (dense[(size_t)m * n_bins + bin] != 0.0) return false;
            return true;
        };
        const int mid_slot = weightless(0) ? 0 : weightless(256) ? 1 : -1;
        auto bf16_rn = [](float x) -> uint32_t {
            uint32_t u; std::memcpy(&u, &x, 4);
            return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;                 // finite inputs only
        };
        auto bf16_f = [](uint32_t h) -> float { const uint32_t u = h << 16; float x; std::memcpy(&x, &u, 4); return x; };
        std::vector<uint32_t> wt((size_t)128 * kFtKP, 0u);
        for (int m = 0; m < p.nfilt && mid_slot >= 0; ++m)
            for (int kk = 0; kk < kFtKP; ++kk) {
                const int c = kk / 8, el = kk % 8;
                int bin = c < 16 ? c + 16 * el : 256 - ((c - 16) + 16 * el);
                if (bin == (mid_slot == 0 ? 0 : 256)) bin = 128;
                const float w = (float)(dense[(size_t)m * n_bins + bin] * (0.25 / (double)p.n_fft));   // the classic kernel's weight
                const uint32_t hi = bf16_rn(w), mid = bf16_rn(w - bf16_f(hi));
                const size_t row = (size_t)m * kFtKP;
                wt[row + kk / 2] |= hi << (16 * (kk & 1));
                wt[row + kFtKP / 2 + kk / 2] |= mid << (16 * (kk & 1));
            }
        if (mid_slot >= 0) {
            while (bb.data.size() % 128) bb.data.push_back(0);
            e->ft_w_bytes = (int)wt.size() * 4;
            e->ft_w_off = bb.add(wt.data(), wt.size() * 4);
            e->ft_mid_slot = mid_slot;
        }
    }
    const int rc = upload(e, bb);
    e->blob_smem = classic_bytes;                                          // what the classic kernel copies to shared memory
    e->kp.blob_bytes = classic_bytes;
    return rc;
}

static int build_entry(const srfe_mfcc_params& p, Entry* e) {
    std::vector<double> win, dense, dct;
    mfcc_window(p, win);
    mfcc_filters(p, dense);
    mfcc_dct(p, dct);
    // The kernel keeps the mel energies as log2 values; 10 log10(x) = kDb * log2(x) is folded into the (linear) DCT here,
    // in double, and into the top_db threshold -- one multiply per filter and frame less in the kernel.
    const double kDb = 10.0 * std::log10(2.0);
    for (double& d : dct) d *= kDb;
    SparseBank sb;
    to_sparse(dense, p.n_mels, p.n_fft / 2 + 1, 0.25, sb);
    BlobBuilder bb;
    add_fft_tables(bb, e->kp, p.n_fft, win);
    { int rc = add_bank(bb, e->kp, sb, p.n_fft, e); if (rc != SRFE_OK) return rc; }
    e->n_fft = p.n_fft;
    e->n_bins = p.n_fft / 2 + 1;
    e->kp.hop = p.hop;
    e->kp.start0 = -(p.n_fft / 2);
    e->kp.n_mfcc = p.n_mfcc;
    e->kp.n_deltas = p.n_deltas;
    e->kp.top_db = p.top_db >= 0.f ? (float)((double)p.top_db / kDb) : p.top_db;
    e->kp.amin = p.amin;
    e->kp.layout = p.layout;
    e->kp.tile_stride = p.n_mels | 1;                                  // P2 units, odd: conflict-free pair-row reads
    double rs0 = 0.0;
    for (int f = 0; f < p.n_mels; ++f) rs0 += dct[f];
    e->kp.dct_row0_sum = (float)rs0;                                   // = sqrt(n_mels)
    // Folded DCT tables (n_mels % 4 == 0).  The DCT-II symmetry D[k][n-1-f] = (-1)^k D[k][f] lets even k work on
    // s[f] = x[f] + x[n-1-f] and odd k on d[f] = x[f] - x[n-1-f], f < n/2.  One kernel thread owns one frame pair and
    // one block of CB same-parity coefficients; a block's table is a float4 plane [f < n/2] (coefficients 0..3) plus,
    // for CB > 4, a second plane of 1 / 2 / 4 floats per f: one LDS.128 and at most one more load per f.  Which CB
    // fills the CTA best depends on the frame count and the launch shape, so every CB gets its own copy after the
    // common tables.
    e->kp.dct_fold = (p.n_mels >= 4 && p.n_mels % 4 == 0) ? 1 : 0;
    e->blob_common = (int)bb.data.size();
    if (e->kp.dct_fold) {
        const int half = p.n_mels / 2, ne = (p.n_mfcc + 1) / 2, no = p.n_mfcc / 2;
        for (int cb : {2, 3, 4, 5, 6, 8}) {
            const int rb = cb <= 4 ? 0 : cb == 5 ? 1 : cb == 6 ? 2 : 4;       // floats per entry in the second plane
            Entry::DctVar v{cb, (ne + cb - 1) / cb, (no + cb - 1) / cb, 0, 0};
            const size_t nent = (size_t)(v.nbe + v.nbo) * half;
            std::vector<float> tab(((nent * (4 + rb) + 3) / 4) * 4, 0.f);
            for (int k = 0; k < p.n_mfcc; ++k) {
                const int j = k >> 1, blk = (k & 1) * v.nbe + j / cb, jj = j % cb;
                for (int f = 0; f < half; ++f) {
                    const size_t ent = (size_t)blk * half + f;
                    const float d = (float)dct[(size_t)k * p.n_mels + f];
                    if (jj < 4) tab[ent * 4 + jj] = d; else tab[nent * 4 + ent * rb + (jj - 4)] = d;
                }
            }
            v.bytes = (int)tab.size() * 4;
            v.off = bb.add(tab.data(), tab.size() * 4);
            e->dct_vars.push_back(v);
        }
    }
    // tcgen05 path (srfe_mfcc_tc.cuh): the folded DCT-II rows as the MMA's B operand, [N x n/2] K-major in the no-swizzle
    // UMMA layout (8-row x 16-byte core matrices, 128 B apart along K, (n/2)/4 * 128 B apart along N), even-k rows and
    // odd-k rows separately, each as a TF32 "hi" copy (13 low mantissa bits cleared) and the fp32 remainder "lo".
    if (p.n_mels % 16 == 0 && p.n_mels <= 128 && p.n_mfcc <= 128) {
        const int half = p.n_mels / 2, ne = (p.n_mfcc + 1) / 2, no = p.n_mfcc / 2;
        e->tc_ne = std::max(16, (ne + 15) / 16 * 16);
        e->tc_no = std::max(16, (no + 15) / 16 * 16);
        std::vector<float> bt((size_t)2 * (e->tc_ne + e->tc_no) * half, 0.f);
        auto fill = [&](size_t base, int N, int par, int count) {
            float* hi = bt.data() + base;
            float* lo = hi + (size_t)N * half;
            for (int j = 0; j < count; ++j)
                for (int f = 0; f < half; ++f) {
                    const float d = (float)dct[(size_t)(2 * j + par) * p.n_mels + f];
                    uint32_t u; std::memcpy(&u, &d, 4); u &= 0xffffe000u;
                    float dh; std::memcpy(&dh, &u, 4);
                    const size_t off = ((size_t)(j / 8) * (half / 4) * 128 + (size_t)(f / 4) * 128 + (j % 8) * 16 + (f % 4) * 4) / 4;
                    hi[off] = dh;
                    lo[off] = d - dh;
                }
        };
        fill(0, e->tc_ne, 0, ne);
        fill((size_t)2 * e->tc_ne * half, e->tc_no, 1, no);
        while (bb.data.size() % 128) bb.data.push_back(0);
        e->tc_b_bytes = (int)bt.size() * 4;
        e->tc_b_off = bb.add(bt.data(), bt.size() * 4);
    }
    std::vector<float> dct_kf((size_t)p.n_mfcc * p.n_mels);
    for (size_t i = 0; i < dct_kf.size(); ++i) dct_kf[i] = (float)dct[i];
    return upload(e, bb, dct_kf);
}

// ------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
struct DevInfo { int sms = 0; int smem_optin = 0; int index = 0; };
static DevInfo g_dev[kMaxDevices];

static int dev_info(DevInfo** out) {
    int dev = 0;
    SRFE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(SRFE_ERR_BAD_ARG, "device index out of range");
    static std::atomic<int> ready[kMaxDevices];
    if (!ready[dev].load(std::memory_order_acquire)) {
        std::unique_lock<std::shared_mutex> lk(g_mu);
        g_dev[dev].index = dev;
        SRFE_CUDA(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
        SRFE_CUDA(cudaDeviceGetAttribute(&g_dev[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        ready[dev].store(1, std::memory_order_release);
    }
    *out = &g_dev[dev];
    return SRFE_OK;
}

template <int NFFT, int FAM, int JLO, int JHI, int NG, unsigned CODE, typename SAMP>
static int launch_k(const KParams& kp, int dev, int grid, int threads, int smem_bytes, cudaStream_t st) {
    auto kern = srfe_kernel<NFFT, FAM, JLO, JHI, NG, CODE, SAMP>;
    // opt-in shared-memory limit of this instantiation, raised on demand, per device; lock-free on the hot path
    // (two threads racing here both set a sufficient value: the attribute only ever grows)
    static std::atomic<int> attr_set[kMaxDevices];
    if (attr_set[dev].load(std::memory_order_acquire) < smem_bytes) {
        SRFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int cur = attr_set[dev].load(std::memory_order_relaxed);
        while (cur < smem_bytes && !attr_set[dev].compare_exchange_weak(cur, smem_bytes, std::memory_order_release)) {}
    }
    kern<<<grid, threads, smem_bytes, st>>>(kp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "srfe_kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SRFE_OK;
}

template <int NFFT, int JLO, int JHI, int NG, unsigned CODE, typename SAMP, int MB = 1>
static int launch_tc(const KParams& kp, int dev, int grid, int smem_bytes, cudaStream_t st) {
    auto kern = srfe_mfcc_tc_kernel<NFFT, JLO, JHI, NG, CODE, SAMP, MB>;
    static std::atomic<int> attr_set[kMaxDevices];
    if (attr_set[dev].load(std::memory_order_acquire) < smem_bytes) {
        SRFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int cur = attr_set[dev].load(std::memory_order_relaxed);
        while (cur < smem_bytes && !attr_set[dev].compare_exchange_weak(cur, smem_bytes, std::memory_order_release)) {}
    }
    kern<<<grid, kTcThreads, smem_bytes, st>>>(kp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "srfe_mfcc_tc_kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SRFE_OK;
}

// Shared-memory plan of the tcgen05 MFCC kernel: [tables][DCT B operand][FFT scratch of the frame warps]
// [ring of pair rows + dummy][frame means][coefficient tile (deltas only)][control block].  Returns false when the
// parameter set / clip length is not eligible (the classic kernel takes it).
static bool plan_tc(const Entry* e, KParams& kp, const DevInfo& di, int* smem_out) {
    if (e->family != FAM_MFCC || e->tc_b_bytes == 0 || kp.T < 3 || kp.T > 4 * (32 - 2 * kp.n_deltas)) return false;
    auto up = [](int x, int a) { return (x + a - 1) / a * a; };
    const int P = (kp.T + 1) / 2, nw = kTcThreads / 32 - kTcEpiWarps;   // FFT scratch for the frame warps only
    const int tmem_need = 2 * kp.n_filt + 2 * (e->tc_ne + e->tc_no);   // A operand (s / d, hi / lo) + two sets of accumulators
    if (tmem_need > 512) return false;
    int cols = 32;
    while (cols < tmem_need) cols *= 2;
    int off = up(e->blob_common, 128);
    kp.tc_off_b = off;
    off = up(off + e->tc_b_bytes, 128);
    kp.sm_scratch = off;
    off += 2 * nw * (e->n_fft == 512 ? FftGeom<512>::SCRATCH_P2 : FftGeom<640>::SCRATCH_P2) * 8;
    kp.sm_tile = off;
    const int ctile = 0;                                               // deltas run in registers (warp shuffles)
    const int row_bytes = kp.tile_stride * 8 + 8;                      // pair row + its two means
    const int left = di.smem_optin - off - ctile - 4 * TC_WORDS - 256;
    int ring = std::min(2 * P, left / row_bytes - 1);
    if (ring < P + 8) return false;
    kp.tc_ring = ring;
    kp.tc_early = ring >= 4 * nw + P ? 1 : 0;                          // room for two pairs per half-warp in flight plus a clip
    off += (ring + 1) * kp.tile_stride * 8;
    kp.tc_off_fmean = off;
    off = up(off + (ring + 1) * 8, 16);
    kp.sm_ctile = off;
    off = up(off + ctile, 16);
    kp.tc_off_ctrl = off;
    off += 4 * TC_WORDS;
    kp.blob_bytes = up(e->blob_common, 16);
    kp.tc_b_src = e->tc_b_off;
    kp.tc_b_bytes = e->tc_b_bytes;
    kp.tc_ne = e->tc_ne;
    kp.tc_no = e->tc_no;
    kp.tc_tmem_cols = cols;
    kp.tc_ring_magic = (unsigned)((0x100000000ULL + (unsigned long long)ring - 1) / (unsigned long long)ring);
    kp.tc_p_magic = (unsigned)((0x100000000ULL + (unsigned long long)P - 1) / (unsigned long long)P);
    *smem_out = up(off, 16);
    return *smem_out <= di.smem_optin;
}

template <int JLO, int JHI, typename SAMP>
static int launch_ft(const KParams& kp, int dev, int grid, int smem_bytes, cudaStream_t st) {
    auto kern = srfe_fbank_tc_kernel<JLO, JHI, SAMP>;
    static std::atomic<int> attr_set[kMaxDevices];
    if (attr_set[dev].load(std::memory_order_acquire) < smem_bytes) {
        SRFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int cur = attr_set[dev].load(std::memory_order_relaxed);
        while (cur < smem_bytes && !attr_set[dev].compare_exchange_weak(cur, smem_bytes, std::memory_order_release)) {}
    }
    kern<<<grid, kFtThreads, smem_bytes, st>>>(kp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "srfe_fbank_tc_kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SRFE_OK;
}

// Shared-memory plan of the tcgen05 FBANK kernel: [FFT tables][FFT scratch of the frame warps][three 48-frame tile buffers]
// [control block].  Returns false when the parameter set is not eligible (the classic kernel takes it).
static bool plan_ft(const Entry* e, KParams& kp, const DevInfo& di, int* smem_out) {
    if (e->family != FAM_FBANK || e->ft_w_bytes == 0 || e->n_fft != 512 || kp.T < 1) return false;
    if ((double)kp.n_clips * ((kp.T + 1) / 2) >= 268435456.0) return false;              // pair-stream indices: ints, magic divisions
    auto up = [](int x, int a) { return (x + a - 1) / a * a; };
    const int P = (kp.T + 1) / 2, nw = kFtThreads / 32 - kTcEpiWarps;
    kp.blob_bytes = up(kp.off_gm, 16);                                                   // FFT tables only (they precede the mel tables)
    int off = up(kp.blob_bytes, 128);
    kp.sm_scratch = off;
    off = up(off + 2 * nw * FftGeom<512>::SCRATCH_P2 * 8, 128);
    kp.tc_off_b = off;
    off += kFtBufs * kFtTileBytes;
    kp.tc_off_ctrl = off;
    off += 4 * FT_WORDS;
    kp.tc_b_src = e->ft_w_off;
    kp.tc_b_bytes = e->ft_w_bytes;
    kp.tc_ne = e->ft_mid_slot;                                                           // which edge bin's K slot carries bin 128
    kp.tc_p_magic = (unsigned)((0x100000000ULL + (unsigned long long)P - 1) / (unsigned long long)P);
    *smem_out = up(off, 16);
    return *smem_out <= di.smem_optin;
}

template <int NFFT, typename SAMP>
static int launch_staged(const KParams& kp, int dev, int grid, int threads, int smem_bytes, cudaStream_t st) {
    auto kern = srfe_spec_staged_kernel<NFFT, SAMP>;
    static std::atomic<int> attr_set[kMaxDevices];
    if (attr_set[dev].load(std::memory_order_acquire) < smem_bytes) {
        SRFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int cur = attr_set[dev].load(std::memory_order_relaxed);
        while (cur < smem_bytes && !attr_set[dev].compare_exchange_weak(cur, smem_bytes, std::memory_order_release)) {}
    }
    kern<<<grid, threads, smem_bytes, st>>>(kp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "srfe_spec_staged_kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SRFE_OK;
}

struct Config { int warps, ctas, cpc, smem, scratch, tile, ctile_off, blob, dct_var, dct_pq; };

// Launch-shape overrides (srfe_set_tuning; 0 = automatic).  A test / tuning hook: results never depend on them
// (tests/test_parity_gpu.py::test_results_do_not_depend_on_launch_configuration); no environment is read on the hot path.
enum Tuning { TUNE_WARPS = 0, TUNE_CTAS, TUNE_CPC, TUNE_DCT_CB, TUNE_DCT_PQ, TUNE_MFCC_TC, TUNE_STAGE, TUNE_FBANK_TC, TUNE_COUNT };
static const char* const kTuningNames[TUNE_COUNT] = {"warps", "ctas", "cpc", "dct_cb", "dct_pq", "mfcc_tc", "stage", "fbank_tc"};
static std::atomic<int> g_tune[TUNE_COUNT];
static int tune(int which) { return g_tune[which].load(std::memory_order_relaxed); }

// Shared-memory plan of a CTA with `warps` warps:  [tables][FFT scratch][MFCC pair-row dB tile + pair means]
// The MFCC epilogue reuses the FFT scratch for the coefficient tile; delta rows may spill into the (dead) dB tile.
static int smem_plan(const Entry* e, const KParams& kp, int warps, int budget, Config* c) {
    (void)budget;
    const int hw = 2 * warps;
    int blob = align16(e->blob_smem);
    c->dct_var = -1;
    c->dct_pq = 1;
    if (!e->dct_vars.empty()) {
        // DCT tile (CB coefficients x PQ frame pairs per thread): modelled cycles per clip = passes over the CTA's
        // threads x max(shared-memory wavefronts of all busy warps, FFMA2 issue on the busiest sub-partition), per f:
        //   LDS.64 of a pair 2 cycles, coefficient LDS.128 (one address) 2 cycles (+ ~1 for a 5th/6th), FFMA2 2 cycles
        //   (scripts/ubench/lds_issue.cu)
        const long long npairs = (kp.T + 1) / 2, nthr = 32LL * warps;
        double best = -1.0;
        for (size_t i = 0; i < e->dct_vars.size(); ++i) {
            const Entry::DctVar& v = e->dct_vars[i];
            for (int pq = 1; pq <= 2; ++pq) {
                const long long items = (long long)(v.nbe + v.nbo) * ((((npairs + pq - 1) / pq) + 15) / 16 * 16);   // half-warp padded
                const long long passes = (items + nthr - 1) / nthr;
                const double busy = std::min((double)items / passes, (double)nthr) / 32.0;      // warps per pass
                const double smem = busy * (2.0 * pq + (v.cb <= 4 ? 2.0 : v.cb <= 6 ? 3.0 : 4.0));
                const double fp = std::ceil(busy / 4.0) * 2.0 * v.cb * pq;
                const double cost = passes * (std::max(smem, fp) + 1.0);
                if (best < 0 || cost < best) { best = cost; c->dct_var = (int)i; c->dct_pq = pq; }
            }
        }
        if (tune(TUNE_DCT_CB) > 0)                                     // override (srfe_set_tuning)
            for (size_t i = 0; i < e->dct_vars.size(); ++i) if (e->dct_vars[i].cb == tune(TUNE_DCT_CB)) c->dct_var = (int)i;
        if (tune(TUNE_DCT_PQ) > 0) c->dct_pq = std::min(2, tune(TUNE_DCT_PQ));
        blob = align16(e->blob_common + e->dct_vars[c->dct_var].bytes);
    }
    c->blob = blob;
    int scratch = hw * (e->n_fft == 512 ? FftGeom<512>::SCRATCH_P2 : FftGeom<640>::SCRATCH_P2) * 8;
    if (e->family == FAM_SPEC && kp.layout == SRFE_LAYOUT_FT)        // CTA-wide [bin][2 hw + 2] transposition tile
        scratch = std::max(scratch, align16(e->n_bins * (2 * hw + 2) * 4));
    int tile = 0;
    c->ctile_off = blob;
    if (e->family == FAM_MFCC) {
        const int npairs = (kp.T + 1) / 2;
        tile = align16((npairs + 1) * kp.tile_stride * 8 + (npairs + 1) * 8);
        const int TC = kp.T + 2;                                     // upper bound of the kernel's row stride
        const int cstat = align16(kp.n_mfcc * TC * 4), call = (1 + kp.n_deltas) * kp.n_mfcc * TC * 4;
        if (cstat > scratch) scratch = cstat;
        if (call > scratch + tile) return -1;
    }
    c->scratch = scratch;
    c->tile = tile;
    c->smem = blob + scratch + tile;
    return c->smem;
}

// Warps per CTA, CTAs per SM and clips per group: fill whole rounds of 4*warps frames, keep as many
// warps resident as the shared-memory budget allows, prefer two CTAs per SM (their phases overlap).
static int pick_config(const Entry* e, const KParams& kp, const DevInfo& di, Config* out) {
    double best = -1.0;
    Config bc{0, 0, 1, 0, 0, 0, 0, 0, -1, 1};
    int cpc_max = (e->family == FAM_MFCC) ? 1 : 8;
    if (8.0 * kp.T * kp.T >= 4294967296.0) cpc_max = 1;              // magic division range (frame_pos)
    const int per_sm = 228 * 1024;                                   // B200: 228 KB per SM, 1 KB reserved per CTA
    for (int ctas = 1; ctas <= 2; ++ctas) {
        const int budget = std::min(di.smem_optin, per_sm / ctas - 1024);
        for (int warps = (ctas == 1 ? 8 : 4); warps <= (ctas == 1 ? kMaxThreads / 32 : kMaxThreads / 64); ++warps) {
            Config pl{};
            const int smem = smem_plan(e, kp, warps, budget, &pl);
            if (smem < 0 || smem > budget) continue;
            for (int cpc = 1; cpc <= cpc_max; ++cpc) {
                // small batches: trading CTAs for clips per group is paid for by the wave factor below (8 clips as one group
                // would run on one SM: 1 / slots), but a grid that is ALMOST full may be the best cut -- 1,024 clips as 147
                // groups of 7 finish after 7 clips' time, as 171 groups of 6 only after 12
                const long long groups = ((long long)kp.n_clips + cpc - 1) / cpc, slots = (long long)di.sms * ctas;
                const long long nf = (long long)cpc * kp.T, per_round = 4LL * warps;
                const long long rounds = (nf + per_round - 1) / per_round;
                // Fitted to scripts/tune.py sweeps on the B200 (profiles/r1_notes.md): throughput ~ round efficiency x
                // (resident warps - 2.3)^0.8.  Kernels with CTA barriers (MFCC epilogue, FT spectrogram tile) gain from a
                // second co-resident CTA whose frame phase fills the first one's serial phase (MFCC ~1.6x at equal warps,
                // and partly filled rounds then cost half as much); barrier-free kernels (TF spectrogram, FBANK) prefer
                // one wide CTA and many clips per group (fewer group prologues).  Round 2: any clips-per-group count (not
                // only powers of two) and the persistent grid's own quantisation -- the busiest CTA runs
                // ceil(groups / slots) groups -- so that BASELINE's small batches (cfg2: 1,024 clips = 6.9 clips per SM)
                // are cut into groups that fill whole rounds AND whole waves.
                const bool ft = e->family == FAM_SPEC && kp.layout == SRFE_LAYOUT_FT;
                const bool mf = e->family == FAM_MFCC;
                double eff = (double)nf / (double)(rounds * per_round);
                if (mf && ctas == 2) eff = std::sqrt(eff);
                eff *= (double)groups / (double)(((groups + slots - 1) / slots) * slots);
                double score = eff * std::pow(std::max(1.0, (double)(ctas * warps) - 2.3), 0.8);
                if (ctas == 2) score *= mf ? 1.6 : ft ? 1.1 : 0.92;
                score *= ft ? 1.0 - 0.03 * std::log2((double)cpc) : 1.0 + 0.01 * std::log2((double)cpc);
                if (score > best + 1e-9) { best = score; bc = pl; bc.warps = warps; bc.ctas = ctas; bc.cpc = cpc; }
            }
        }
    }
    if (best < 0) return fail(SRFE_ERR_TOO_LARGE, "clip too long: the per-clip tile does not fit in shared memory");
    // overrides (srfe_set_tuning)
    const int ow = tune(TUNE_WARPS), oc = tune(TUNE_CTAS), op = tune(TUNE_CPC);
    if (ow > 0 || oc > 0 || op > 0) {
        const int w_ = ow > 0 ? std::min(kMaxThreads / 32, std::max(1, ow)) : bc.warps;
        const int c_ = oc > 0 ? std::min(8, std::max(1, oc)) : bc.ctas;
        const int p_ = (op > 0 && cpc_max > 1) ? op : bc.cpc;
        const int budget = std::min(di.smem_optin, per_sm / c_ - 1024);
        const int smem = smem_plan(e, kp, w_, budget, &bc);
        bc.warps = w_; bc.ctas = c_; bc.cpc = p_;
        if (smem < 0 || smem > budget) return fail(SRFE_ERR_TOO_LARGE, "srfe_set_tuning: the warps / ctas override does not fit in shared memory");
    }
    *out = bc;
    return SRFE_OK;
}

struct ConfigKey {
    const Entry* e; int T, n_clips, layout; const DevInfo* di;
    bool operator<(const ConfigKey& o) const {
        return std::tie(e, T, n_clips, layout, di) < std::tie(o.e, o.T, o.n_clips, o.layout, o.di);
    }
};
static std::map<ConfigKey, Config> g_configs;                       // guarded by g_mu

static int launch(const Entry* e, KParams kp, bool i16, cudaStream_t st) {
    if (e->family == FAM_MFCC && kp.n_deltas > 0 && kp.T < 2) return fail(SRFE_ERR_UNSUPPORTED, "mfcc: deltas need at least 2 frames");
    if (kp.n_clips == 0 || kp.T == 0) return SRFE_OK;
    DevInfo* di = nullptr;
    int rc = dev_info(&di);
    if (rc != SRFE_OK) return rc;
    const int jlo = kp.w_lo / 32, jhi = (kp.w_hi + 31) / 32;
    const bool a400 = e->mel_ng == 8 && e->mel_code == 0xa400u, e500 = e->mel_ng == 8 && e->mel_code == 0xe500u;
    // MFCC: the tcgen05 kernel when the parameter set and clip length fit it (n_mels % 16 == 0, <= 128 frames, 112 with
    // deltas) and the clip has at least 48 frames.  Measured on the B200 (16,384 clips, M clips/s, classic / tcgen05, after
    // the read-out warps stopped polling in parallel): R-MFCC parameters at 32 / 40 / 51 / 64 / 101 frames 26.0 / 20.3,
    // 18.0 / 20.1, 15.7 / 17.0, 13.8 / 15.0, 8.5 / 10.2; C-MFCC parameters 32.5 / 23.2, 23.1 / 22.3, 20.8 / 21.3, 19.2 / 20.5,
    // 11.5 / 15.0 -- short clips fill too little of the 128-row MMA tile.  The rule depends on the clip length only, never on
    // the batch size (a clip's features must not depend on its batch); at 64 clips per call the classic kernel is the
    // quicker one by ~5 us.  srfe_set_tuning("mfcc_tc", 1 | 2) forces either.
    if (e->family == FAM_MFCC && tune(TUNE_MFCC_TC) != 1 && (kp.T >= 48 || tune(TUNE_MFCC_TC) == 2)) {
        int smem_tc = 0;
        KParams kt = kp;
        if (plan_tc(e, kt, *di, &smem_tc)) {
            const int grid = std::min(kt.n_clips, di->sms);
#define SRFE_GO_TC(N, JLO, JHI, NG, CODE, MB)                                                                    \
    return i16 ? launch_tc<N, JLO, JHI, NG, CODE, short, MB>(kt, di->index, grid, smem_tc, st)                   \
               : launch_tc<N, JLO, JHI, NG, CODE, float, MB>(kt, di->index, grid, smem_tc, st)
            // mel batch (mel_project): all eight band-sum groups before their emits when the frame warps set the pace
            // (no deltas: +2.5 % on the headline shape), one group at a time when the read-out warps do (deltas)
            if (e->n_fft == 512) {
                if (jlo >= 1 && jhi <= 15) {
                    if (a400 && kt.n_deltas == 0) SRFE_GO_TC(512, 1, 15, 8, 0xa400u, 8);
                    if (a400) SRFE_GO_TC(512, 1, 15, 8, 0xa400u, 1);
                    SRFE_GO_TC(512, 1, 15, 0, 0u, 1);
                }
                SRFE_GO_TC(512, 0, 16, 0, 0u, 1);
            } else {
                if (e500 && kt.n_deltas == 0) SRFE_GO_TC(640, 0, 20, 8, 0xe500u, 8);
                if (e500) SRFE_GO_TC(640, 0, 20, 8, 0xe500u, 1);
                SRFE_GO_TC(640, 0, 20, 0, 0u, 1);
            }
#undef SRFE_GO_TC
        } else if (tune(TUNE_MFCC_TC) == 2) {
            return fail(SRFE_ERR_UNSUPPORTED, "mfcc_tc = 2: this parameter set / clip length does not fit the tcgen05 MFCC kernel");
        }
    }
    // FBANK: the tcgen05 kernel (filter projection as bf16 hi / mid MMAs, weights resident in TMEM, srfe_fbank_tc.cuh).
    // Measured against the classic kernel on the B200 (M clips/s, classic / tcgen05): R-FBANK 15.1 / 14.5 at 1,024 clips,
    // 17.4 / 17.4 at 4,096, 18.4 / 18.5 at 16,384; C-FBANK (40 filters: two of the four read-out quadrants idle) 16.5 / 15.6,
    // 17.7 / 18.0, 18.5 / 19.2; below ~1,000 clips the classic kernel wins by 5-15 % (pipeline fill per CTA).  A tie -- and
    // the two kernels agree to 1.4e-4, not bit for bit, so a batch-size rule would make a clip's features depend on the batch
    // (or shard) it arrives in, which the product promises they never do (test_large_batch_properties,
    // test_shards_concatenate_bit_exact).  Hence OPT-IN only: srfe_set_tuning("fbank_tc", 2); 0 / 1 = the classic kernel.
    if (e->family == FAM_FBANK && tune(TUNE_FBANK_TC) == 2) {
        int smem_ft = 0;
        KParams kt = kp;
        if (plan_ft(e, kt, *di, &smem_ft)) {
            const int grid = std::min(kt.n_clips, di->sms);
#define SRFE_GO_FT(JLO, JHI)                                                                                     \
    return i16 ? launch_ft<JLO, JHI, short>(kt, di->index, grid, smem_ft, st)                                     \
               : launch_ft<JLO, JHI, float>(kt, di->index, grid, smem_ft, st)
            if (jlo == 0 && jhi <= 13) SRFE_GO_FT(0, 13);
            SRFE_GO_FT(0, 16);
#undef SRFE_GO_FT
        } else {
            return fail(SRFE_ERR_UNSUPPORTED, "fbank_tc = 2: this parameter set does not fit the tcgen05 FBANK kernel");
        }
    }
    Config cfg;
    {   // the choice depends on (parameter set, frames per clip, batch size up to the point where the grid is full):
        // remember it -- small-batch callers (the reference's ensemble drivers run batch_size = 1) pay for the search once
        bool overridden = false;
        for (int k = TUNE_WARPS; k <= TUNE_DCT_PQ; ++k) overridden = overridden || tune(k) > 0;
        const ConfigKey key{e, kp.T, std::min(kp.n_clips, 64 * di->sms), kp.layout, di};   // beyond that the grid's quantisation is < 2 %
        bool hit = false;
        if (!overridden) {
            std::shared_lock<std::shared_mutex> lk(g_mu);
            auto it = g_configs.find(key);
            if (it != g_configs.end()) { cfg = it->second; hit = true; }
        }
        if (!hit) {
            rc = pick_config(e, kp, *di, &cfg);
            if (rc != SRFE_OK) return rc;
            if (!overridden) {
                std::unique_lock<std::shared_mutex> lk(g_mu);
                if (g_configs.size() >= 4096) g_configs.clear();    // variable-length callers: bounded memory
                g_configs[key] = cfg;
            }
        }
    }
    kp.cpc = cfg.cpc;
    kp.n_groups = (kp.n_clips + cfg.cpc - 1) / cfg.cpc;
    kp.sm_scratch = cfg.blob;
    kp.sm_tile = kp.sm_scratch + cfg.scratch;
    kp.blob_bytes = cfg.blob;
    if (cfg.dct_var >= 0) {
        const Entry::DctVar& v = e->dct_vars[cfg.dct_var];
        kp.blob_bytes = e->blob_common;
        kp.dct_cb = v.cb; kp.dct_pq = cfg.dct_pq; kp.dct_nbe = v.nbe; kp.dct_nbo = v.nbo;
        kp.dct_src = v.off; kp.dct_bytes = v.bytes; kp.off_dfold = e->blob_common;
    }
    kp.sm_ctile = cfg.ctile_off;
    kp.t_magic = (unsigned)((0x100000000ULL + (unsigned long long)kp.T - 1) / (unsigned long long)kp.T);
#ifdef SRFE_DEV
    { const char* v = getenv("SRFE_DEBUG"); kp.debug = (v && *v) ? atoi(v) : 0; }   // developer build only: phase-skipping switches
#endif
    const int smem = cfg.smem;
    const int grid = std::min(kp.n_groups, di->sms * cfg.ctas), threads = 32 * cfg.warps;
    if (e->family == FAM_SPEC && tune(TUNE_STAGE) == 2) {
        // Spectrogram with TMA-staged frames (srfe_spec_staged_kernel): every half-warp prefetches its next frame pair into
        // shared memory with cp.async.bulk + mbarrier while it transforms the current one.  Built, bit-identical, measured on
        // B200 -- and NOT dispatched by default: C-SPEC TF 33.4 vs 35.4 M clips/s, C-SPEC FT 24.4 vs 24.5, R-SPEC TF (int16)
        // 22.1 vs 25.6; the direct LDG.64 path already keeps enough loads in flight, and the staged copy costs an extra
        // trip through the shared-memory pipe, which is the scarcer resource here (profiles/r2_notes.md).  Opt in with
        // srfe_set_tuning("stage", 2).  Needs 16-byte aligned frame starts and room for one 2-frame buffer per half-warp.
        if (cfg.ctas != 1) return fail(SRFE_ERR_UNSUPPORTED, "stage = 2: needs the one-CTA-per-SM plan");
        const size_t elem = i16 ? 2 : 4;
        const int stage = 2 * cfg.warps * 2 * e->n_fft * (int)elem, mb = 2 * cfg.warps * 8;
        const bool aligned = ((uintptr_t)kp.pcm % 16 == 0) && ((size_t)kp.clip_stride * elem % 16 == 0) && ((size_t)kp.hop * elem % 16 == 0);
        const int total = ((smem + 127) / 128) * 128 + stage + mb;
        if (aligned && total <= di->smem_optin) {
            kp.sm_stage = ((smem + 127) / 128) * 128;
            kp.sm_stage_mb = kp.sm_stage + stage;
            kp.stage_phase0 = 0;
            if (e->n_fft == 512) return i16 ? launch_staged<512, short>(kp, di->index, grid, threads, total, st) : launch_staged<512, float>(kp, di->index, grid, threads, total, st);
            return i16 ? launch_staged<640, short>(kp, di->index, grid, threads, total, st) : launch_staged<640, float>(kp, di->index, grid, threads, total, st);
        } else if (tune(TUNE_STAGE) == 2) {
            return fail(SRFE_ERR_UNSUPPORTED, "stage = 2: frames not 16-byte aligned or no shared memory left for the staging buffers");
        }
    }
    // Curated instantiation list.  Window extents (units of 32 samples) and the preset shapes of the mel
    // bank / DCT get specialised kernels; everything else runs the generic ones.
    //   mel {8 groups, 0xa400}: 128 Slaney mels @ n_fft 512 and the reference's 120 HTK bands @ 512
    //   mel {8 groups, 0xe500}: 128 Slaney mels @ n_fft 640
#define SRFE_GO(N, FAM, JLO, JHI, NG, CODE)                                                                      \
    return i16 ? launch_k<N, FAM, JLO, JHI, NG, CODE, short>(kp, di->index, grid, threads, smem, st)             \
               : launch_k<N, FAM, JLO, JHI, NG, CODE, float>(kp, di->index, grid, threads, smem, st)
    if (e->n_fft == 512) {
        switch (e->family) {
            case FAM_SPEC: SRFE_GO(512, FAM_SPEC, 0, 16, 0, 0u);
            case FAM_FBANK:
                if (jlo == 0 && jhi <= 13) {
                    if (a400) SRFE_GO(512, FAM_FBANK, 0, 13, 8, 0xa400u);
                    SRFE_GO(512, FAM_FBANK, 0, 13, 0, 0u);
                }
                SRFE_GO(512, FAM_FBANK, 0, 16, 0, 0u);
            default:
                if (jlo >= 1 && jhi <= 15) {
                    if (a400) SRFE_GO(512, FAM_MFCC, 1, 15, 8, 0xa400u);
                    SRFE_GO(512, FAM_MFCC, 1, 15, 0, 0u);
                }
                SRFE_GO(512, FAM_MFCC, 0, 16, 0, 0u);
        }
    } else {
        switch (e->family) {
            case FAM_SPEC: SRFE_GO(640, FAM_SPEC, 0, 20, 0, 0u);
            case FAM_FBANK: SRFE_GO(640, FAM_FBANK, 0, 20, 0, 0u);
            default:
                if (e500) SRFE_GO(640, FAM_MFCC, 0, 20, 8, 0xe500u);
                SRFE_GO(640, FAM_MFCC, 0, 20, 0, 0u);
        }
    }
#undef SRFE_GO
}

// ------------------------------------------------------------------------------
// spectrogram + fbank of the same batch in one launch (srfe_spec_fbank_kernel)
// ------------------------------------------------------------------------------
template <int NA, int NB, int JLO, int JHI, int NG, unsigned CODE, typename SAMP>
static int launch_fused_k(const KParams& ka, const KParams& kb, int off_b, int dev, int grid, int smem_bytes, cudaStream_t st) {
    auto kern = srfe_spec_fbank_kernel<NA, NB, JLO, JHI, NG, CODE, SAMP>;
    static std::atomic<int> attr_set[kMaxDevices];
    if (attr_set[dev].load(std::memory_order_acquire) < smem_bytes) {
        SRFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int cur = attr_set[dev].load(std::memory_order_relaxed);
        while (cur < smem_bytes && !attr_set[dev].compare_exchange_weak(cur, smem_bytes, std::memory_order_release)) {}
    }
    kern<<<grid, kMaxThreads, smem_bytes, st>>>(ka, kb, off_b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "srfe_spec_fbank_kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SRFE_OK;
}

static int run_fused(const void* pcm, bool i16, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                     const srfe_spec_params* ps, const srfe_fbank_params* pf, float* out_spec, float* out_fbank,
                     cudaStream_t st, int64_t Ta, int64_t Tb) {
    Entry *ea = nullptr, *eb = nullptr;
    int rc = get_entry(FAM_SPEC, *ps, &ea);
    if (rc == SRFE_OK) rc = get_entry(FAM_FBANK, *pf, &eb);
    if (rc != SRFE_OK) return rc;
    if (eb->n_fft != 512) return fail(SRFE_ERR_UNSUPPORTED, "spec+fbank in one launch: the fbank side needs n_fft = 512");
    if (n_clips == 0 || Ta == 0 || Tb == 0) return SRFE_OK;
    DevInfo* di = nullptr;
    rc = dev_info(&di);
    if (rc != SRFE_OK) return rc;
    auto up = [](int x, int a) { return (x + a - 1) / a * a; };
    const int warps = kMaxThreads / 32, hw = 2 * warps;
    auto fill = [&](const Entry* e, KParams& k, float* out, int64_t T) {
        k = e->kp;
        k.pcm = pcm; k.out = out; k.clip_stride = clip_stride; k.n_clips = (int)n_clips; k.n_samples = (int)n_samples; k.T = (int)T;
        k.blob_bytes = align16(e->blob_smem);
        int scratch = hw * (e->n_fft == 512 ? FftGeom<512>::SCRATCH_P2 : FftGeom<640>::SCRATCH_P2) * 8;
        if (e->family == FAM_SPEC && k.layout == SRFE_LAYOUT_FT) scratch = std::max(scratch, align16(e->n_bins * (2 * hw + 2) * 4));
        k.sm_scratch = up(k.blob_bytes, 128);
        k.sm_tile = k.sm_scratch + scratch;
        k.t_magic = (unsigned)((0x100000000ULL + (unsigned long long)T - 1) / (unsigned long long)T);
        return up(k.sm_tile, 128);
    };
    KParams ka, kb;
    const int off_b = fill(ea, ka, out_spec, Ta);
    const int smem = off_b + fill(eb, kb, out_fbank, Tb);
    if (smem > di->smem_optin) return fail(SRFE_ERR_TOO_LARGE, "spec+fbank in one launch: shared memory plan does not fit");
    int cpc = 8;                                                       // small batches: never trade CTAs for clips per group
    while (cpc > 1 && ((n_clips + cpc - 1) / cpc < 2 * di->sms || 8.0 * cpc * std::max(Ta, Tb) >= 4294967296.0 / std::max(Ta, Tb))) cpc /= 2;
    if (tune(TUNE_CPC) > 0) cpc = tune(TUNE_CPC);
    ka.cpc = kb.cpc = cpc;
    ka.n_groups = kb.n_groups = (int)((n_clips + cpc - 1) / cpc);
    const int grid = std::min(ka.n_groups, di->sms);
    const int jlo = kb.w_lo / 32, jhi = (kb.w_hi + 31) / 32;
    const bool a400 = eb->mel_ng == 8 && eb->mel_code == 0xa400u;
#define SRFE_GO_F(NA, JLO, JHI, NG, CODE)                                                                                   \
    return i16 ? launch_fused_k<NA, 512, JLO, JHI, NG, CODE, short>(ka, kb, off_b, di->index, grid, smem, st)               \
               : launch_fused_k<NA, 512, JLO, JHI, NG, CODE, float>(ka, kb, off_b, di->index, grid, smem, st)
    if (ea->n_fft == 640) {
        if (jlo == 0 && jhi <= 13) { if (a400) SRFE_GO_F(640, 0, 13, 8, 0xa400u); SRFE_GO_F(640, 0, 13, 0, 0u); }
        SRFE_GO_F(640, 0, 16, 0, 0u);
    }
    if (jlo == 0 && jhi <= 13) { if (a400) SRFE_GO_F(512, 0, 13, 8, 0xa400u); SRFE_GO_F(512, 0, 13, 0, 0u); }
    SRFE_GO_F(512, 0, 16, 0, 0u);
#undef SRFE_GO_F
}

static int check_buffers(const void* pcm, int elem, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const void* p,
                         const float* out) {
    if (!p) return fail(SRFE_ERR_BAD_ARG, "params is NULL");
    if (n_clips < 0 || n_samples <= 0) return fail(SRFE_ERR_BAD_ARG, "n_clips must be >= 0 and n_samples > 0");
    if (n_clips > 0 && (!pcm || !out)) return fail(SRFE_ERR_BAD_ARG, "pcm/out is NULL");
    if (clip_stride < n_samples) return fail(SRFE_ERR_BAD_ARG, "clip_stride < n_samples");
    if (n_clips > 1 && (clip_stride & 1)) return fail(SRFE_ERR_BAD_ARG, "clip_stride must be even (rows aligned to a sample pair)");
    if (((uintptr_t)pcm & (uintptr_t)(2 * elem - 1)) || ((uintptr_t)out & 3))
        return fail(SRFE_ERR_BAD_ARG, "pcm must be aligned to a sample pair (8 bytes for float32, 4 for int16), out to 4 bytes");
    if (n_clips > 0x7fffffffLL || n_samples > (1 << 24)) return fail(SRFE_ERR_TOO_LARGE, "n_clips / n_samples too large");
    return SRFE_OK;
}

template <typename P>
static int run_device(int family, const void* pcm, bool i16, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                      const P* p, float* out, cudaStream_t st, int64_t T) {
    Entry* e = nullptr;
    int rc = get_entry(family, *p, &e);
    if (rc != SRFE_OK) return rc;
    KParams kp = e->kp;
    kp.pcm = pcm;
    kp.out = out;
    kp.clip_stride = clip_stride;
    kp.n_clips = (int)n_clips;
    kp.n_samples = (int)n_samples;
    kp.T = (int)T;
    return launch(e, kp, i16, st);
}

// ------------------------------------------------------------------------------
// host-buffer path: chunked H2D -> kernel -> D2H on two streams
// ------------------------------------------------------------------------------
// One workspace per device (not per thread): two streams, two device in / out buffers and -- only when a caller hands
// over PAGEABLE memory, which is what the reference's DataLoader does (training.py:77, no pin_memory) -- two pinned
// staging buffers each way.  A per-device mutex serialises host-path calls on the same GPU (they would share its
// PCIe link anyway); srfe_release_host_workspace() frees everything.
struct HostWs {
    std::mutex mu;
    bool init = false;
    cudaStream_t st[2] = {nullptr, nullptr};
    void* d_in[2] = {nullptr, nullptr};
    float* d_out[2] = {nullptr, nullptr};
    void* h_in[2] = {nullptr, nullptr};         // pinned staging (pageable callers only)
    float* h_out[2] = {nullptr, nullptr};
    size_t cap_in = 0, cap_out = 0, cap_hin = 0, cap_hout = 0;
};
static HostWs g_ws[kMaxDevices];

template <typename T, typename AllocFn, typename FreeFn>
static int grow_pair(T* ptr[2], size_t& cap, size_t bytes, AllocFn alloc, FreeFn release, const char* what) {
    if (bytes <= cap) return SRFE_OK;
    for (int i = 0; i < 2; ++i) { if (ptr[i]) release(ptr[i]); ptr[i] = nullptr; }
    cap = 0;                                     // nothing usable until both allocations succeed
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = alloc((void**)&ptr[i], bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();                  // do not leave the error for the next launch to trip over
            for (int j = 0; j < 2; ++j) { if (ptr[j]) release(ptr[j]); ptr[j] = nullptr; }
            return cuda_fail(e, what);
        }
    }
    cap = bytes;
    return SRFE_OK;
}

static int host_ws_prepare(HostWs& w, size_t in_bytes, size_t out_bytes, bool stage_in, bool stage_out) {
    if (!w.init) {
        for (int i = 0; i < 2; ++i) SRFE_CUDA(cudaStreamCreateWithFlags(&w.st[i], cudaStreamNonBlocking));
        w.init = true;
    }
    auto dmalloc = [](void** q, size_t n) { return cudaMalloc(q, n); };
    auto dfree = [](void* q) { cudaFree(q); };
    auto hmalloc = [](void** q, size_t n) { return cudaMallocHost(q, n); };
    auto hfree = [](void* q) { cudaFreeHost(q); };
    int rc = grow_pair(w.d_in, w.cap_in, in_bytes, dmalloc, dfree, "cudaMalloc (host-path input buffer)");
    if (rc == SRFE_OK) rc = grow_pair(w.d_out, w.cap_out, out_bytes, dmalloc, dfree, "cudaMalloc (host-path output buffer)");
    if (rc == SRFE_OK && stage_in) rc = grow_pair(w.h_in, w.cap_hin, in_bytes, hmalloc, hfree, "cudaMallocHost (input staging)");
    if (rc == SRFE_OK && stage_out) rc = grow_pair(w.h_out, w.cap_hout, out_bytes, hmalloc, hfree, "cudaMallocHost (output staging)");
    return rc;
}

static bool is_pageable(const void* host_ptr) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Large staging copies with non-temporal stores: the destination (a pinned staging buffer on the way in, the caller's
// result buffer on the way out) is not read again by this core, and a cached store would first fetch every destination
// line (read-for-ownership): 3 bytes of memory traffic per byte copied instead of 2.  glibc's memcpy only switches to
// streaming stores above a per-call size the per-thread slices here never reach.
static void copy_stream(char* dst, const char* src, size_t n) {
#if defined(__x86_64__) || defined(_M_X64)
    if (n >= (size_t)(256u << 10)) {
        size_t head = (size_t)((16 - ((uintptr_t)dst & 15)) & 15);
        if (head) { std::memcpy(dst, src, head); dst += head; src += head; n -= head; }
        const size_t blocks = n / 64;
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(src) + 0), b = _mm_loadu_si128((const __m128i*)(src) + 1);
            const __m128i c = _mm_loadu_si128((const __m128i*)(src) + 2), d = _mm_loadu_si128((const __m128i*)(src) + 3);
            _mm_stream_si128((__m128i*)(dst) + 0, a); _mm_stream_si128((__m128i*)(dst) + 1, b);
            _mm_stream_si128((__m128i*)(dst) + 2, c); _mm_stream_si128((__m128i*)(dst) + 3, d);
            src += 64; dst += 64;
        }
        n -= blocks * 64;
        if (n) std::memcpy(dst, src, n);
        _mm_sfence();
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

// rows of `row_bytes` from a strided source into a dense destination, split over a few host threads (one copy
// stream per thread: a single core moves ~10 GB/s, the PCIe link ~50)
static void gather_rows(char* dst, const char* src, size_t rows, size_t row_bytes, size_t src_stride_bytes) {
    const size_t total = rows * row_bytes;
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = std::max(1u, std::min(nthr ? nthr / 2 : 1u, 8u));
    if (total < (size_t)(4u << 20)) nthr = 1;
    auto work = [=](size_t r0, size_t r1) {
        if (src_stride_bytes == row_bytes) copy_stream(dst + r0 * row_bytes, src + r0 * row_bytes, (r1 - r0) * row_bytes);
        else for (size_t r = r0; r < r1; ++r) std::memcpy(dst + r * row_bytes, src + r * src_stride_bytes, row_bytes);
    };
    if (nthr == 1) { work(0, rows); return; }
    std::vector<std::thread> th;
    const size_t per = (rows + nthr - 1) / nthr;
    size_t done_to = std::min(rows, per);                  // rows [0, per) are the calling thread's
    for (unsigned t = 1; t < nthr; ++t) {
        const size_t r0 = std::min(rows, t * per), r1 = std::min(rows, r0 + per);
        if (r0 >= r1) break;
        try {
            th.emplace_back(work, r0, r1);
        } catch (...) {                                    // no more threads to be had: the caller copies the rest itself
            break;
        }
        done_to = r1;
    }
    work(0, std::min(rows, per));
    if (done_to < rows) work(done_to, rows);
    for (auto& t : th) t.join();
}

template <typename P>
static int run_host(int family, const void* pcm, bool i16, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                    const P* p, float* out, int device, int64_t T, int64_t out_per_clip) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return fail(SRFE_ERR_NO_DEVICE, "no CUDA device"); }
    if (device < 0 || device >= kMaxDevices) return fail(SRFE_ERR_BAD_ARG, "device index out of range");
    if (n_clips == 0) return SRFE_OK;
    SRFE_CUDA(cudaSetDevice(device));
    const size_t elem = i16 ? 2 : 4;
    const size_t row_bytes = (size_t)n_samples * elem, out_row_bytes = (size_t)out_per_clip * 4;
    const bool stage_in = is_pageable(pcm), stage_out = is_pageable(out);
    // ~32 MB chunks (524 one-second float32 clips).  Pageable callers: the host-side gather of chunk i+1 overlaps the
    // transfer and kernel of chunk i.  Pinned callers: the H2D engine is busy from the first byte to the last either way, so
    // what a chunk size decides is the tail after the last H2D byte -- the last chunk's kernel and D2H (0.74 ms with
    // 2048-clip chunks, 0.19 ms now, of a 20 ms step of 16,384 clips).
    int64_t chunk = std::max<int64_t>(1, (int64_t)((32u << 20) / row_bytes));
    if (n_clips < chunk) chunk = n_clips;
    HostWs& w = g_ws[device];
    std::lock_guard<std::mutex> lk(w.mu);
    int rc = host_ws_prepare(w, (size_t)chunk * row_bytes, (size_t)chunk * out_row_bytes, stage_in, stage_out);
    int64_t pend_c0[2] = {-1, -1}, pend_nc[2] = {0, 0};          // staged outputs not yet copied back, per slot
    auto drain = [&](int s) -> int {
        if (pend_c0[s] < 0) return SRFE_OK;
        cudaError_t ce = cudaStreamSynchronize(w.st[s]);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamSynchronize");
        // (a fresh pageable result buffer takes its first-touch page faults here: spread them over the copy threads)
        gather_rows((char*)(out + pend_c0[s] * out_per_clip), (const char*)w.h_out[s], (size_t)pend_nc[s], out_row_bytes, out_row_bytes);
        pend_c0[s] = -1;
        return SRFE_OK;
    };
    for (int64_t c0 = 0, i = 0; rc == SRFE_OK && c0 < n_clips; c0 += chunk, ++i) {
        const int64_t nc = (n_clips - c0 < chunk) ? n_clips - c0 : chunk;
        const int s = (int)(i & 1);
        const char* src = (const char*)pcm + (size_t)c0 * clip_stride * elem;
        cudaError_t ce;
        if (stage_in) {
            // slot s was last used by chunk i-2: its H2D must be done before the staging buffer is overwritten
            if (stage_out) rc = drain(s); else { ce = cudaStreamSynchronize(w.st[s]); if (ce != cudaSuccess) rc = cuda_fail(ce, "cudaStreamSynchronize"); }
            if (rc != SRFE_OK) break;
            gather_rows((char*)w.h_in[s], src, (size_t)nc, row_bytes, (size_t)clip_stride * elem);
            ce = cudaMemcpyAsync(w.d_in[s], w.h_in[s], (size_t)nc * row_bytes, cudaMemcpyHostToDevice, w.st[s]);
        } else {
            if (stage_out) { rc = drain(s); if (rc != SRFE_OK) break; }
            ce = cudaMemcpy2DAsync(w.d_in[s], row_bytes, src, (size_t)clip_stride * elem, row_bytes, (size_t)nc,
                                   cudaMemcpyHostToDevice, w.st[s]);
        }
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "H2D"); break; }
        rc = run_device(family, w.d_in[s], i16, nc, n_samples, n_samples, p, w.d_out[s], w.st[s], T);
        if (rc != SRFE_OK) break;
        float* dst = stage_out ? w.h_out[s] : out + c0 * out_per_clip;
        ce = cudaMemcpyAsync(dst, w.d_out[s], (size_t)nc * out_row_bytes, cudaMemcpyDeviceToHost, w.st[s]);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "D2H"); break; }
        if (stage_out) { pend_c0[s] = c0; pend_nc[s] = nc; }
    }
    for (int i = 0; i < 2; ++i) {
        if (!w.st[i]) continue;
        if (stage_out && rc == SRFE_OK) { rc = drain(i); continue; }
        cudaError_t ce = cudaStreamSynchronize(w.st[i]);
        if (ce != cudaSuccess && rc == SRFE_OK) rc = cuda_fail(ce, "cudaStreamSynchronize");
    }
    cudaSetDevice(prev);
    return rc;
}

// Host rows -> a device buffer the caller owns, through the same pinned staging ring (pageable sources) or directly
// (pinned sources), stream-ordered on `st`: what `x.to(device)` does in the reference's forward, at PCIe speed instead
// of the driver's single staged copy.  The staging buffers are reused chunk by chunk, so the gathers run on the calling
// thread while earlier chunks are in flight; the function returns when the last chunk has been handed to the copy engine
// and its staging buffer may be reused (the device side completes in stream order).
static int upload_rows(const void* host, int64_t rows, int64_t row_bytes, int64_t row_stride_bytes, void* dev, int device, cudaStream_t st) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return fail(SRFE_ERR_NO_DEVICE, "no CUDA device"); }
    if (device < 0 || device >= kMaxDevices) return fail(SRFE_ERR_BAD_ARG, "device index out of range");
    if (rows == 0) return SRFE_OK;
    SRFE_CUDA(cudaSetDevice(device));
    int rc = SRFE_OK;
    if (!is_pageable(host)) {
        cudaError_t ce = cudaMemcpy2DAsync(dev, (size_t)row_bytes, host, (size_t)row_stride_bytes, (size_t)row_bytes, (size_t)rows, cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "H2D");
        cudaSetDevice(prev);
        return rc;
    }
    HostWs& w = g_ws[device];
    std::lock_guard<std::mutex> lk(w.mu);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(rows, (int64_t)((32u << 20) / (size_t)row_bytes)));
    rc = host_ws_prepare(w, 0, 0, false, false);
    if (rc == SRFE_OK) rc = grow_pair(w.h_in, w.cap_hin, (size_t)chunk * row_bytes, [](void** q, size_t n) { return cudaMallocHost(q, n); },
                                      [](void* q) { cudaFreeHost(q); }, "cudaMallocHost (input staging)");
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int i = 0; i < 2 && rc == SRFE_OK; ++i) {
        cudaError_t ce = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "cudaEventCreate");
    }
    bool used[2] = {false, false};
    for (int64_t r0 = 0, i = 0; rc == SRFE_OK && r0 < rows; r0 += chunk, ++i) {
        const int64_t nr = std::min(chunk, rows - r0);
        const int s2 = (int)(i & 1);
        if (used[s2]) { cudaError_t ce = cudaEventSynchronize(ev[s2]); if (ce != cudaSuccess) { rc = cuda_fail(ce, "cudaEventSynchronize"); break; } }
        gather_rows((char*)w.h_in[s2], (const char*)host + (size_t)r0 * row_stride_bytes, (size_t)nr, (size_t)row_bytes, (size_t)row_stride_bytes);
        cudaError_t ce = cudaMemcpyAsync((char*)dev + (size_t)r0 * row_bytes, w.h_in[s2], (size_t)nr * row_bytes, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[s2], st);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "H2D"); break; }
        used[s2] = true;
    }
    for (int i = 0; i < 2; ++i) {
        if (!ev[i]) continue;
        if (used[i]) { cudaError_t ce = cudaEventSynchronize(ev[i]); if (ce != cudaSuccess && rc == SRFE_OK) rc = cuda_fail(ce, "cudaEventSynchronize"); }
        cudaEventDestroy(ev[i]);
    }
    cudaSetDevice(prev);
    return rc;
}

static int release_host_ws() {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return SRFE_OK; }     // no device: nothing was allocated
    for (int d = 0; d < kMaxDevices; ++d) {
        HostWs& w = g_ws[d];
        std::lock_guard<std::mutex> lk(w.mu);
        if (!w.init) continue;
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
        for (int i = 0; i < 2; ++i) {
            if (w.st[i]) { cudaStreamSynchronize(w.st[i]); cudaStreamDestroy(w.st[i]); w.st[i] = nullptr; }
            if (w.d_in[i]) cudaFree(w.d_in[i]);
            if (w.d_out[i]) cudaFree(w.d_out[i]);
            if (w.h_in[i]) cudaFreeHost(w.h_in[i]);
            if (w.h_out[i]) cudaFreeHost(w.h_out[i]);
            w.d_in[i] = nullptr; w.d_out[i] = nullptr; w.h_in[i] = nullptr; w.h_out[i] = nullptr;
        }
        w.cap_in = w.cap_out = w.cap_hin = w.cap_hout = 0;
        w.init = false;
    }
    cudaSetDevice(prev);
    return SRFE_OK;
}

}  // namespace srfe

using namespace srfe;

// ------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------
extern "C" {

int srfe_version(void) { return SRFE_VERSION_MAJOR * 1000 + SRFE_VERSION_MINOR; }
const char* srfe_last_error_string(void) { return g_err.c_str(); }
int srfe_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SRFE_ERR_NO_DEVICE, cudaGetErrorString(e)); }
    return n;
}
int64_t srfe_launch_count(void) { return g_launches.load(); }
int srfe_release_host_workspace(void) { return release_host_ws(); }
int srfe_upload(const void* host, int64_t rows, int64_t row_bytes, int64_t row_stride_bytes, void* dev, int device, void* stream) {
    if (rows < 0 || row_bytes <= 0 || row_stride_bytes < row_bytes) return fail(SRFE_ERR_BAD_ARG, "upload: bad sizes");
    if (rows > 0 && (!host || !dev)) return fail(SRFE_ERR_BAD_ARG, "upload: host/dev is NULL");
    return upload_rows(host, rows, row_bytes, row_stride_bytes, dev, device, (cudaStream_t)stream);
}

static int spec_fbank(const void* pcm, bool i16, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* ps,
                      const srfe_fbank_params* pf, float* out_spec, float* out_fbank, void* stream) {
    int rc = check_buffers(pcm, i16 ? 2 : 4, n_clips, n_samples, clip_stride, ps, out_spec);
    if (rc == SRFE_OK) rc = check_buffers(pcm, i16 ? 2 : 4, n_clips, n_samples, clip_stride, pf, out_fbank);
    if (rc != SRFE_OK) return rc;
    const int64_t Ta = srfe_spec_out_shape(ps, n_samples, nullptr);
    if (Ta < 0) return (int)Ta;
    const int64_t Tb = srfe_fbank_out_shape(pf, n_samples, nullptr);
    if (Tb < 0) return (int)Tb;
    return run_fused(pcm, i16, n_clips, n_samples, clip_stride, ps, pf, out_spec, out_fbank, (cudaStream_t)stream, Ta, Tb);
}
int srfe_spec_fbank_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* ps,
                        const srfe_fbank_params* pf, float* out_spec, float* out_fbank, void* stream) {
    return spec_fbank(pcm, false, n_clips, n_samples, clip_stride, ps, pf, out_spec, out_fbank, stream);
}
int srfe_spec_fbank_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* ps,
                        const srfe_fbank_params* pf, float* out_spec, float* out_fbank, void* stream) {
    return spec_fbank(pcm, true, n_clips, n_samples, clip_stride, ps, pf, out_spec, out_fbank, stream);
}

int srfe_augment_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const int8_t* kind,
                     int64_t first_clip_index, const int16_t* noise_bank, const int64_t* noise_offsets, int32_t n_noise_files,
                     const srfe_augment_params* p, float* out, int8_t* op_out, void* stream) {
    if (!p) return fail(SRFE_ERR_BAD_ARG, "params is NULL");
    if (n_clips < 0 || n_samples <= 0 || clip_stride < n_samples) return fail(SRFE_ERR_BAD_ARG, "augment: bad sizes");
    if (n_clips > 0 && (!pcm || !out)) return fail(SRFE_ERR_BAD_ARG, "pcm/out is NULL");
    if (n_noise_files < 0 || (n_noise_files > 0 && (!noise_bank || !noise_offsets))) return fail(SRFE_ERR_BAD_ARG, "augment: noise bank is NULL");
    if (n_clips > 0x7fffffffLL || n_samples > (1 << 24) || first_clip_index < 0) return fail(SRFE_ERR_TOO_LARGE, "n_clips / n_samples too large");
    const bool needs_noise = kind != nullptr || p->noise_hi > p->noise_lo || p->snr_hi > p->snr_lo;
    if (needs_noise && n_noise_files == 0) return fail(SRFE_ERR_BAD_ARG, "augment: the noise / silence ops need a background-noise bank");
    if (p->shift_range < 0 || p->shift_range > n_samples) return fail(SRFE_ERR_BAD_ARG, "augment: shift_range out of [0, n_samples]");
    if (n_clips == 0) return SRFE_OK;
    AugParams a{};
    a.pcm = pcm; a.out = out; a.kind = (const signed char*)kind; a.op_out = (signed char*)op_out;
    a.bank = noise_bank; a.bank_off = (const long long*)noise_offsets;
    a.clip_stride = clip_stride; a.first_index = first_clip_index;
    a.n_clips = (int)n_clips; a.n_samples = (int)n_samples; a.n_files = n_noise_files;
    a.key0 = (unsigned)(p->seed & 0xffffffffu); a.key1 = (unsigned)(p->seed >> 32);
    a.shift_lo = p->shift_lo; a.shift_hi = p->shift_hi; a.noise_lo = p->noise_lo; a.noise_hi = p->noise_hi;
    a.snr_lo = p->snr_lo; a.snr_hi = p->snr_hi; a.pitch_lo = p->pitch_lo; a.pitch_hi = p->pitch_hi;
    a.speed_lo = p->speed_lo; a.speed_hi = p->speed_hi;
    a.shift_range = p->shift_range; a.noise_upper = p->noise_upper;
    const double snr_db[4] = {-5.0, 0.0, 5.0, 10.0};
    for (int i = 0; i < 4; ++i) a.snr_div[i] = std::pow(10.0, snr_db[i] / 10.0);
    srfe_augment_kernel<<<(unsigned)n_clips, 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "srfe_augment_kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SRFE_OK;
}
int srfe_set_tuning(const char* name, int value) {
    if (!name || value < 0) return fail(SRFE_ERR_BAD_ARG, "srfe_set_tuning: name is NULL or value < 0");
    for (int k = 0; k < TUNE_COUNT; ++k)
        if (std::strcmp(name, kTuningNames[k]) == 0) { g_tune[k].store(value); return SRFE_OK; }
    return fail(SRFE_ERR_BAD_ARG, std::string("srfe_set_tuning: unknown knob '") + name + "'");
}

#define SRFE_VALIDATE(p)                                            \
    if (!(p)) return fail(SRFE_ERR_BAD_ARG, "params is NULL");      \
    { const char* why = ""; int rc__ = validate(*(p), &why); if (rc__ != SRFE_OK) return fail(rc__, why); }

int64_t srfe_spec_out_shape(const srfe_spec_params* p, int64_t n, int64_t shape2[2]) {
    SRFE_VALIDATE(p)
    if (n < p->nperseg) return fail(SRFE_ERR_BAD_ARG, "spec: n_samples < nperseg");
    const int64_t T = spec_frames(*p, n), F = p->nperseg / 2 + 1;
    if (shape2) { shape2[0] = p->layout == SRFE_LAYOUT_FT ? F : T; shape2[1] = p->layout == SRFE_LAYOUT_FT ? T : F; }
    return T;
}
int64_t srfe_fbank_out_shape(const srfe_fbank_params* p, int64_t n, int64_t shape2[2]) {
    SRFE_VALIDATE(p)
    if (n < 1) return fail(SRFE_ERR_BAD_ARG, "fbank: n_samples < 1");
    const int64_t T = fbank_frames(*p, n);
    if (shape2) { shape2[0] = T; shape2[1] = p->nfilt; }
    return T;
}
int64_t srfe_mfcc_out_shape(const srfe_mfcc_params* p, int64_t n, int64_t shape2[2]) {
    SRFE_VALIDATE(p)
    if (n <= p->n_fft / 2) return fail(SRFE_ERR_BAD_ARG, "mfcc: reflect padding needs n_samples > n_fft/2");
    const int64_t T = mfcc_frames(*p, n), R = (int64_t)(1 + p->n_deltas) * p->n_mfcc;
    if (shape2) { shape2[0] = p->layout == SRFE_LAYOUT_FT ? R : T; shape2[1] = p->layout == SRFE_LAYOUT_FT ? T : R; }
    return T;
}

int64_t srfe_spec_bytes_per_clip(const srfe_spec_params* p, int64_t n) {
    int64_t s[2]; const int64_t T = srfe_spec_out_shape(p, n, s);
    return T < 0 ? T : n * 4 + s[0] * s[1] * 4;
}
int64_t srfe_fbank_bytes_per_clip(const srfe_fbank_params* p, int64_t n) {
    int64_t s[2]; const int64_t T = srfe_fbank_out_shape(p, n, s);
    return T < 0 ? T : n * 4 + s[0] * s[1] * 4;
}
int64_t srfe_mfcc_bytes_per_clip(const srfe_mfcc_params* p, int64_t n) {
    int64_t s[2]; const int64_t T = srfe_mfcc_out_shape(p, n, s);
    return T < 0 ? T : n * 4 + s[0] * s[1] * 4;
}

// The device entry points need no global-memory workspace (everything between PCM and features lives in registers and
// shared memory): the query exists so that a caller written against "caller owns output and workspace" has its answer.
int64_t srfe_spec_workspace_bytes(const srfe_spec_params* p, int64_t n_clips, int64_t n) {
    const int64_t T = srfe_spec_out_shape(p, n, nullptr);
    return T < 0 ? T : (n_clips < 0 ? (int64_t)fail(SRFE_ERR_BAD_ARG, "n_clips < 0") : 0);
}
int64_t srfe_fbank_workspace_bytes(const srfe_fbank_params* p, int64_t n_clips, int64_t n) {
    const int64_t T = srfe_fbank_out_shape(p, n, nullptr);
    return T < 0 ? T : (n_clips < 0 ? (int64_t)fail(SRFE_ERR_BAD_ARG, "n_clips < 0") : 0);
}
int64_t srfe_mfcc_workspace_bytes(const srfe_mfcc_params* p, int64_t n_clips, int64_t n) {
    const int64_t T = srfe_mfcc_out_shape(p, n, nullptr);
    return T < 0 ? T : (n_clips < 0 ? (int64_t)fail(SRFE_ERR_BAD_ARG, "n_clips < 0") : 0);
}

static int copy_out(const std::vector<double>& v, double* dst) {
    if (!dst) return fail(SRFE_ERR_BAD_ARG, "output pointer is NULL");
    std::memcpy(dst, v.data(), v.size() * sizeof(double));
    return SRFE_OK;
}
int srfe_spec_window_f64(const srfe_spec_params* p, double* win) { SRFE_VALIDATE(p) std::vector<double> w; spec_window(*p, w); return copy_out(w, win); }
int srfe_fbank_window_f64(const srfe_fbank_params* p, double* win) { SRFE_VALIDATE(p) std::vector<double> w; fbank_window(*p, w); return copy_out(w, win); }
int srfe_mfcc_window_f64(const srfe_mfcc_params* p, double* win) { SRFE_VALIDATE(p) std::vector<double> w; mfcc_window(*p, w); return copy_out(w, win); }
int srfe_fbank_filters_f64(const srfe_fbank_params* p, double* w) { SRFE_VALIDATE(p) std::vector<double> d; fbank_filters(*p, d); return copy_out(d, w); }
int srfe_mfcc_filters_f64(const srfe_mfcc_params* p, double* w) { SRFE_VALIDATE(p) std::vector<double> d; mfcc_filters(*p, d); return copy_out(d, w); }
int srfe_mfcc_dct_f64(const srfe_mfcc_params* p, double* d) { SRFE_VALIDATE(p) std::vector<double> v; mfcc_dct(*p, v); return copy_out(v, d); }

int srfe_spec_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* p,
                  float* out, void* stream) {
    int rc = check_buffers(pcm, sizeof(float), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    const int64_t T = srfe_spec_out_shape(p, n_samples, nullptr);
    if (T < 0) return (int)T;
    return run_device(FAM_SPEC, pcm, false, n_clips, n_samples, clip_stride, p, out, (cudaStream_t)stream, T);
}
int srfe_spec_host_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* p,
                       float* out, int device) {
    int rc = check_buffers(pcm, sizeof(float), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    int64_t s[2]; const int64_t T = srfe_spec_out_shape(p, n_samples, s);
    if (T < 0) return (int)T;
    return run_host(FAM_SPEC, pcm, false, n_clips, n_samples, clip_stride, p, out, device, T, s[0] * s[1]);
}
int srfe_spec_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* p,
                  float* out, void* stream) {
    int rc = check_buffers(pcm, sizeof(int16_t), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    const int64_t T = srfe_spec_out_shape(p, n_samples, nullptr);
    if (T < 0) return (int)T;
    return run_device(FAM_SPEC, pcm, true, n_clips, n_samples, clip_stride, p, out, (cudaStream_t)stream, T);
}
int srfe_spec_host_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_spec_params* p,
                       float* out, int device) {
    int rc = check_buffers(pcm, sizeof(int16_t), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    int64_t s[2]; const int64_t T = srfe_spec_out_shape(p, n_samples, s);
    if (T < 0) return (int)T;
    return run_host(FAM_SPEC, pcm, true, n_clips, n_samples, clip_stride, p, out, device, T, s[0] * s[1]);
}
int srfe_fbank_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_fbank_params* p,
                  float* out, void* stream) {
    int rc = check_buffers(pcm, sizeof(float), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    const int64_t T = srfe_fbank_out_shape(p, n_samples, nullptr);
    if (T < 0) return (int)T;
    return run_device(FAM_FBANK, pcm, false, n_clips, n_samples, clip_stride, p, out, (cudaStream_t)stream, T);
}
int srfe_fbank_host_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_fbank_params* p,
                       float* out, int device) {
    int rc = check_buffers(pcm, sizeof(float), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    int64_t s[2]; const int64_t T = srfe_fbank_out_shape(p, n_samples, s);
    if (T < 0) return (int)T;
    return run_host(FAM_FBANK, pcm, false, n_clips, n_samples, clip_stride, p, out, device, T, s[0] * s[1]);
}
int srfe_fbank_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_fbank_params* p,
                  float* out, void* stream) {
    int rc = check_buffers(pcm, sizeof(int16_t), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    const int64_t T = srfe_fbank_out_shape(p, n_samples, nullptr);
    if (T < 0) return (int)T;
    return run_device(FAM_FBANK, pcm, true, n_clips, n_samples, clip_stride, p, out, (cudaStream_t)stream, T);
}
int srfe_fbank_host_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_fbank_params* p,
                       float* out, int device) {
    int rc = check_buffers(pcm, sizeof(int16_t), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    int64_t s[2]; const int64_t T = srfe_fbank_out_shape(p, n_samples, s);
    if (T < 0) return (int)T;
    return run_host(FAM_FBANK, pcm, true, n_clips, n_samples, clip_stride, p, out, device, T, s[0] * s[1]);
}
int srfe_mfcc_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_mfcc_params* p,
                  float* out, void* stream) {
    int rc = check_buffers(pcm, sizeof(float), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    const int64_t T = srfe_mfcc_out_shape(p, n_samples, nullptr);
    if (T < 0) return (int)T;
    return run_device(FAM_MFCC, pcm, false, n_clips, n_samples, clip_stride, p, out, (cudaStream_t)stream, T);
}
int srfe_mfcc_host_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_mfcc_params* p,
                       float* out, int device) {
    int rc = check_buffers(pcm, sizeof(float), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    int64_t s[2]; const int64_t T = srfe_mfcc_out_shape(p, n_samples, s);
    if (T < 0) return (int)T;
    return run_host(FAM_MFCC, pcm, false, n_clips, n_samples, clip_stride, p, out, device, T, s[0] * s[1]);
}
int srfe_mfcc_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_mfcc_params* p,
                  float* out, void* stream) {
    int rc = check_buffers(pcm, sizeof(int16_t), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    const int64_t T = srfe_mfcc_out_shape(p, n_samples, nullptr);
    if (T < 0) return (int)T;
    return run_device(FAM_MFCC, pcm, true, n_clips, n_samples, clip_stride, p, out, (cudaStream_t)stream, T);
}
int srfe_mfcc_host_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const srfe_mfcc_params* p,
                       float* out, int device) {
    int rc = check_buffers(pcm, sizeof(int16_t), n_clips, n_samples, clip_stride, p, out);
    if (rc != SRFE_OK) return rc;
    int64_t s[2]; const int64_t T = srfe_mfcc_out_shape(p, n_samples, s);
    if (T < 0) return (int)T;
    return run_host(FAM_MFCC, pcm, true, n_clips, n_samples, clip_stride, p, out, device, T, s[0] * s[1]);
}

}  // extern "C"
