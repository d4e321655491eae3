// srfe_tables.h -- host-side (double precision) builders for the immutable tables
// the kernels consume: analysis windows, FFT twiddles, sparse mel filterbanks,
// DCT-II rows.  Pure C++, no CUDA: also compiled into tests/emu.
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/srfe.h"

namespace srfe {

struct F2 { float x, y; };

// --- frame counts (reference formulas) -------------------------------------------
int64_t spec_frames(const srfe_spec_params& p, int64_t n);    // scipy: (n - noverlap) // (nperseg - noverlap)
int64_t fbank_frames(const srfe_fbank_params& p, int64_t n);  // model_fbanks_cnn.py:29
int64_t mfcc_frames(const srfe_mfcc_params& p, int64_t n);    // librosa center=True: 1 + n // hop

// --- validation: 0 or a negative srfe_status, message in *why ----------------------
int validate(const srfe_spec_params& p, const char** why);
int validate(const srfe_fbank_params& p, const char** why);
int validate(const srfe_mfcc_params& p, const char** why);

// --- windows, length n_fft, zero outside the analysis window -----------------------
void window_tukey_periodic(int n, double alpha, std::vector<double>& w);          // scipy get_window(('tukey',a), n)
void spec_window(const srfe_spec_params& p, std::vector<double>& w);              // [nperseg]
void fbank_window(const srfe_fbank_params& p, std::vector<double>& w);            // hamming(frame_len) then zeros
void mfcc_window(const srfe_mfcc_params& p, std::vector<double>& w);              // hann(win) centred in n_fft

// --- dense filter matrices [n_filters][n_fft/2+1] -----------------------------------
void fbank_filters(const srfe_fbank_params& p, std::vector<double>& w);           // model_fbanks_cnn.py:46-59
void mfcc_filters(const srfe_mfcc_params& p, std::vector<double>& w);             // librosa.filters.mel (slaney, norm=1)
void mfcc_dct(const srfe_mfcc_params& p, std::vector<double>& d);                 // [n_mfcc][n_mels]

// --- sparse (CSR) form of a filter matrix, weights pre-scaled ------------------------
struct SparseBank {
    std::vector<int32_t> start, count, offset;   // per filter: first bin, #bins, index into weight[]
    std::vector<float> weight;
    int max_count = 0;
};
void to_sparse(const std::vector<double>& dense, int n_filters, int n_bins, double scale, SparseBank& out);

// --- ELL form for the kernels: filters in groups of 16 (one per lane), every group padded to a common
//     multiple-of-4 run length so the inner loop has a uniform trip count and reads float4 weights
struct EllBank {
    std::vector<int32_t> gmeta;      // per group: {offset into w4 (float4 units, /16 lanes), n4 = run length / 4}
    std::vector<int32_t> start;      // per filter (padded to 16 * groups): first bin read (bank-skewed, <= first nonzero bin)
    std::vector<float> w4;           // [((offset + q4) * 2 + plane) * 16 + lane] float2: weights 2 plane, 2 plane + 1 of step q4
    int groups = 0;
    int max_reach = 0;               // max over filters of start + 4 * n4 (bins read, incl. padding)
};
void to_ell(const SparseBank& sb, EllBank& out);

// --- FFT twiddles (see srfe_fft.cuh) ---------------------------------------------------
void fft_twiddles(int n_fft, std::vector<F2>& tw1, std::vector<F2>& twu, std::vector<F2>& tw16);

}  // namespace srfe
