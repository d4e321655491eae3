/*
 * srfe.h -- C ABI of the B200-native speech-recognition feature front end.
 *
 * Drop-in boundary for the acoustic feature hot path of
 * remit0/SpeechRecognitionProject.  The reference has no FFI of its own: the
 * boundary is five Python functions and the per-clip loops that call them
 * (citations are into /root/reference):
 *
 *   srfe_spec_f32   replaces  compute_spec   models/model_spec_bgru.py:11-17
 *                                            models/model_spec_cnn.py:12-18 (layout TF)
 *                   + loops   models/model_spec_bgru.py:27-32, model_spec_cnn.py:38-43
 *   srfe_fbank_f32  replaces  filter_banks   models/model_fbanks_cnn.py:15-66
 *                   + loop    models/model_fbanks_cnn.py:84-88
 *   srfe_mfcc_f32   replaces  compute_mfcc   models/model_mfcc_bgru.py:11-19
 *                                            models/model_mfrn_bgru.py:11-19
 *                   + loops   models/model_mfcc_bgru.py:29-34, model_mfrn_bgru.py:128-133
 *   srfe_*_host_f32 the same call with HOST buffers (what Network.forward holds
 *                   today: a CPU float32 [B,16000] batch, training.py:86), doing
 *                   H2D -> kernel -> D2H inside the call.
 *
 * Conventions
 *   - plain C, POD structs, no C++/torch types; every function returns 0
 *     (SRFE_OK) or a negative srfe_status; srfe_last_error_string() gives the
 *     thread-local detail.  There is NO CPU fallback: without a CUDA device the
 *     compute entry points return SRFE_ERR_NO_DEVICE / SRFE_ERR_CUDA.
 *   - device entry points are asynchronous and stream-ordered on `cuda_stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream); they never
 *     synchronise and never allocate on the hot call once the (device, params)
 *     tables exist (first call per parameter set builds and uploads them).
 *   - the caller owns `pcm` and `out`; the library owns only immutable tables.
 *   - pcm: float32, `n_clips` rows of `n_samples`, row stride `clip_stride`
 *     elements (>= n_samples), int16-scale values as dataset.py:117 produces.
 *     Rows must be 8-byte aligned (clip_stride even, base 8-byte aligned).  The *_i16 variants take the same
 *     layout with int16 samples (rows 4-byte aligned).
 *   - out: contiguous float32, shape given by srfe_*_out_shape().
 */
#ifndef SRFE_H_
#define SRFE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRFE_VERSION_MAJOR 0
#define SRFE_VERSION_MINOR 1

typedef enum srfe_status {
    SRFE_OK = 0,
    SRFE_ERR_BAD_ARG = -1,       /* null pointer, negative size, misaligned pointer/stride */
    SRFE_ERR_UNSUPPORTED = -2,   /* parameter combination the kernels do not implement */
    SRFE_ERR_CUDA = -3,          /* CUDA runtime / launch error (see error string) */
    SRFE_ERR_NO_DEVICE = -4,     /* no CUDA device visible */
    SRFE_ERR_TOO_LARGE = -5      /* clip too long for the fused per-clip kernels */
} srfe_status;

typedef enum srfe_layout {
    SRFE_LAYOUT_FT = 0,          /* [clip][feature/frequency][time]  (spec_bgru, mfcc) */
    SRFE_LAYOUT_TF = 1           /* [clip][time][feature/frequency]  (spec_cnn, fbank; GRU-ready) */
} srfe_layout;

/* log-spectrogram: scipy.signal.spectrogram(fs, nperseg, noverlap, detrend=False)
 * with its defaults (periodic Tukey(0.25), density scaling, one-sided), then
 * optionally ln(S + log_eps).  model_spec_bgru.py:13-14. */
typedef struct srfe_spec_params {
    int32_t sample_rate;         /* 16000 */
    int32_t nperseg;             /* 640 (reference) or 512; == n_fft */
    int32_t noverlap;            /* 320 (reference) */
    int32_t take_log;            /* 1: ln(S + log_eps); 0: raw PSD */
    float   log_eps;             /* 1e-10 */
    int32_t layout;              /* srfe_layout; FT -> [B, nperseg/2+1, T], TF -> [B, T, nperseg/2+1] */
} srfe_spec_params;

/* log mel filterbank energies, model_fbanks_cnn.py:15-66 (python_speech_features
 * style): float32 pre-emphasis, Hamming(frame_len), rfft(n_fft), |X|^2/n_fft,
 * HTK mel triangles on floor((n_fft+1) f / fs) bins, 0 -> DBL_EPSILON, 20 log10.
 * Output always TF: [B, T, nfilt]. */
typedef struct srfe_fbank_params {
    int32_t sample_rate;         /* 16000 */
    int32_t frame_len;           /* 400 */
    int32_t frame_step;          /* 160 */
    int32_t n_fft;               /* 512 (or 640) ; frame_len <= n_fft */
    float   preemph;             /* 0.97 */
    int32_t nfilt;               /* 120 (reference), 40 (BASELINE cfg2) */
    float   vtlp_alpha;          /* 0 = off.  Vocal-tract-length perturbation of the filter centre frequencies before they
                                  * are floored to bins (legacy/model_8/dataset_top.py:251-252; alpha ~ U(0.9, 1.1) there):
                                  * hz < 4800 min(alpha,1)/alpha -> hz alpha, else mapped linearly onto the rest up to
                                  * fs/2.  One alpha per call: the table is cached per (device, parameter set). */
} srfe_fbank_params;

/* MFCC + deltas, model_mfcc_bgru.py:13-16 with librosa-0.6 semantics:
 * stft(center, reflect pad, periodic Hann(win_length) centred in n_fft), |X|^2,
 * Slaney mel (area-normalised), 10 log10(max(amin, .)), clamp to clip max - top_db,
 * orthonormal DCT-II, np.gradient applied n_deltas times along time.
 * Output FT: [B, (1+n_deltas)*n_mfcc, T]  or TF: [B, T, (1+n_deltas)*n_mfcc]. */
typedef struct srfe_mfcc_params {
    int32_t sample_rate;         /* 16000 */
    int32_t n_fft;               /* 640 (reference) or 512 */
    int32_t win_length;          /* <= n_fft; 0 means n_fft */
    int32_t hop;                 /* 320 (reference) */
    int32_t n_mels;              /* 128 */
    float   fmin;                /* 0 */
    float   fmax;                /* <= 0 means sample_rate / 2 */
    int32_t n_mfcc;              /* 13 (reference), 40 (BASELINE cfg1) */
    int32_t n_deltas;            /* 0, 1 or 2 (reference: 2) */
    float   top_db;              /* 80; < 0 disables the clamp */
    float   amin;                /* 1e-10 */
    int32_t layout;              /* srfe_layout */
} srfe_mfcc_params;

/* ---- library / device ---------------------------------------------------- */
int         srfe_version(void);                 /* major*1000 + minor */
const char* srfe_last_error_string(void);       /* thread-local, never NULL */
int         srfe_device_count(void);            /* >= 0, or negative srfe_status */

/* ---- shapes (host only, no device needed) --------------------------------- */
/* Writes {rows, cols} of one clip's feature matrix in the params' layout and
 * returns the number of frames T (>= 0), or a negative srfe_status. */
int64_t srfe_spec_out_shape (const srfe_spec_params*  p, int64_t n_samples, int64_t shape2[2]);
int64_t srfe_fbank_out_shape(const srfe_fbank_params* p, int64_t n_samples, int64_t shape2[2]);
int64_t srfe_mfcc_out_shape (const srfe_mfcc_params*  p, int64_t n_samples, int64_t shape2[2]);

/* ---- host-side table builders (double precision; used by the kernels' table
 *      cache and exposed so they can be verified without a GPU) -------------- */
/* window of length n_fft as the kernels apply it (zero outside the analysis window) */
int srfe_spec_window_f64 (const srfe_spec_params*  p, double* win /* [nperseg] */);
int srfe_fbank_window_f64(const srfe_fbank_params* p, double* win /* [n_fft]   */);
int srfe_mfcc_window_f64 (const srfe_mfcc_params*  p, double* win /* [n_fft]   */);
/* dense filter matrices, row-major [n_filters][n_fft/2+1] */
int srfe_fbank_filters_f64(const srfe_fbank_params* p, double* w);
int srfe_mfcc_filters_f64 (const srfe_mfcc_params*  p, double* w);
/* orthonormal DCT-II rows, row-major [n_mfcc][n_mels] */
int srfe_mfcc_dct_f64(const srfe_mfcc_params* p, double* d);

/* ---- device entry points (pcm/out are DEVICE pointers) -------------------- */
int srfe_spec_f32 (const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   const srfe_spec_params* p, float* out, void* cuda_stream);
int srfe_fbank_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   const srfe_fbank_params* p, float* out, void* cuda_stream);
int srfe_mfcc_f32 (const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   const srfe_mfcc_params* p, float* out, void* cuda_stream);

/* int16 ingest: the wav's native sample type (dataset.py:103), converted on load inside the
 * kernel -- replaces dataset.py:117's astype(float32) for callers that keep PCM as int16 and halves the bytes
 * that cross PCIe / HBM.  Integer-valued float32 input and its int16 copy give bit-identical features.
 * Rows must be 4-byte aligned (clip_stride even). */
int srfe_spec_i16 (const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   const srfe_spec_params* p, float* out, void* cuda_stream);
int srfe_fbank_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   const srfe_fbank_params* p, float* out, void* cuda_stream);
int srfe_mfcc_i16 (const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                   const srfe_mfcc_params* p, float* out, void* cuda_stream);

/* ---- several feature sets of the same batch in ONE launch ---- */
/* The reference's ensemble pushes one batch through model_spec_* and model_fbanks_cnn in turn
 * (analyst_training.py:91-94, predictions.py:58-60), each recomputing its features from the PCM.  These entry points
 * produce the spectrogram AND the log-fbank features of the same clips with one kernel: a CTA takes a group of clips
 * through both frame loops back to back, so HBM delivers every sample once for both outputs.  Results are bit-identical
 * to srfe_spec_* and srfe_fbank_* called separately.  The fbank side needs n_fft = 512 (else SRFE_ERR_UNSUPPORTED). */
int srfe_spec_fbank_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_spec_params* ps, const srfe_fbank_params* pf, float* out_spec, float* out_fbank,
                        void* cuda_stream);
int srfe_spec_fbank_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_spec_params* ps, const srfe_fbank_params* pf, float* out_spec, float* out_fbank,
                        void* cuda_stream);

/* ---- host entry points (pcm/out are HOST pointers; synchronous) ----------- */
/* H2D of the batch, the same kernels, D2H of the features, on `device`, chunked over two streams so that the
 * transfers of one chunk overlap the kernel of the other.  Pinned (page-locked) caller buffers are copied from / to
 * directly; PAGEABLE buffers -- what the reference's DataLoader hands over, training.py:77 has no pin_memory -- are
 * gathered by a few host threads into a pinned staging ring first (and the features scattered back from one).  The
 * workspace (2 streams, device in / out buffers, staging buffers when needed) is per device, grows on demand, is
 * guarded by a per-device mutex and is freed by srfe_release_host_workspace(). */
int srfe_spec_host_f32 (const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_spec_params* p, float* out, int device);
int srfe_fbank_host_f32(const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_fbank_params* p, float* out, int device);
int srfe_mfcc_host_f32 (const float* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_mfcc_params* p, float* out, int device);

int srfe_spec_host_i16 (const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_spec_params* p, float* out, int device);
int srfe_fbank_host_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_fbank_params* p, float* out, int device);
int srfe_mfcc_host_i16 (const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride,
                        const srfe_mfcc_params* p, float* out, int device);

/* ---- on-device augmentation + silence synthesis -------------- */
/* What the reference's Dataset does per clip on the host with global random state -- generate_silence_sample
 * (dataset.py:148-161), add_noise_uniform (:185-191), add_noise_snr (:163-183), time_stretching (:193-202) and the band
 * selection of __getitem__ (:107-116) -- as ONE launch over a batch of int16 clips that stays on the device and feeds the
 * front end on the same stream.  Randomness is a counter-based contract: Philox4x32-10, key = seed, counter = (global clip
 * index = first_clip_index + row, draw block); the draw layout is in oracle/augment.py, whose numpy restatement this
 * kernel matches bit for bit.  A clip takes the op whose band [lo, hi) contains its uniform draw; the reference's pitch-shift
 * and speed-tune bands (librosa / cv2 resampling) are host-only: such clips are copied through and flagged. */
typedef enum srfe_augment_op {
    SRFE_AUG_NONE = 0, SRFE_AUG_SHIFT = 1, SRFE_AUG_NOISE_UNIFORM = 2, SRFE_AUG_NOISE_SNR = 3,
    SRFE_AUG_SILENCE_ZERO = 4, SRFE_AUG_SILENCE_NOISE = 5, SRFE_AUG_HOST_PITCH = 6, SRFE_AUG_HOST_SPEED = 7
} srfe_augment_op;

typedef struct srfe_augment_params {
    uint64_t seed;
    float shift_lo, shift_hi;    /* time shift band          (dataset.py:112-113: 0.4, 0.6) */
    float noise_lo, noise_hi;    /* add_noise_uniform band   (dataset.py:114-115: 0.6, 0.8) */
    float snr_lo, snr_hi;        /* add_noise_snr band       (never called by __getitem__: 0, 0 = off) */
    float pitch_lo, pitch_hi;    /* host-only bands, reported in op_out (0, 0.2) / (0.2, 0.4) */
    float speed_lo, speed_hi;
    int32_t shift_range;         /* 4800 */
    float noise_upper;           /* 0.1 */
} srfe_augment_params;

/* pcm: DEVICE int16 [n_clips][clip_stride]; kind: DEVICE int8 per clip (0 = clip, 1 = silence of zeros, 2 = silence from a
 * noise slice) or NULL; noise_bank: DEVICE int16, the background files back to back; noise_offsets: DEVICE int64
 * [n_noise_files + 1] (every file at least n_samples long); out: DEVICE float32 [n_clips][n_samples] (dataset.py:117);
 * op_out: DEVICE int8 [n_clips] (srfe_augment_op) or NULL. */
int srfe_augment_i16(const int16_t* pcm, int64_t n_clips, int64_t n_samples, int64_t clip_stride, const int8_t* kind,
                     int64_t first_clip_index, const int16_t* noise_bank, const int64_t* noise_offsets, int32_t n_noise_files,
                     const srfe_augment_params* p, float* out, int8_t* op_out, void* cuda_stream);

/* Host rows -> a DEVICE buffer the caller owns, stream-ordered on `cuda_stream` of `device`: the PCM upload of the
 * reference's forward (`x.to(DEVICE)` of a DataLoader batch, pageable: training.py:77) through libsrfe's pinned staging ring
 * and copy threads instead of the driver's single staged copy; pinned sources are copied directly.  `dev` receives the rows
 * densely packed (row_bytes apart).  Returns once the host buffer may be reused. */
int srfe_upload(const void* host, int64_t rows, int64_t row_bytes, int64_t row_stride_bytes, void* dev, int device,
                void* cuda_stream);

/* frees the host entry points' per-device workspaces (streams, device buffers, pinned staging); they are re-created on
 * the next host call.  Returns SRFE_OK. */
int srfe_release_host_workspace(void);

/* ---- introspection for benchmarks / tests --------------------------------- */
/* Launch-shape override for tests and tuning sweeps: `name` in {"warps", "ctas", "cpc", "dct_cb", "dct_pq",
 * "mfcc_tc", "stage", "fbank_tc"}, value 0 = automatic (the default).  Results never depend on these beyond the stated
 * tolerance (the parity suite checks it); "mfcc_tc" / "fbank_tc": 1 = force the classic CUDA-core kernel of the family,
 * 2 = require the tcgen05 one (SRFE_ERR_UNSUPPORTED when the parameter set does not fit it; the FBANK one is opt-in only:
 * it ties the classic kernel and is not bit-identical to it); "stage": 2 = TMA-staged spectrogram frames.
 * Process-wide, thread-safe. */
int srfe_set_tuning(const char* name, int value);
/* number of kernel launches issued by this process through the entry points */
int64_t srfe_launch_count(void);
/* algorithmic bytes per clip (fp32 PCM in + fp32 features out), SURVEY.md 8d */
int64_t srfe_spec_bytes_per_clip (const srfe_spec_params*  p, int64_t n_samples);
int64_t srfe_fbank_bytes_per_clip(const srfe_fbank_params* p, int64_t n_samples);
int64_t srfe_mfcc_bytes_per_clip (const srfe_mfcc_params*  p, int64_t n_samples);

/* device (global-memory) workspace the device entry points need for this call: always 0 -- PCM to features runs in
 * registers and shared memory; returns a negative srfe_status for invalid parameters.  (Host entry points keep a
 * per-device workspace of their own, see above.) */
int64_t srfe_spec_workspace_bytes (const srfe_spec_params*  p, int64_t n_clips, int64_t n_samples);
int64_t srfe_fbank_workspace_bytes(const srfe_fbank_params* p, int64_t n_clips, int64_t n_samples);
int64_t srfe_mfcc_workspace_bytes (const srfe_mfcc_params*  p, int64_t n_clips, int64_t n_samples);

#ifdef __cplusplus
}
#endif
#endif /* SRFE_H_ */
