"""CPU restatement of the reference feature front end (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to
/root/reference).  Two flavours per feature family:

* ``*_ref``   -- mirrors the reference's dtype path exactly (what the models see)
* ``*_truth`` -- the same algorithm carried in float64 throughout ("truth" for
                 judging an fp32 GPU pipeline without inheriting the oracle's own
                 single-precision noise)

Parameters that are literals in the reference are lifted to dataclass fields so
the BASELINE.json shapes ("C-presets") run through the same code.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
import scipy.signal

__all__ = [
    "SpecParams", "FbankParams", "MfccParams", "PRESETS",
    "R_SPEC", "C_SPEC", "R_FBANK", "C_FBANK", "R_MFCC", "C_MFCC", "C_MFCC_D2",
]


# --------------------------------------------------------------------------- #
# parameter sets
# --------------------------------------------------------------------------- #
@dataclass(frozen=True)
class SpecParams:
    """models/model_spec_bgru.py:13-14 literals, lifted."""
    fs: int = 16000
    nperseg: int = 640
    noverlap: int = 320
    log: bool = True            # np.log(S + eps)   (model_spec_bgru.py:14)
    eps: float = 1e-10
    layout: str = "ft"          # 'ft' = [freq, time] (spec_bgru), 'tf' = .T (spec_cnn:15)


@dataclass(frozen=True)
class FbankParams:
    """models/model_fbanks_cnn.py:18-45 literals, lifted."""
    fs: int = 16000
    frame_len: int = 400        # round(0.025 * fs)     (:22,27)
    frame_step: int = 160       # round(0.010 * fs)     (:23,28)
    nfft: int = 512             # (:41)
    preemph: float = 0.97       # (:19)
    nfilt: int = 120            # (:45)
    vtlp_alpha: float = 0.0     # 0 = off; legacy/model_8/dataset_top.py:251-252 (alpha ~ U(0.9, 1.1) there)


@dataclass(frozen=True)
class MfccParams:
    """models/model_mfcc_bgru.py:13-16 call + librosa-0.6 defaults, lifted."""
    sr: int = 16000
    n_fft: int = 640
    win_length: Optional[int] = None    # librosa: defaults to n_fft
    hop: int = 320
    n_mels: int = 128                   # librosa.feature.melspectrogram default
    fmin: float = 0.0
    fmax: Optional[float] = None        # librosa: sr / 2
    n_mfcc: int = 13
    n_deltas: int = 2                   # np.gradient applied 0, 1 or 2 times (:14-16)
    top_db: float = 80.0                # librosa.power_to_db default
    amin: float = 1e-10

    @property
    def win(self) -> int:
        return self.n_fft if self.win_length is None else self.win_length

    @property
    def f_hi(self) -> float:
        return self.sr / 2.0 if self.fmax is None else self.fmax


R_SPEC = SpecParams()
C_SPEC = SpecParams(nperseg=512, noverlap=256)
R_FBANK = FbankParams()
C_FBANK = FbankParams(nfilt=40)
R_MFCC = MfccParams()
C_MFCC = MfccParams(n_fft=512, win_length=400, hop=160, n_mfcc=40, n_deltas=0)
C_MFCC_D2 = replace(C_MFCC, n_deltas=2)

PRESETS = {
    "R-SPEC": R_SPEC, "C-SPEC": C_SPEC,
    "R-FBANK": R_FBANK, "C-FBANK": C_FBANK,
    "R-MFCC": R_MFCC, "C-MFCC": C_MFCC, "C-MFCC-D2": C_MFCC_D2,
}


# --------------------------------------------------------------------------- #
# windows (closed forms; checked against scipy in tests/test_oracle.py)
# --------------------------------------------------------------------------- #
def tukey_periodic(n: int, alpha: float = 0.25) -> np.ndarray:
    """scipy.signal.get_window(('tukey', alpha), n) == tukey(n+1, sym=True)[:-1]
    (the default window of scipy.signal.spectrogram, model_spec_bgru.py:13)."""
    m = n + 1                                   # symmetric length
    k = np.arange(n, dtype=np.float64)
    width = int(np.floor(alpha * (m - 1) / 2.0))
    w = np.ones(n, dtype=np.float64)
    a = k[: width + 1]
    w[: width + 1] = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * a / alpha / (m - 1))))
    c = k[m - width - 1:]
    w[m - width - 1:] = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * c / alpha / (m - 1))))
    return w


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) (librosa.stft default)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def hamming_symmetric(n: int) -> np.ndarray:
    """np.hamming(n) (model_fbanks_cnn.py:40)."""
    if n == 1:
        return np.ones(1)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / (n - 1))


# --------------------------------------------------------------------------- #
# frame counts
# --------------------------------------------------------------------------- #
def spec_num_frames(n: int, p: SpecParams = R_SPEC) -> int:
    """scipy _fft_helper: (n - noverlap) // (nperseg - noverlap); no padding."""
    return (n - p.noverlap) // (p.nperseg - p.noverlap)


def fbank_num_frames(n: int, p: FbankParams = R_FBANK) -> int:
    """model_fbanks_cnn.py:29."""
    return int(np.ceil(float(np.abs(n - p.frame_len)) / p.frame_step))


def mfcc_num_frames(n: int, p: MfccParams = R_MFCC) -> int:
    """librosa.stft(center=True): 1 + n // hop."""
    return 1 + n // p.hop


# --------------------------------------------------------------------------- #
# R-SPEC / C-SPEC
# --------------------------------------------------------------------------- #
def spec_ref(x: np.ndarray, p: SpecParams = R_SPEC) -> np.ndarray:
    """Mirror of compute_spec (models/model_spec_bgru.py:11-17; spec_cnn.py:12-18
    adds .T).  The arithmetic is scipy.signal.spectrogram's, called exactly like
    the reference does; for float32 input scipy stays in single precision."""
    _, _, s = scipy.signal.spectrogram(x, fs=p.fs, nperseg=p.nperseg,
                                       noverlap=p.noverlap, detrend=False)
    if p.log:
        s = np.log(s.astype(np.float32) + p.eps)        # :14 (float32 + python float)
    s = s.astype(np.float32)
    return np.ascontiguousarray(s.T) if p.layout == "tf" else s


def spec_truth(x: np.ndarray, p: SpecParams = R_SPEC) -> np.ndarray:
    """float64 restatement of scipy's _spectral_helper for the reference call
    (periodic Tukey-0.25, density scaling, one-sided doubling, no detrend,
    no padding)."""
    w = tukey_periodic(p.nperseg, 0.25)
    step = p.nperseg - p.noverlap
    fr = np.lib.stride_tricks.sliding_window_view(
        np.asarray(x, dtype=np.float64), p.nperseg)[::step]
    X = np.fft.rfft(fr * w, p.nperseg, axis=1)
    P = (X.real ** 2 + X.imag ** 2) / (p.fs * (w * w).sum())
    if p.nperseg % 2 == 0:
        P[:, 1:-1] *= 2.0
    else:
        P[:, 1:] *= 2.0
    if p.log:
        P = np.log(P + p.eps)
    P = P.T                                              # [freq, time]
    return np.ascontiguousarray(P.T) if p.layout == "tf" else np.ascontiguousarray(P)


# --------------------------------------------------------------------------- #
# R-FBANK / C-FBANK
# --------------------------------------------------------------------------- #
def htk_floor_filterbank(p: FbankParams = R_FBANK) -> np.ndarray:
    """The nfilt x (nfft/2+1) triangle matrix of model_fbanks_cnn.py:46-59
    (HTK mel, bin = floor((NFFT+1)*hz/fs), left-closed half-open edges)."""
    high = 2595.0 * np.log10(1.0 + (p.fs / 2.0) / 700.0)            # :47
    mel_points = np.linspace(0.0, high, p.nfilt + 2)                 # :48
    hz_points = 700.0 * (10.0 ** (mel_points / 2595.0) - 1.0)        # :49
    if p.vtlp_alpha:                                                 # legacy/model_8/dataset_top.py:251-252, 8000 -> fs/2
        a, nyq = float(np.float32(p.vtlp_alpha)), p.fs / 2.0
        m1 = min(a, 1.0)
        hz_points = np.array([h * a if h < (4800 * m1 / a) else nyq - ((nyq - 4800 * m1) / (nyq - 4800 * (m1 / a))) * (nyq - h)
                              for h in hz_points])
    bins = np.floor((p.nfft + 1) * hz_points / p.fs)                 # :50
    fb = np.zeros((p.nfilt, p.nfft // 2 + 1))                        # :51
    for m in range(1, p.nfilt + 1):                                  # :52
        lo, ce, hi = int(bins[m - 1]), int(bins[m]), int(bins[m + 1])
        k = np.arange(lo, ce)
        fb[m - 1, lo:ce] = (k - bins[m - 1]) / (bins[m] - bins[m - 1]) if ce > lo else 0  # :57
        k = np.arange(ce, hi)
        fb[m - 1, ce:hi] = (bins[m + 1] - k) / (bins[m + 1] - bins[m]) if hi > ce else 0  # :59
    return fb


def _fbank_core(emph: np.ndarray, p: FbankParams) -> np.ndarray:
    n = emph.shape[0]
    nfr = fbank_num_frames(n, p)                                      # :29
    pad_len = nfr * p.frame_step + p.frame_len                        # :31
    padded = np.append(emph, np.zeros(pad_len - n))                   # :32-33 (promotes to float64)
    idx = (np.arange(p.frame_len)[None, :]
           + np.arange(0, nfr * p.frame_step, p.frame_step)[:, None])  # :35-36
    frames = padded[idx] * hamming_symmetric(p.frame_len)             # :37,40
    mag = np.absolute(np.fft.rfft(frames, p.nfft))                    # :42
    powf = (1.0 / p.nfft) * (mag ** 2)                                # :43
    fb = np.dot(powf, htk_floor_filterbank(p).T)                      # :60
    fb = np.where(fb == 0, np.finfo(float).eps, fb)                   # :61
    return 20.0 * np.log10(fb)                                        # :62


def fbank_ref(x: np.ndarray, p: FbankParams = R_FBANK) -> np.ndarray:
    """Mirror of filter_banks (models/model_fbanks_cnn.py:15-66): pre-emphasis in
    the input dtype (float32, with fl32(0.97)), everything after in float64,
    final cast float32.  Output [frames, nfilt]."""
    x = np.asarray(x)
    emph = np.append(x[0], x[1:] - p.preemph * x[:-1])                # :20
    return _fbank_core(emph, p).astype(np.float32)                    # :64-65


def fbank_truth(x: np.ndarray, p: FbankParams = R_FBANK) -> np.ndarray:
    """Same algorithm, float64 from the framing on, but with the reference's
    float32-rounded pre-emphasis kept (it is part of the reference's semantics:
    the float32 samples e[n] ARE the signal that gets analysed)."""
    x = np.asarray(x)
    emph = np.append(x[0], x[1:] - p.preemph * x[:-1]).astype(np.float64)
    return _fbank_core(emph, p)


# --------------------------------------------------------------------------- #
# R-MFCC / C-MFCC   (librosa 0.6.x restated -- third-party, un-vendored, un-pinned)
# --------------------------------------------------------------------------- #
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3.0)
    logstep = np.log(6.4) / 27.0
    return np.where(f >= 1000.0,
                    15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / logstep, lin)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), (200.0 / 3.0) * m)


def slaney_mel_filterbank(sr=16000, n_fft=640, n_mels=128, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1):
    Slaney mel scale, triangles in Hz, area normalisation.  [n_mels, 1+n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    return w * (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]


def dct2_ortho_matrix(n_out: int, n_in: int) -> np.ndarray:
    """librosa.filters.dct(n_out, n_in) == first rows of the orthonormal DCT-II."""
    k = np.arange(n_out, dtype=np.float64)[:, None]
    n = np.arange(n_in, dtype=np.float64)[None, :]
    d = np.cos(np.pi * k * (2.0 * n + 1.0) / (2.0 * n_in)) * np.sqrt(2.0 / n_in)
    d[0] = 1.0 / np.sqrt(n_in)
    return d


def _mfcc_window(p: MfccParams) -> np.ndarray:
    """periodic Hann of win_length, zero-padded centred to n_fft (librosa.util.pad_center)."""
    w = hann_periodic(p.win)
    lpad = (p.n_fft - p.win) // 2
    out = np.zeros(p.n_fft)
    out[lpad:lpad + p.win] = w
    return out


def _mfcc_core(y: np.ndarray, p: MfccParams, quantise_stft: bool) -> np.ndarray:
    yp = np.pad(y, p.n_fft // 2, mode="reflect")                      # librosa.stft(center=True)
    fr = np.lib.stride_tricks.sliding_window_view(yp, p.n_fft)[:: p.hop]
    X = np.fft.rfft(_mfcc_window(p) * fr, p.n_fft, axis=1).T          # [1+n_fft/2, frames]
    if quantise_stft:
        X = X.astype(np.complex64)                                    # librosa stores STFT as complex64
        S = (np.abs(X) ** 2)                                          # float32
    else:
        S = X.real ** 2 + X.imag ** 2
    M = slaney_mel_filterbank(p.sr, p.n_fft, p.n_mels, p.fmin, p.f_hi) @ S.astype(np.float64)
    dB = 10.0 * np.log10(np.maximum(p.amin, M))                       # power_to_db, ref=1.0
    if p.top_db is not None:
        dB = np.maximum(dB, dB.max() - p.top_db)                      # max over the WHOLE clip
    C = dct2_ortho_matrix(p.n_mfcc, p.n_mels) @ dB                    # [n_mfcc, frames]
    out = [C]
    g = C
    for _ in range(p.n_deltas):                                       # model_mfcc_bgru.py:14-16
        g = np.gradient(g, axis=1)
        out.append(g)
    return np.concatenate(out)


def mfcc_ref(y: np.ndarray, p: MfccParams = R_MFCC) -> np.ndarray:
    """compute_mfcc (models/model_mfcc_bgru.py:11-19) with librosa-0.6 internals
    restated: float32 signal, float64 FFT, complex64 STFT storage, float64
    mel/log/DCT, float32 result.  Output [(1+n_deltas)*n_mfcc, frames]."""
    return _mfcc_core(np.asarray(y), p, quantise_stft=True).astype(np.float32)


def mfcc_truth(y: np.ndarray, p: MfccParams = R_MFCC) -> np.ndarray:
    """Same chain in float64 throughout (no complex64 quantisation)."""
    return _mfcc_core(np.asarray(y, dtype=np.float64), p, quantise_stft=False)
