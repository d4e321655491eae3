"""Independent implementations of the librosa MFCC chain that ship in this image (TEST INFRASTRUCTURE ONLY):
``transformers.audio_utils`` + ``scipy.fft.dct`` and ``torchaudio.transforms.MFCC``.  Used by
tests/test_mfcc_crosscheck.py (live) and oracle/make_golden.py (recorded report) to pin the restatement of the
un-vendored third-party arithmetic behind models/model_mfcc_bgru.py:13."""
from __future__ import annotations

import numpy as np
import scipy.fft

from .features import MfccParams


def hf_mel_matrix(p: MfccParams) -> np.ndarray:
    from transformers import audio_utils as au
    return au.mel_filter_bank(num_frequency_bins=p.n_fft // 2 + 1, num_mel_filters=p.n_mels, min_frequency=p.fmin,
                              max_frequency=p.f_hi, sampling_rate=p.sr, norm="slaney", mel_scale="slaney")


def hf_mfcc(x64: np.ndarray, p: MfccParams) -> np.ndarray:
    """static coefficients [n_mfcc, frames], float64 (HF stores its STFT as complex64)"""
    from transformers import audio_utils as au
    fb = hf_mel_matrix(p)
    # HF frames win_length samples and zero-pads them at the END of the FFT buffer; librosa centres the window in
    # n_fft.  Same samples, same window values, a linear phase apart: identical power spectra.
    s = au.spectrogram(x64, au.window_function(p.win, "hann"), frame_length=p.win, hop_length=p.hop,
                       fft_length=p.n_fft, power=2.0, center=True, pad_mode="reflect", mel_filters=fb,
                       mel_floor=p.amin, log_mel="dB", reference=1.0, min_value=p.amin, db_range=p.top_db)
    return scipy.fft.dct(s, axis=0, type=2, norm="ortho")[: p.n_mfcc]


def torchaudio_mfcc(x64: np.ndarray, p: MfccParams) -> np.ndarray:
    """static coefficients [n_mfcc, frames], double-precision transform over float32-built mel / DCT tables"""
    import torch
    import torchaudio as ta
    t = ta.transforms.MFCC(p.sr, p.n_mfcc, log_mels=False, melkwargs=dict(
        n_fft=p.n_fft, win_length=p.win, hop_length=p.hop, n_mels=p.n_mels, f_min=p.fmin, f_max=p.f_hi,
        norm="slaney", mel_scale="slaney", center=True, pad_mode="reflect", power=2.0)).double()
    return t(torch.from_numpy(x64)).numpy()            # unbatched: top_db is relative to this clip's maximum
