"""``sys.modules`` shim standing in for the absent third-party package librosa
(TEST INFRASTRUCTURE ONLY).

The reference imports ``from librosa.feature import mfcc``
(models/model_mfcc_bgru.py:5, models/model_mfrn_bgru.py:5) and
``from librosa.effects import pitch_shift`` (dataset.py:9).  librosa is not
installed, not vendored and not pinned by the reference, so the reference's MFCC
modules cannot be imported as they are.  ``install()`` registers a minimal module
whose ``feature.mfcc`` is the restated librosa-0.6 algorithm of
``oracle.features`` -- enough to import and run the *unmodified* reference model
files in this container (golden generation, drop-in tests).
"""
from __future__ import annotations

import sys
import types

import numpy as np

from .features import MfccParams, _mfcc_core


def _mfcc(y, sr=22050, S=None, n_mfcc=20, **kwargs):
    """librosa.feature.mfcc(y, sr, n_mfcc=, n_fft=, hop_length=) -> float64 [n_mfcc, frames]."""
    if S is not None:
        raise NotImplementedError("shim only supports the reference's call form")
    p = MfccParams(sr=int(sr), n_fft=int(kwargs.pop("n_fft", 2048)),
                   hop=int(kwargs.pop("hop_length", 512)),
                   n_mels=int(kwargs.pop("n_mels", 128)),
                   fmin=float(kwargs.pop("fmin", 0.0)), fmax=kwargs.pop("fmax", None),
                   n_mfcc=int(n_mfcc), n_deltas=0)
    if kwargs:
        raise TypeError(f"unsupported librosa.feature.mfcc arguments: {sorted(kwargs)}")
    return _mfcc_core(np.asarray(y), p, quantise_stft=True)


def _pitch_shift(*_a, **_k):
    raise NotImplementedError("librosa.effects.pitch_shift is outside the hot path (dataset.py:206-216)")


def install() -> types.ModuleType:
    if "librosa" in sys.modules and not getattr(sys.modules["librosa"], "__srfe_shim__", False):
        return sys.modules["librosa"]           # a real librosa is present: use it
    root = types.ModuleType("librosa")
    root.__srfe_shim__ = True
    root.__version__ = "0.6-restated-shim"
    feature = types.ModuleType("librosa.feature")
    feature.mfcc = _mfcc
    effects = types.ModuleType("librosa.effects")
    effects.pitch_shift = _pitch_shift
    root.feature, root.effects = feature, effects
    sys.modules["librosa"] = root
    sys.modules["librosa.feature"] = feature
    sys.modules["librosa.effects"] = effects
    return root
