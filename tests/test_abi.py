"""CPU: the C-ABI shared library loads, exports every symbol include/srfe.h declares, and its
host-side logic (shapes, validation, table builders) matches the oracle -- no compute calls."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import speechrecognitionproject_b200 as S
from speechrecognitionproject_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = C.POINTER(C.c_double)


def _tab(fn, cp, n):
    a = np.zeros(n)
    _lib.check(fn(C.byref(cp), a.ctypes.data_as(DP)))
    return a


def test_header_symbols_all_exported(srfe_lib):
    hdr = open(os.path.join(ROOT, "include", "srfe.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(srfe_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(srfe_lib, name), f"{name} declared in include/srfe.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert srfe_lib.srfe_version() == 1


def test_tuning_knobs_agree_between_header_python_and_library(srfe_lib):
    """Every knob the Python mirror knows is documented in include/srfe.h and accepted by srfe_set_tuning (host-side only:
    the hook stores an integer); unknown names are refused."""
    from speechrecognitionproject_b200 import features as F
    hdr = open(os.path.join(ROOT, "include", "srfe.h")).read()
    doc = hdr[hdr.index("Launch-shape override"):hdr.index("int srfe_set_tuning")]
    for k in F._TUNING_KNOBS:
        assert f'"{k}"' in doc, f"tuning knob {k} is not documented in include/srfe.h"
        assert srfe_lib.srfe_set_tuning(k.encode(), 0) == 0
    assert srfe_lib.srfe_set_tuning(b"no_such_knob", 1) != 0


def test_out_shapes_and_bytes(srfe_lib):
    expect = {"R-SPEC": ((321, 49), 126916), "C-SPEC": ((257, 61), 126708), "R-FBANK": ((98, 120), 111040),
              "C-FBANK": ((98, 40), 79680), "R-MFCC": ((39, 51), 71956), "C-MFCC": ((40, 101), 80160)}
    for name, (shape, nbytes) in expect.items():
        assert S.out_shape(S.PRESETS[name], 16000) == shape
        assert S.bytes_per_clip(S.PRESETS[name]) == nbytes
    from dataclasses import replace
    assert S.out_shape(replace(S.R_SPEC, layout="tf"), 16000) == (49, 321)
    assert S.out_shape(replace(S.R_MFCC, layout="tf"), 16000) == (51, 39)
    for n in (640, 8000, 12346, 16384, 48000):
        assert S.out_shape(S.R_SPEC, n)[1] == oracle.spec_num_frames(n)
        assert S.out_shape(S.R_FBANK, n)[0] == oracle.fbank_num_frames(n)
        assert S.out_shape(S.R_MFCC, n)[1] == oracle.mfcc_num_frames(n)


def test_workspace_query(srfe_lib):
    """The device entry points need no global workspace; the query validates its parameters like the others."""
    import ctypes as C
    import speechrecognitionproject_b200 as S
    for fam, p in (("spec", S.R_SPEC), ("fbank", S.R_FBANK), ("mfcc", S.C_MFCC)):
        fn = getattr(srfe_lib, f"srfe_{fam}_workspace_bytes")
        assert fn(C.byref(p.to_c()), 262144, 16000) == 0
        assert fn(C.byref(p.to_c()), -1, 16000) < 0
    bad = S.MfccParams(n_fft=500).to_c()
    assert srfe_lib.srfe_mfcc_workspace_bytes(C.byref(bad), 1, 16000) < 0


def test_validation_errors(srfe_lib):
    bad = [S.SpecParams(nperseg=600, noverlap=300), S.SpecParams(noverlap=640), S.SpecParams(noverlap=319),
           S.FbankParams(frame_len=700), S.FbankParams(nfilt=0), S.FbankParams(frame_step=161),
           S.MfccParams(n_fft=1024), S.MfccParams(n_mfcc=200), S.MfccParams(n_deltas=3), S.MfccParams(hop=321),
           S.MfccParams(win_length=800)]
    for p in bad:
        with pytest.raises(_lib.SrfeError) as ei:
            S.out_shape(p, 16000)
        assert ei.value.code in (-1, -2) and str(ei.value)
    with pytest.raises(_lib.SrfeError):
        S.out_shape(S.R_MFCC, 100)                         # reflect padding needs n > n_fft/2
    with pytest.raises(_lib.SrfeError):
        S.out_shape(S.R_SPEC, 100)


def test_no_device_is_loud(srfe_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert srfe_lib.srfe_device_count() < 0
    with pytest.raises(RuntimeError):
        S.mfcc(torch.zeros(2, 16000))                      # CPU tensor -> host entry point -> no device -> error, no fallback
    cp = S.R_SPEC.to_c()
    rc = srfe_lib.srfe_spec_f32(8, 1, 16000, 16000, C.byref(cp), 8, None)
    assert rc in (-3, -4)


def test_argument_checks(srfe_lib):
    cp = S.R_MFCC.to_c()
    assert srfe_lib.srfe_mfcc_f32(None, 1, 16000, 16000, C.byref(cp), None, None) == -1
    assert srfe_lib.srfe_mfcc_f32(8, 2, 16000, 15999, C.byref(cp), 8, None) == -1      # stride < n_samples
    assert srfe_lib.srfe_mfcc_f32(8, 2, 16000, 16001, C.byref(cp), 8, None) == -1      # odd stride
    assert srfe_lib.srfe_mfcc_f32(4, 1, 16000, 16000, C.byref(cp), 8, None) == -1      # misaligned pcm
    assert b"aligned" in srfe_lib.srfe_last_error_string()
    assert srfe_lib.srfe_mfcc_f32(8, 1, 16000, 16000, None, 8, None) == -1


def test_windows_match_oracle(srfe_lib):
    for name in ("R-SPEC", "C-SPEC"):
        p = S.PRESETS[name]
        np.testing.assert_allclose(_tab(srfe_lib.srfe_spec_window_f64, p.to_c(), p.nperseg),
                                   oracle.tukey_periodic(p.nperseg), atol=1e-15)
    w = _tab(srfe_lib.srfe_fbank_window_f64, S.R_FBANK.to_c(), 512)
    np.testing.assert_allclose(w[:400], np.hamming(400), atol=1e-15)
    assert not w[400:].any()
    for name in ("R-MFCC", "C-MFCC"):
        p = S.PRESETS[name]
        w = _tab(srfe_lib.srfe_mfcc_window_f64, p.to_c(), p.n_fft)
        np.testing.assert_allclose(w, oracle.features._mfcc_window(oracle.PRESETS[name]), atol=1e-15)


def test_filterbanks_match_oracle(srfe_lib):
    for name in ("R-FBANK", "C-FBANK"):
        p = S.PRESETS[name]
        fb = _tab(srfe_lib.srfe_fbank_filters_f64, p.to_c(), p.nfilt * 257).reshape(p.nfilt, 257)
        np.testing.assert_array_equal(fb, oracle.htk_floor_filterbank(oracle.PRESETS[name]))   # floor() bins: exact
    # VTLP-warped banks (legacy/model_8/dataset_top.py:251-252), alpha over the reference's U(0.9, 1.1) range and its ends
    from dataclasses import replace
    base = oracle.htk_floor_filterbank(oracle.R_FBANK)
    moved = 0
    for alpha in (0.9, 0.93, 0.987, 1.0, 1.013, 1.07, 1.1):
        for name in ("R-FBANK", "C-FBANK"):
            p = replace(S.PRESETS[name], vtlp_alpha=alpha)
            fb = _tab(srfe_lib.srfe_fbank_filters_f64, p.to_c(), p.nfilt * 257).reshape(p.nfilt, 257)
            want = oracle.htk_floor_filterbank(replace(oracle.PRESETS[name], vtlp_alpha=alpha))
            np.testing.assert_array_equal(fb, want)
            assert not fb[:, 256].any() or alpha > 1.0
            moved += name == "R-FBANK" and not np.array_equal(fb, base)
    assert moved >= 5                                                    # the warp does move the triangles
    for name in ("R-MFCC", "C-MFCC"):
        p = S.PRESETS[name]
        nb = p.n_fft // 2 + 1
        mb = _tab(srfe_lib.srfe_mfcc_filters_f64, p.to_c(), p.n_mels * nb).reshape(p.n_mels, nb)
        ref = oracle.slaney_mel_filterbank(p.sr, p.n_fft, p.n_mels)
        np.testing.assert_allclose(mb, ref, atol=1e-15)
        assert ((mb != 0) == (ref != 0)).all()
        d = _tab(srfe_lib.srfe_mfcc_dct_f64, p.to_c(), p.n_mfcc * p.n_mels).reshape(p.n_mfcc, p.n_mels)
        np.testing.assert_allclose(d, oracle.dct2_ortho_matrix(p.n_mfcc, p.n_mels), atol=1e-15)
