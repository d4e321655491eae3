"""CPU: the oracle restatement against the golden vectors produced by the UNMODIFIED
reference functions (oracle/make_golden.py), the analytic known answers of SURVEY.md
section 4, and its own float64 'truth' flavour."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest
import scipy.signal

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(ROOT, "tests", "golden", "known_answers.json")) as f:
        return json.load(f)


def test_windows_closed_forms():
    for n in (512, 640):
        np.testing.assert_allclose(oracle.tukey_periodic(n), scipy.signal.get_window(("tukey", 0.25), n), atol=1e-15)
        np.testing.assert_allclose(oracle.hann_periodic(n), scipy.signal.get_window("hann", n, fftbins=True), atol=1e-15)
    np.testing.assert_allclose(oracle.hamming_symmetric(400), np.hamming(400), atol=1e-15)
    w = oracle.tukey_periodic(640)
    assert w[0] == 0 and (w == 1).sum() == 481 and abs((w * w).sum() - 540.0) < 1e-9


def test_spec_restated_equals_reference_golden(golden):
    x = golden["x"]
    for i in range(x.shape[0]):
        np.testing.assert_array_equal(oracle.spec_ref(x[i]), golden["spec_ft"][i])
        np.testing.assert_array_equal(oracle.spec_ref(x[i], oracle.SpecParams(layout="tf")), golden["spec_tf"][i])


def test_fbank_restated_equals_reference_golden(golden):
    x = golden["x"]
    for i in range(x.shape[0]):
        got = oracle.fbank_ref(x[i])
        # identical arithmetic (vectorised filter build): bit-exact float32 outputs
        np.testing.assert_array_equal(got, golden["fbank"][i])


def test_mfcc_golden_is_the_shim_path_selfcheck(golden):
    """NOT an independent pin: golden['mfcc'] is the unmodified compute_mfcc (models/model_mfcc_bgru.py:11-19) run
    with librosa replaced by oracle/librosa_shim.py, i.e. the restatement seen through the reference's own wrapper
    (tensor conversion, np.gradient calls, concatenation order, float32 cast).  The independent evidence for the
    librosa arithmetic is live in tests/test_mfcc_crosscheck.py."""
    x = golden["x"]
    for i in range(x.shape[0]):
        np.testing.assert_array_equal(oracle.mfcc_ref(x[i]), golden["mfcc"][i])


def test_truth_vs_reference_dtype_paths(golden):
    """float64 'truth' and the reference's own dtype path agree far inside the GPU tolerances
    wherever the reference is not itself limited by single precision."""
    x = golden["x"][:6]
    for i in range(x.shape[0]):
        assert np.abs(oracle.mfcc_truth(x[i]) - oracle.mfcc_ref(x[i])).max() < 2e-4
        t, r = oracle.fbank_truth(x[i]), oracle.fbank_ref(x[i]).astype(np.float64)
        m = t >= t.max() - 100
        assert np.abs(t - r)[m].max() < 5e-5


def test_known_answers(kat):
    z = np.zeros(16000, np.float32)
    assert np.allclose(oracle.spec_ref(z), kat["zeros_spec"], atol=1e-5)
    assert np.allclose(oracle.fbank_ref(z), kat["zeros_fbank"], atol=1e-4)
    m = oracle.mfcc_ref(z)
    assert np.allclose(m[0], kat["zeros_mfcc_c0"], atol=1e-3) and np.abs(m[1:]).max() < 1e-9
    fb = oracle.htk_floor_filterbank()
    assert [i for i in range(120) if not fb[i].any()] == kat["fbank_empty_filters"]
    assert (fb != 0).sum() == 397 and not fb[:, 256].any()
    assert (oracle.slaney_mel_filterbank() != 0).sum() == 631
    n = np.arange(16000)
    tone = (1000.0 * np.sin(2 * np.pi * 1000.0 * n / 16000.0)).astype(np.float32)
    psd = oracle.spec_truth(tone, oracle.SpecParams(log=False))
    assert psd[:, 10].argmax() == 40 and abs(psd[40, 10] / kat["tone1k_bin40_psd"] - 1) < 1e-6
    g1 = np.gradient(np.array([kat["gradient_in"]], float), axis=1)
    assert g1[0].tolist() == kat["gradient_d1"] and np.gradient(g1, axis=1)[0].tolist() == kat["gradient_d2"]


def test_shapes_and_frame_counts():
    assert oracle.spec_num_frames(16000) == 49 and oracle.fbank_num_frames(16000) == 98 and oracle.mfcc_num_frames(16000) == 51
    x = oracle.synthetic_corpus(1, 5)[0]
    assert oracle.spec_ref(x, oracle.C_SPEC).shape == (257, 61)
    assert oracle.fbank_ref(x, oracle.C_FBANK).shape == (98, 40)
    assert oracle.mfcc_ref(x, oracle.C_MFCC).shape == (40, 101)
    assert oracle.mfcc_ref(x, oracle.C_MFCC_D2).shape == (120, 101)


def test_corpus_definition():
    c = oracle.synthetic_corpus(100, 1)
    assert c.dtype == np.float32 and c.shape == (100, 16000)
    assert not c[96].any()                                  # every 97th clip: digital silence
    assert np.abs(c[88] - np.round(c[88])).max() > 0        # every 89th: non-integer floats
    assert np.array_equal(c[:3], np.round(c[:3]))
    np.testing.assert_array_equal(oracle.synthetic_corpus(2, 1, start=50), c[50:52])   # shard-independent
