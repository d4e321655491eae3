"""GPU parity: CUDA kernels (through the C ABI) vs the CPU oracle on seeded inputs.

Reference behaviour being matched (citations into /root/reference):
compute_spec (models/model_spec_bgru.py:11-17, model_spec_cnn.py:12-18),
filter_banks (models/model_fbanks_cnn.py:15-66), compute_mfcc
(models/model_mfcc_bgru.py:11-19).  Tolerances: tests/tolerances.py.
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np
import pytest
import torch

import oracle
import speechrecognitionproject_b200 as S
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def corpus():
    return oracle.synthetic_corpus(24, config_index=1)


@pytest.fixture(scope="module")
def edges():
    e = oracle.edge_suite()
    return list(e.keys()), np.stack(list(e.values()))


def _gpu(fn, x, p, **kw):
    y = fn(torch.from_numpy(x).cuda(), p, **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


# ---------------------------------------------------------------- golden (reference) ----
def test_golden_spec(srfe_lib, golden):
    x = golden["x"]
    got = _gpu(S.spec, x, S.R_SPEC)
    assert got.shape == golden["spec_ft"].shape == (x.shape[0], 321, 49)
    H.check_logspec(got, golden["spec_ft"].astype(np.float64), "golden spec ft")
    got_tf = _gpu(S.spec, x, S.R_SPEC, layout="tf")
    assert got_tf.shape == (x.shape[0], 49, 321)
    np.testing.assert_array_equal(got_tf, np.ascontiguousarray(got.transpose(0, 2, 1)))
    H.check_logspec(got_tf, golden["spec_tf"].astype(np.float64), "golden spec tf")


def test_golden_fbank(srfe_lib, golden):
    x = golden["x"]
    got = _gpu(S.fbank, x, S.R_FBANK)
    assert got.shape == golden["fbank"].shape == (x.shape[0], 98, 120)
    H.check_logmel(got, golden["fbank"].astype(np.float64), "golden fbank")


def test_golden_mfcc(srfe_lib, golden):
    x = golden["x"]
    got = _gpu(S.mfcc, x, S.R_MFCC)
    assert got.shape == golden["mfcc"].shape == (x.shape[0], 39, 51)
    H.check_mfcc(got, golden["mfcc"].astype(np.float64), "golden mfcc")


# ---------------------------------------------------------------- presets vs oracle ----
@pytest.mark.parametrize("name", ["R-SPEC", "C-SPEC"])
def test_spec_presets(srfe_lib, corpus, name):
    p = S.PRESETS[name]
    op = H.to_oracle_params(p)
    got = _gpu(S.spec, corpus, p)
    truth = H.oracle_batch(oracle.spec_truth, corpus, op)
    assert got.shape == truth.shape
    H.check_logspec(got, truth, name)
    # raw PSD: relative error on the power spectra themselves
    praw = replace(p, log=False)
    got_raw = _gpu(S.spec, corpus, praw)
    truth_raw = H.oracle_batch(oracle.spec_truth, corpus, H.to_oracle_params(praw))
    H.check_psd(got_raw, truth_raw, name + " raw")
    # and against the imported-reference flavour (scipy, single precision)
    ref = H.oracle_batch(oracle.spec_ref, corpus, op)
    H.check_logspec(got, ref.astype(np.float64), name + " vs scipy-fp32")


@pytest.mark.parametrize("name", ["R-FBANK", "C-FBANK"])
def test_fbank_presets(srfe_lib, corpus, name):
    p = S.PRESETS[name]
    op = H.to_oracle_params(p)
    got = _gpu(S.fbank, corpus, p)
    truth = H.oracle_batch(oracle.fbank_truth, corpus, op)
    assert got.shape == truth.shape
    stats = H.check_logmel(got, truth, name)
    assert stats["frac_outside"] < 0.12, stats          # R-FBANK: 10 of 120 filters are structurally empty
    ref = H.oracle_batch(oracle.fbank_ref, corpus, op)
    H.check_logmel(got, ref.astype(np.float64), name + " vs reference dtype path")


@pytest.mark.parametrize("name", ["R-MFCC", "C-MFCC", "C-MFCC-D2"])
def test_mfcc_presets(srfe_lib, corpus, name):
    p = S.PRESETS[name]
    op = H.to_oracle_params(p)
    got = _gpu(S.mfcc, corpus, p)
    truth = H.oracle_batch(oracle.mfcc_truth, corpus, op)
    assert got.shape == truth.shape
    H.check_mfcc(got, truth, name)
    ref = H.oracle_batch(oracle.mfcc_ref, corpus, op)
    H.check_mfcc(got, ref.astype(np.float64), name + " vs restated librosa dtype path")
    got_tf = _gpu(S.mfcc, corpus, p, layout="tf")
    np.testing.assert_array_equal(got_tf, np.ascontiguousarray(got.transpose(0, 2, 1)))


# ---------------------------------------------------------------- edge clips -----------
def test_edge_suite(srfe_lib, edges):
    names, x = edges
    got = _gpu(S.mfcc, x, S.R_MFCC)
    truth = H.oracle_batch(oracle.mfcc_truth, x, oracle.R_MFCC)
    for i, n in enumerate(names):
        H.check_mfcc(got[i:i + 1], truth[i:i + 1], f"edge mfcc {n}")
    got = _gpu(S.fbank, x, S.R_FBANK)
    truth = H.oracle_batch(oracle.fbank_truth, x, oracle.R_FBANK)
    for i, n in enumerate(names):
        H.check_logmel(got[i:i + 1], truth[i:i + 1], f"edge fbank {n}")
    got = _gpu(S.spec, x, S.R_SPEC)
    truth = H.oracle_batch(oracle.spec_truth, x, oracle.R_SPEC)
    for i, n in enumerate(names):
        H.check_logspec(got[i:i + 1], truth[i:i + 1], f"edge spec {n}")


def test_known_answers(srfe_lib):
    z = torch.zeros(2, 16000, device="cuda")
    s = S.spec(z).cpu().numpy()
    np.testing.assert_allclose(s, np.log(np.float32(1e-10)), rtol=0, atol=2e-5)
    f = S.fbank(z).cpu().numpy()
    np.testing.assert_allclose(f, 20 * np.log10(np.finfo(float).eps), rtol=0, atol=1e-3)
    m = S.mfcc(z).cpu().numpy()
    np.testing.assert_allclose(m[:, 0, :], -100.0 * np.sqrt(128.0), rtol=0, atol=1e-3)
    assert np.abs(m[:, 1:, :]).max() <= 1e-3
    # the 10 structurally empty filters of the reference's 120-band bank are constant for ANY input
    x = torch.from_numpy(oracle.synthetic_corpus(3, 2)).cuda()
    f = S.fbank(x).cpu().numpy()
    empty = [0, 2, 4, 7, 9, 11, 14, 17, 21, 25]
    np.testing.assert_allclose(f[:, :, empty], 20 * np.log10(np.finfo(float).eps), rtol=0, atol=1e-3)
    # pure tone at a bin centre: PSD peak 2 (A sum(w)/2)^2 / (fs sum(w^2)) at bin 40
    n = np.arange(16000)
    tone = (1000.0 * np.sin(2 * np.pi * 1000.0 * n / 16000.0)).astype(np.float32)
    psd = S.spec(torch.from_numpy(tone).cuda(), replace(S.R_SPEC, log=False)).cpu().numpy()
    w = oracle.tukey_periodic(640)
    expect = 2.0 * (1000.0 * w.sum() / 2.0) ** 2 / (16000.0 * (w * w).sum())
    assert psd.shape == (321, 49) and psd[:, 10].argmax() == 40
    np.testing.assert_allclose(psd[40], expect, rtol=2e-4)


# ---------------------------------------------------------------- interface behaviour ---
def test_single_clip_signatures(srfe_lib, corpus):
    s = torch.from_numpy(corpus[0])
    for dev in ("cpu", "cuda"):
        x = s.to(dev)
        assert S.compute_spec(x).shape == (321, 49)
        assert S.compute_spec(x, transpose=True).shape == (49, 321)
        assert S.filter_banks(x).shape == (98, 120)
        m = S.compute_mfcc(x)
        assert m.shape == (39, 51) and m.dtype == torch.float32 and m.device.type == dev and m.is_contiguous()


def test_host_path_matches_device_path(srfe_lib, corpus):
    x = torch.from_numpy(corpus)
    for fn, p in ((S.spec, S.R_SPEC), (S.fbank, S.R_FBANK), (S.mfcc, S.R_MFCC), (S.mfcc, S.C_MFCC)):
        a = fn(x.cuda(), p).cpu()
        b = fn(x, p)                       # host entry point: H2D -> kernel -> D2H
        c = fn(x.pin_memory(), p)
        assert not b.is_cuda and torch.equal(a, b) and torch.equal(a, c)


def test_ragged_and_strided_inputs(srfe_lib, corpus):
    x = torch.from_numpy(corpus).cuda()
    full = S.mfcc(x)
    # empty batch
    assert S.mfcc(x[:0]).shape == (0, 39, 51)
    assert S.fbank(x[:0]).shape == (0, 98, 120)
    # batch of one, and every clip independent of its neighbours (bit-exact)
    for i in (0, 7, 23):
        assert torch.equal(S.mfcc(x[i:i + 1])[0], full[i])
    # row-strided view (clip_stride > n_samples), odd-offset view -> copied
    wide = torch.zeros(x.size(0), 16384, device="cuda")
    wide[:, :16000] = x
    assert torch.equal(S.mfcc(wide[:, :16000]), full)
    shifted = torch.zeros(x.size(0), 16001, device="cuda")
    shifted[:, 1:] = x
    assert torch.equal(S.mfcc(shifted[:, 1:]), full)
    # other clip lengths: frame counts follow the reference formulas
    for n in (8000, 12345 * 2 // 2 + 1, 16384):
        xs = x[:3, :n] if n <= 16000 else torch.cat([x[:3], x[:3, : n - 16000]], 1)
        xs_np = xs.cpu().numpy()
        if n % 2 == 0:
            got = S.mfcc(xs).cpu().numpy()
            H.check_mfcc(got, H.oracle_batch(oracle.mfcc_truth, xs_np, oracle.R_MFCC), f"mfcc n={n}")
            got = S.spec(xs).cpu().numpy()
            H.check_logspec(got, H.oracle_batch(oracle.spec_truth, xs_np, oracle.R_SPEC), f"spec n={n}")
        got = S.fbank(xs).cpu().numpy()
        H.check_logmel(got, H.oracle_batch(oracle.fbank_truth, xs_np, oracle.R_FBANK), f"fbank n={n}")


def test_errors_are_loud(srfe_lib):
    x = torch.zeros(2, 16000, device="cuda")
    with pytest.raises(RuntimeError):
        S.spec(x, S.SpecParams(nperseg=600, noverlap=300))
    with pytest.raises(RuntimeError):
        S.mfcc(x, replace(S.R_MFCC, n_mfcc=200))
    with pytest.raises(TypeError):
        S.mfcc(x.double())
    with pytest.raises(RuntimeError):
        S.mfcc(torch.zeros(1, 200, device="cuda"))


def test_deterministic_and_stream_ordered(srfe_lib, corpus):
    x = torch.from_numpy(corpus).cuda()
    a = S.fbank(x)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        b = S.fbank(x)
    st.synchronize()
    assert torch.equal(a, b)


def test_shards_concatenate_bit_exact(srfe_lib, corpus):
    """Multi-GPU contract: features of a batch == concatenation of its shards."""
    from speechrecognitionproject_b200.sharding import shard_range
    x = torch.from_numpy(corpus).cuda()
    for fn, p in ((S.spec, S.R_SPEC), (S.fbank, S.C_FBANK), (S.mfcc, S.C_MFCC)):
        whole = fn(x, p)
        for world in (2, 4, 8):
            parts = [fn(x[slice(*shard_range(x.size(0), r, world))], p) for r in range(world)]
            assert torch.equal(torch.cat(parts), whole)


def test_launch_counter(srfe_lib, corpus):
    x = torch.from_numpy(corpus).cuda()
    n0 = S.launch_count()
    S.mfcc(x); S.spec(x); S.fbank(x)
    assert S.launch_count() == n0 + 3


# ---------------------------------------------------------------- generic (non-preset) shapes ----
# The presets hit shape-specialised instantiations (compile-time mel bank shape / DCT N-tiles);
# these cases force the generic code paths: runtime ELL metadata, runtime N-tile count, CUDA-core
# DCT (n_mels not a multiple of 8), other window / hop / filter counts, short clips.
GENERIC_MFCC = [
    S.MfccParams(n_fft=512, win_length=400, hop=160, n_mels=40, n_mfcc=13, n_deltas=2),
    S.MfccParams(n_fft=512, win_length=512, hop=256, n_mels=64, n_mfcc=20, n_deltas=1),
    S.MfccParams(n_fft=640, win_length=480, hop=240, n_mels=80, n_mfcc=24, n_deltas=0, top_db=None),
    S.MfccParams(n_fft=640, hop=320, n_mels=100, n_mfcc=13, n_deltas=2),          # CUDA-core DCT path
    S.MfccParams(n_fft=512, win_length=320, hop=160, n_mels=128, n_mfcc=64, n_deltas=0, fmin=20.0, fmax=7600.0),
    S.MfccParams(n_fft=640, hop=320, n_mels=128, n_mfcc=13, n_deltas=2, top_db=40.0),
]


@pytest.mark.parametrize("idx", range(len(GENERIC_MFCC)))
def test_mfcc_generic_shapes(srfe_lib, corpus, idx):
    p = GENERIC_MFCC[idx]
    x = corpus[:6]
    got = _gpu(S.mfcc, x, p)
    truth = H.oracle_batch(oracle.mfcc_truth, x, H.to_oracle_params(p))
    assert got.shape == truth.shape
    H.check_mfcc(got, truth, f"generic mfcc {idx}")
    got_tf = _gpu(S.mfcc, x, p, layout="tf")
    np.testing.assert_array_equal(got_tf, np.ascontiguousarray(got.transpose(0, 2, 1)))


GENERIC_FBANK = [
    S.FbankParams(nfilt=26), S.FbankParams(nfilt=64, frame_len=512, frame_step=128),
    S.FbankParams(nfft=640, frame_len=640, frame_step=320, nfilt=80, preemph=0.95),
    S.FbankParams(nfft=640, frame_len=400, frame_step=160, nfilt=120), S.FbankParams(nfilt=200, preemph=0.0),
]


@pytest.mark.parametrize("idx", range(len(GENERIC_FBANK)))
def test_fbank_generic_shapes(srfe_lib, corpus, idx):
    p = GENERIC_FBANK[idx]
    x = corpus[:6]
    got = _gpu(S.fbank, x, p)
    truth = H.oracle_batch(oracle.fbank_truth, x, H.to_oracle_params(p))
    assert got.shape == truth.shape
    H.check_logmel(got, truth, f"generic fbank {idx}")


@pytest.mark.parametrize("nperseg,noverlap", [(512, 384), (640, 0), (512, 0), (640, 512)])
def test_spec_generic_shapes(srfe_lib, corpus, nperseg, noverlap):
    x = corpus[:5]
    for layout in ("ft", "tf"):
        p = S.SpecParams(nperseg=nperseg, noverlap=noverlap, layout=layout)
        got = _gpu(S.spec, x, p)
        truth = H.oracle_batch(oracle.spec_truth, x, H.to_oracle_params(p))
        assert got.shape == truth.shape
        H.check_logspec(got, truth, f"spec {nperseg}/{noverlap} {layout}")


def test_too_long_for_fused_mfcc_is_an_error_not_a_fallback(srfe_lib):
    x = torch.zeros(2, 16000, device="cuda")
    with pytest.raises(RuntimeError, match="SRFE_ERR_TOO_LARGE"):
        S.mfcc(x, S.MfccParams(n_fft=512, win_length=320, hop=40, n_mels=128, n_mfcc=64, n_deltas=0))


def test_short_and_long_clips(srfe_lib):
    rng = np.random.default_rng(5)
    for n in (700, 1024, 4000, 32000):
        x = np.round(rng.standard_normal((3, n)) * 2000).astype(np.float32)
        got = _gpu(S.mfcc, x, S.R_MFCC)
        H.check_mfcc(got, H.oracle_batch(oracle.mfcc_truth, x, oracle.R_MFCC), f"mfcc n={n}")
        got = _gpu(S.fbank, x, S.R_FBANK)
        H.check_logmel(got, H.oracle_batch(oracle.fbank_truth, x, oracle.R_FBANK), f"fbank n={n}")
        got = _gpu(S.spec, x, S.R_SPEC, layout="tf")
        H.check_logspec(got, H.oracle_batch(oracle.spec_truth, x, replace(oracle.R_SPEC, layout="tf")), f"spec n={n}")


def test_large_batch_properties(srfe_lib):
    """Full-size batch (BASELINE cfg2/cfg3 sizes): size-independent properties instead of a CPU oracle run --
    every clip equals its stand-alone result, duplicated clips give identical features, silence gives the known constants."""
    torch.manual_seed(0)
    B = 4096
    x = (torch.randn(B, 16000, device="cuda") * 2500).round()
    x[1000] = x[7]; x[4095] = x[7]; x[2048] = 0
    for fn, p in ((S.spec, S.C_SPEC), (S.fbank, S.C_FBANK), (S.mfcc, S.C_MFCC), (S.mfcc, S.R_MFCC)):
        y = fn(x, p)
        assert torch.isfinite(y).all()
        assert torch.equal(y[1000], y[7]) and torch.equal(y[4095], y[7])
        for i in (0, 7, 2048, 4095):
            assert torch.equal(fn(x[i:i + 1], p)[0], y[i])
    assert torch.allclose(S.mfcc(x[2048:2049], S.R_MFCC)[0, 0], torch.tensor(-100.0 * 128 ** 0.5, device="cuda"), atol=1e-3)


def test_int16_ingest_bit_identical(srfe_lib, corpus):
    """int16 PCM (the wav's own type, dataset.py:103) converted on load inside the kernel gives exactly the
    features of its float32 copy (dataset.py:117) -- device and host entry points, every family."""
    xi = torch.from_numpy(np.round(corpus).clip(-32768, 32767).astype(np.int16))
    xf = xi.float()
    for fn, p in ((S.spec, S.R_SPEC), (S.spec, replace(S.C_SPEC, layout="tf")), (S.fbank, S.R_FBANK), (S.fbank, S.C_FBANK),
                  (S.mfcc, S.R_MFCC), (S.mfcc, S.C_MFCC), (S.mfcc, GENERIC_MFCC[3])):
        a = fn(xf.cuda(), p)
        b = fn(xi.cuda(), p)
        assert b.dtype == torch.float32 and torch.equal(a, b)
        c = fn(xi, p)                                   # host entry point, int16 H2D
        assert not c.is_cuda and torch.equal(a.cpu(), c)
    odd = torch.zeros(4, 16001, dtype=torch.int16)
    odd[:, 1:] = xi[:4]
    assert torch.equal(S.mfcc(odd[:, 1:].cuda()), S.mfcc(xf[:4].cuda()))      # misaligned view -> copied
    with pytest.raises(TypeError):
        S.mfcc(xi.to(torch.int32).cuda())


def test_results_do_not_depend_on_launch_configuration(srfe_lib, corpus):
    """compute-sanitizer is closed on this pool; as a race / hazard screen every family is run under several launch
    configurations (srfe_set_tuning: warps / ctas / cpc, and for the classic MFCC kernel every DCT tile shape
    dct_cb x dct_pq) and repeatedly: results must be bit-identical everywhere -- no arithmetic depends on the
    configuration (each DCT output is one ascending-f FFMA chain whatever the tile)."""
    x = torch.from_numpy(np.concatenate([corpus, corpus[:7]])).cuda()        # 31 clips: ragged last group
    cases = [(S.spec, S.R_SPEC), (S.spec, replace(S.C_SPEC, layout="tf")), (S.fbank, S.R_FBANK), (S.fbank, S.C_FBANK),
             (S.mfcc, S.R_MFCC), (S.mfcc, S.C_MFCC_D2)]
    configs = [None, (4, 2, 1), (8, 2, 4), (5, 2, 2), (9, 1, 1), (16, 1, 8), (13, 1, 2)]
    try:
        for fn, p in cases:
            S.set_tuning()
            if fn is S.mfcc:
                S.set_tuning(mfcc_tc=1)                  # the classic kernel: the one these launch shapes apply to
            base = fn(x, p)
            trials = [(cfg, None) for cfg in configs]
            if fn is S.mfcc:
                trials += [(cfg, (cb, pq)) for cfg in (None, (5, 2, 1)) for cb in (2, 3, 4, 5, 6, 8) for pq in (1, 2)]
            for cfg, dct in trials:
                S.set_tuning(warps=0, ctas=0, cpc=0, dct_cb=0, dct_pq=0)
                if cfg is not None:
                    S.set_tuning(warps=cfg[0], ctas=cfg[1], cpc=cfg[2])
                if dct is not None:
                    S.set_tuning(dct_cb=dct[0], dct_pq=dct[1])
                try:
                    y1 = fn(x, p)
                except RuntimeError as e:                    # a forced configuration may not fit in shared memory
                    assert "SRFE_ERR_TOO_LARGE" in str(e)
                    continue
                for _ in range(3):
                    assert torch.equal(fn(x, p), y1), f"non-deterministic: {type(p).__name__} cfg={cfg} dct={dct}"
                assert torch.equal(y1, base), f"configuration-dependent result: {type(p).__name__} cfg={cfg} dct={dct}"
    finally:
        S.set_tuning()


def test_thread_safety_two_host_threads(srfe_lib, corpus):
    """Re-entrancy (SURVEY 8b 'Threading'): two host threads, each on its own stream, hammer different feature
    families (first calls race on the table cache); results must equal the single-threaded ones."""
    import threading
    x = torch.from_numpy(corpus).cuda()
    p1 = S.MfccParams(n_fft=512, win_length=400, hop=160, n_mels=64, n_mfcc=20, n_deltas=1)     # fresh cache entries
    p2 = S.FbankParams(nfilt=48)
    want = {}
    errs = []

    def work(name, fn, p):
        try:
            st = torch.cuda.Stream()
            outs = []
            with torch.cuda.stream(st):
                for _ in range(20):
                    outs.append(fn(x, p))
            st.synchronize()
            assert all(torch.equal(o, outs[0]) for o in outs)
            want[name] = outs[0]
        except Exception as e:          # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=("a", S.mfcc, p1)), threading.Thread(target=work, args=("b", S.fbank, p2)),
          threading.Thread(target=work, args=("c", S.mfcc, p1))]
    for t in ts: t.start()
    for t in ts: t.join()
    assert not errs, errs
    assert torch.equal(want["a"], want["c"]) and torch.equal(want["a"], S.mfcc(x, p1)) and torch.equal(want["b"], S.fbank(x, p2))
    # host entry points from two threads (per-thread staging workspaces)
    xc = torch.from_numpy(corpus)
    res = {}
    def host(name):
        res[name] = S.mfcc(xc, S.R_MFCC)
    hs = [threading.Thread(target=host, args=(i,)) for i in range(3)]
    for t in hs: t.start()
    for t in hs: t.join()
    ref = S.mfcc(x, S.R_MFCC).cpu()
    assert all(torch.equal(res[i], ref) for i in range(3))


def test_one_process_two_devices(srfe_lib, corpus):
    """one process driving several GPUs is legal -- tables are cached per device and the launch goes to the
    device that owns the tensor, whatever the current device is.  Skipped on single-GPU boxes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    x0 = torch.from_numpy(corpus).to("cuda:0")
    x1 = x0.to("cuda:1")
    for fn, p in ((S.mfcc, S.R_MFCC), (S.spec, S.R_SPEC), (S.fbank, S.R_FBANK)):
        y0 = fn(x0, p)
        with torch.cuda.device(0):
            y1 = fn(x1, p)                      # current device 0, data on device 1
        assert y1.device == x1.device
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        assert torch.equal(y0.cpu(), y1.cpu())
    xc = torch.from_numpy(corpus)
    assert torch.equal(S.mfcc(xc, S.R_MFCC, device=1), S.mfcc(xc, S.R_MFCC, device=0))      # host entry points, per device


@pytest.mark.parametrize("seed", [11, 12])
def test_random_parameter_sets_and_lengths(srfe_lib, seed):
    """Random feature parameters, clip lengths and batch sizes against the float64 oracle: walks launch shapes, window
    extents, filterbank shapes and DCT tile choices the presets never reach (scripts/fuzz_shapes.py runs more)."""
    rng = np.random.default_rng(seed)
    done = 0
    for case in range(30):
        n_fft = int(rng.choice([512, 640]))
        n_samples = int(rng.choice([4000, 8000, 16000, 24000, 12346]))
        x = oracle.synthetic_corpus(int(rng.choice([1, 3])), config_index=30 + case % 5, n_samples=n_samples)
        xd = torch.from_numpy(x).cuda()
        fam = rng.choice(["mfcc", "mfcc", "fbank", "spec"])
        try:
            if fam == "mfcc":
                n_mels = int(rng.choice([20, 26, 40, 64, 80, 128, 200]))
                p = S.MfccParams(n_fft=n_fft, win_length=int(rng.choice([n_fft, 400, 320, 256])),
                                 hop=int(rng.choice([80, 128, 160, 200, 320])), n_mels=n_mels,
                                 n_mfcc=int(rng.integers(1, min(64, n_mels) + 1)), n_deltas=int(rng.integers(0, 3)),
                                 layout=str(rng.choice(["ft", "tf"])))
                got = S.mfcc(xd, p).cpu().numpy()
                truth = H.oracle_batch(oracle.mfcc_truth, x, H.to_oracle_params(p))
                if p.layout == "tf":
                    truth = truth.transpose(0, 2, 1)                  # the oracle is always [coeff, time]
                H.check_mfcc(got, truth, str(p))
            elif fam == "fbank":
                p = S.FbankParams(nfft=n_fft, frame_len=int(rng.choice([400, 320, 512])), frame_step=int(rng.choice([80, 160, 200])),
                                  nfilt=int(rng.choice([13, 26, 40, 64, 120])))
                got = S.fbank(xd, p).cpu().numpy()
                H.check_logmel(got, H.oracle_batch(oracle.fbank_truth, x, H.to_oracle_params(p)), str(p))
            else:
                p = S.SpecParams(nperseg=n_fft, noverlap=int(rng.choice([n_fft // 2, n_fft // 4, n_fft - 160, 0])), log=False,
                                 layout=str(rng.choice(["ft", "tf"])))
                got = S.spec(xd, p).cpu().numpy()
                H.check_psd(got, H.oracle_batch(oracle.spec_truth, x, H.to_oracle_params(p)), str(p))
            done += 1
        except RuntimeError as e:                                     # shapes the kernels decline are declined loudly
            assert "SRFE_ERR_TOO_LARGE" in str(e) or "SRFE_ERR_UNSUPPORTED" in str(e) or "SRFE_ERR_BAD_ARG" in str(e), str(e)
    assert done >= 20


def test_cuda_graph_capture_and_replay(srfe_lib, corpus):
    """The device entry points are plain stream-ordered launches (no allocation, no synchronisation once the tables of
    a parameter set exist): they can be captured into a CUDA graph and replayed on new data."""
    x = torch.from_numpy(corpus).cuda()
    static_in = x.clone()
    want = [S.mfcc(x, S.R_MFCC, layout="tf"), S.fbank(x, S.R_FBANK), S.spec(x, S.R_SPEC)]       # warm-up: tables, attributes
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [S.mfcc(static_in, S.R_MFCC, layout="tf"), S.fbank(static_in, S.R_FBANK), S.spec(static_in, S.R_SPEC)]
    g.replay()
    torch.cuda.synchronize()
    for o, w in zip(outs, want):
        assert torch.equal(o, w)
    static_in.copy_(x.flip(0))                          # new data, same graph
    g.replay()
    torch.cuda.synchronize()
    for o, w in zip(outs, want):
        assert torch.equal(o, w.flip(0))


# ---------------------------------------------------------------- tcgen05 MFCC kernel ----
TC_CASES = [
    ("C-MFCC", {}), ("C-MFCC", {"layout": "tf"}), ("C-MFCC-D2", {}), ("C-MFCC-D2", {"layout": "tf"}),
    ("R-MFCC", {}), ("R-MFCC", {"layout": "tf"}), ("R-MFCC", {"n_deltas": 1}), ("R-MFCC", {"n_deltas": 0, "n_mfcc": 20}),
    ("C-MFCC", {"n_mels": 80, "n_mfcc": 13, "n_deltas": 2}), ("C-MFCC", {"n_mels": 64, "n_mfcc": 33, "top_db": None}),
    ("C-MFCC", {"n_mels": 32, "n_mfcc": 12, "n_deltas": 1, "layout": "tf"}), ("C-MFCC", {"n_mels": 128, "n_mfcc": 1}),
]


@pytest.mark.parametrize("idx", range(len(TC_CASES)))
def test_mfcc_tensor_core_kernel_vs_classic_and_oracle(srfe_lib, idx):
    """srfe_mfcc_tc_kernel (DCT-II as 3xTF32 tcgen05 MMAs, accumulators in TMEM) against the classic CUDA-core kernel and
    the float64 oracle: every preset, both layouts, 0 / 1 / 2 deltas, odd coefficient counts, batch sizes that leave
    CTAs with 1, 2 and 3 clips, int16 ingest."""
    name, over = TC_CASES[idx]
    p = replace(S.PRESETS[name], **over)
    x = oracle.synthetic_corpus(333, config_index=9)
    x[5] = 0.0                                               # digital silence: amin floor, c0 only
    x[6, 8000:] = 0.0                                        # half-silent: the top_db clamp bites
    xd = torch.from_numpy(x).cuda()
    try:
        for n in (1, 7, 148, 149, 333):
            S.set_tuning(mfcc_tc=2)
            tc = S.mfcc(xd[:n], p)
            tc2 = S.mfcc(xd[:n], p)
            S.set_tuning(mfcc_tc=1)
            classic = S.mfcc(xd[:n], p)
            assert torch.equal(tc, tc2), "tcgen05 kernel is not deterministic"
            assert float((tc - classic).abs().max()) <= 2e-4, (name, over, n)
        S.set_tuning(mfcc_tc=2)
        got = S.mfcc(xd[:12], p).cpu().numpy()
        truth = H.oracle_batch(oracle.mfcc_truth, x[:12], H.to_oracle_params(p))     # the oracle is [rows, frames]
        if p.layout == "tf":
            truth = truth.transpose(0, 2, 1)
        H.check_mfcc(got, truth, f"tc {name} {over}")
        got16 = S.mfcc(xd[:12].to(torch.int16), p).cpu().numpy()
        np.testing.assert_array_equal(got16, got)
    finally:
        S.set_tuning()


def test_mfcc_tensor_core_kernel_limits(srfe_lib):
    """128 frames without deltas, 112 with two (each TMEM quadrant carries a halo of n_deltas frames); beyond that the
    tcgen05 kernel refuses (mfcc_tc = 2) and the default dispatch uses the classic kernel -- same results either way."""
    base = replace(S.C_MFCC, hop=160)
    try:
        for n_deltas, frames_ok in ((0, 128), (2, 112)):
            p = replace(base, n_deltas=n_deltas)
            for T, ok in ((frames_ok, True), (frames_ok + 1, False)):
                x = torch.from_numpy(oracle.synthetic_corpus(5, config_index=10, n_samples=160 * (T - 1) + 40)).cuda()
                assert S.out_shape(p, x.size(1))[1] == T
                S.set_tuning(mfcc_tc=1)
                classic = S.mfcc(x, p)
                S.set_tuning(mfcc_tc=2)
                if ok:
                    assert float((S.mfcc(x, p) - classic).abs().max()) <= 2e-4
                else:
                    with pytest.raises(RuntimeError, match="SRFE_ERR_UNSUPPORTED"):
                        S.mfcc(x, p)
                S.set_tuning()
                assert float((S.mfcc(x, p) - classic).abs().max()) <= 2e-4
        # parameter sets it cannot take (n_mels not a multiple of 16) go to the classic kernel silently, refuse when forced
        p = replace(S.C_MFCC, n_mels=40, n_mfcc=13)
        x = torch.from_numpy(oracle.synthetic_corpus(3, config_index=10)).cuda()
        S.set_tuning()
        y = S.mfcc(x, p)
        S.set_tuning(mfcc_tc=2)
        with pytest.raises(RuntimeError, match="SRFE_ERR_UNSUPPORTED"):
            S.mfcc(x, p)
        S.set_tuning(mfcc_tc=1)
        assert torch.equal(S.mfcc(x, p), y)
    finally:
        S.set_tuning()


FT_CASES = [
    ("R-FBANK", {}), ("C-FBANK", {}), ("R-FBANK", {"nfilt": 26}), ("R-FBANK", {"nfilt": 128}), ("R-FBANK", {"nfilt": 64, "preemph": 0.0}),
    ("R-FBANK", {"frame_len": 512, "frame_step": 128, "nfilt": 80}), ("R-FBANK", {"vtlp_alpha": 0.9}), ("C-FBANK", {"vtlp_alpha": 1.08}),
]


@pytest.mark.parametrize("idx", range(len(FT_CASES)))
def test_fbank_tensor_core_kernel_vs_classic_and_oracle(srfe_lib, idx):
    """srfe_fbank_tc_kernel (filter projection as bf16 hi/mid tcgen05 MMAs, weights resident in TMEM) against the classic
    CUDA-core kernel and the float64 oracle: presets, other bank sizes / framings (window 512 = the generic instantiation),
    VTLP-warped banks, batch sizes that leave CTAs with 1, 2 and 3 clips and partly filled tiles, clip lengths with an odd
    frame count, int16 ingest."""
    name, over = FT_CASES[idx]
    p = replace(S.PRESETS[name], **over)
    x = oracle.synthetic_corpus(333, config_index=11)
    x[5] = 0.0                                               # digital silence: every band sum is an exact zero -> eps
    x[6, 8000:] = 0.0
    xd = torch.from_numpy(x).cuda()
    try:
        for n in (1, 2, 7, 148, 149, 333):
            S.set_tuning(fbank_tc=2)
            tc = S.fbank(xd[:n], p)
            tc2 = S.fbank(xd[:n], p)
            S.set_tuning(fbank_tc=1)
            classic = S.fbank(xd[:n], p)
            assert torch.equal(tc, tc2), "tcgen05 kernel is not deterministic"
            assert torch.equal(tc[5:6], classic[5:6]) or n <= 5            # exact zeros: bit-identical
            inside = classic > classic.amax(dim=(1, 2), keepdim=True) - 100.0
            assert float((tc - classic).abs()[inside].max()) <= 4e-4, (name, over, n)
        for n_samples in (15840, 15700, 1000, 640):          # odd frame counts repeat the last frame in the pair stream
            xs = xd[:9, :n_samples].contiguous()
            S.set_tuning(fbank_tc=2)
            tc = S.fbank(xs, p)
            S.set_tuning(fbank_tc=1)
            classic = S.fbank(xs, p)
            assert tc.shape == classic.shape
            inside = classic > classic.amax(dim=(1, 2), keepdim=True) - 100.0
            assert float((tc - classic).abs()[inside].max()) <= 4e-4, (name, over, n_samples)
        S.set_tuning(fbank_tc=2)
        got = S.fbank(xd[:12], p).cpu().numpy()
        truth = H.oracle_batch(oracle.fbank_truth, x[:12], H.to_oracle_params(p))
        H.check_logmel(got, truth, f"tc {name} {over}")
        got16 = S.fbank(xd[:12].to(torch.int16), p).cpu().numpy()
        np.testing.assert_array_equal(got16, got)
    finally:
        S.set_tuning()


def test_fbank_tensor_core_kernel_limits(srfe_lib, corpus):
    """Parameter sets the tcgen05 FBANK kernel does not take (n_fft 640, more than 128 filters) stay on the classic kernel
    silently and are refused when it is forced; the default dispatch is always the classic kernel (opt-in only)."""
    x = torch.from_numpy(np.concatenate([corpus] * 200)[:4200]).cuda()
    try:
        for p in (S.FbankParams(nfft=640, frame_len=640, frame_step=320, nfilt=80), S.FbankParams(nfilt=200)):
            S.set_tuning()
            y = S.fbank(x[:20], p)
            S.set_tuning(fbank_tc=2)
            with pytest.raises(RuntimeError, match="SRFE_ERR_UNSUPPORTED"):
                S.fbank(x[:20], p)
            S.set_tuning(fbank_tc=1)
            assert torch.equal(S.fbank(x[:20], p), y)
        for p in (S.R_FBANK, S.C_FBANK):                     # never dispatched automatically: a clip's features must not
            for n in (20, 4200):                             # depend on the batch it arrives in (the kernels agree to 1.4e-4)
                S.set_tuning()
                y = S.fbank(x[:n], p)
                S.set_tuning(fbank_tc=1)
                assert torch.equal(S.fbank(x[:n], p), y)
                S.set_tuning(fbank_tc=2)
                b = S.fbank(x[:n], p)
                inside = y > y.amax(dim=(1, 2), keepdim=True) - 100.0
                assert float((y - b).abs()[inside].max()) <= 4e-4
    finally:
        S.set_tuning()


@pytest.mark.parametrize("alpha", [0.9, 1.0, 1.08])
def test_fbank_vtlp_warped_bank(srfe_lib, corpus, alpha):
    """the VTLP warp of legacy/model_8/dataset_top.py:251-252 as a filter-bank option (one alpha per call)."""
    p = replace(S.R_FBANK, vtlp_alpha=alpha)
    got = _gpu(S.fbank, corpus[:8], p)
    truth = H.oracle_batch(oracle.fbank_truth, corpus[:8], H.to_oracle_params(p))
    H.check_logmel(got, truth, f"vtlp {alpha}")
    if alpha != 1.0:
        assert np.abs(got - _gpu(S.fbank, corpus[:8], S.R_FBANK)).max() > 0.5       # and it is a different bank
    with pytest.raises(RuntimeError, match="SRFE_ERR_BAD_ARG"):
        S.fbank(torch.from_numpy(corpus[:2]).cuda(), replace(S.R_FBANK, vtlp_alpha=3.0))


@pytest.mark.parametrize("names", [("R-SPEC", "R-FBANK"), ("C-SPEC", "C-FBANK"), ("C-SPEC", "R-FBANK"), ("R-SPEC", "C-FBANK")])
def test_spec_and_fbank_in_one_launch(srfe_lib, corpus, names):
    """srfe_spec_fbank_* = the two feature sets the reference's ensemble computes from the same batch
    (analyst_training.py:91-94) out of one kernel; bit-identical to the separate entry points."""
    ps, pf = S.PRESETS[names[0]], S.PRESETS[names[1]]
    x = torch.from_numpy(np.concatenate([corpus] * 14)[:331]).cuda()
    for n in (1, 5, 331):
        for layout in ("ft", "tf"):
            n0 = S.launch_count()
            ys, yf = S.spec_fbank(x[:n], ps, pf, layout=layout)
            assert S.launch_count() - n0 == 1
            assert torch.equal(ys, S.spec(x[:n], ps, layout=layout)) and torch.equal(yf, S.fbank(x[:n], pf))
    ys16, yf16 = S.spec_fbank(x.to(torch.int16), ps, pf)
    assert torch.equal(ys16, S.spec(x, ps)) and torch.equal(yf16, S.fbank(x, pf))
    y1s, y1f = S.spec_fbank(x[3], ps, pf)                                   # single clip
    assert torch.equal(y1s, S.spec(x[3], ps)) and torch.equal(y1f, S.fbank(x[3], pf))
    with pytest.raises(RuntimeError, match="SRFE_ERR_UNSUPPORTED"):
        S.spec_fbank(x[:4], ps, replace(pf, nfft=640, frame_len=640))
    with pytest.raises(TypeError):
        S.spec_fbank(x[:4].cpu(), ps, pf)


def test_spectrogram_with_tma_staged_frames(srfe_lib, corpus):
    """srfe_spec_staged_kernel (cp.async.bulk + mbarrier prefetch of each half-warp's next frame pair; opt-in, see
    srfe_abi.cu for why it is not the default): bit-identical to the direct-load kernel."""
    x = torch.from_numpy(np.concatenate([corpus] * 14)[:333]).cuda()
    try:
        for p, xx in ((S.C_SPEC, x), (replace(S.C_SPEC, layout="tf"), x), (replace(S.R_SPEC, layout="tf"), x.to(torch.int16)),
                      (replace(S.C_SPEC, log=False), x.to(torch.int16))):
            for n in (1, 7, 148, 333):
                S.set_tuning(stage=1, warps=16, ctas=1)
                direct = S.spec(xx[:n], p)
                S.set_tuning(stage=2, warps=16, ctas=1)
                assert torch.equal(S.spec(xx[:n], p), direct), (p, n)
        S.set_tuning(stage=2, warps=16, ctas=1)
        with pytest.raises(RuntimeError, match="SRFE_ERR_UNSUPPORTED"):        # fp32 R-SPEC: 32 x 5 KB of staging do not fit
            S.spec(x[:8], S.R_SPEC)
    finally:
        S.set_tuning()


def test_upload_through_the_staging_ring(srfe_lib, corpus):
    """srfe_upload / features.to_device: the reference forward's `x.to(DEVICE)` for pageable, pinned, strided and int16
    batches; bytes arrive unchanged, stream-ordered."""
    x = torch.from_numpy(np.concatenate([corpus] * 30))                      # 720 clips = 46 MB: several staging chunks
    assert not x.is_pinned()
    for src in (x, x.pin_memory(), x[:, :8000], x[::2], x.to(torch.int16), x[:1], x[:0]):
        got = S.to_device(src)
        assert got.is_cuda and got.dtype == src.dtype and got.shape == src.shape and got.is_contiguous()
        assert torch.equal(got.cpu(), src)
    y = S.to_device(x)
    assert torch.equal(S.mfcc(y), S.mfcc(x.cuda()))
    assert S.to_device(y) is y
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        z = S.to_device(x)
        f = S.mfcc(z)
    st.synchronize()
    assert torch.equal(f, S.mfcc(y))
