"""CPU: lane-by-lane emulation of the packed (two frames per lane) half-warp FFT -- the same
__host__ __device__ templates the kernels use -- against numpy.  Catches permutation / twiddle /
pairing mistakes without a GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

VP = C.c_void_p


def _p(a):
    return a.ctypes.data_as(VP)


@pytest.mark.parametrize("n", [4, 5, 16, 20])
def test_codelets(emu_lib, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, 2)).astype(np.float32)
    b = rng.standard_normal((n, 2)).astype(np.float32)
    oa, ob = np.zeros_like(a), np.zeros_like(b)
    assert emu_lib.emu_dft2(n, _p(a), _p(b), _p(oa), _p(ob)) == 0
    for x, o in ((a, oa), (b, ob)):
        ref = np.fft.fft(x[:, 0].astype(np.float64) + 1j * x[:, 1])
        assert np.abs((o[:, 0] + 1j * o[:, 1]) - ref).max() < 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("nfft", [512, 640])
def test_real_fft_power_pairs(emu_lib, nfft):
    rng = np.random.default_rng(nfft)
    frames = []
    frames.append((rng.standard_normal(nfft) * 3000).astype(np.float32))
    f = (rng.standard_normal(nfft) * 300).astype(np.float32); f[400:] = 0; frames.append(f)      # 400-in-512 zero padding
    f = np.zeros(nfft, np.float32); f[7] = 1000; frames.append(f)                                 # impulse: flat spectrum
    frames.append(np.full(nfft, 1234.0, np.float32))                                              # DC
    frames.append(np.zeros(nfft, np.float32))                                                     # silence
    for i in range(len(frames)):
        xa, xb = frames[i], frames[(i + 1) % len(frames)]      # different content in the two packed lanes
        pa = np.zeros(nfft // 2 + 1, np.float32)
        pb = np.zeros_like(pa)
        assert emu_lib.emu_power2(nfft, _p(xa), _p(xb), _p(pa), _p(pb)) == 0
        for x, p in ((xa, pa), (xb, pb)):
            ref = np.abs(np.fft.rfft(x.astype(np.float64))) ** 2
            assert np.abs(p - ref).max() <= 2e-6 * max(ref.max(), 1e-30)


@pytest.mark.parametrize("name", ["R-FBANK", "C-FBANK", "R-MFCC", "C-MFCC", "generic-fbank-26", "generic-mfcc-80"])
def test_ell_filterbank_tables(emu_lib, name):
    """Host table logic of the mel / filterbank stage (srfe_tables.cpp: to_sparse -> to_ell): walking the planar,
    bank-skewed ELL tables the way the kernel does reproduces the dense matrix product, padded reads carry zero weights,
    the runs stay inside the per-half-warp power buffer, and at most two of the 16 filters of a group start in the same
    bank pair for the presets (the skew's purpose; unskewed banks have 4- to 8-way groups)."""
    import ctypes as C
    import speechrecognitionproject_b200 as S
    from speechrecognitionproject_b200 import _lib
    presets = {"R-FBANK": S.R_FBANK, "C-FBANK": S.C_FBANK, "R-MFCC": S.R_MFCC, "C-MFCC": S.C_MFCC,
               "generic-fbank-26": S.FbankParams(nfilt=26), "generic-mfcc-80": S.MfccParams(n_fft=512, win_length=400, hop=160, n_mels=80, n_mfcc=13)}
    p = presets[name]
    is_mfcc = isinstance(p, S.MfccParams)
    lib = _lib.lib()
    n_fft = p.n_fft if is_mfcc else p.nfft
    n_bins, n_filt = n_fft // 2 + 1, (p.n_mels if is_mfcc else p.nfilt)
    dense = np.zeros((n_filt, n_bins))
    pc = p.to_c()
    fn = lib.srfe_mfcc_filters_f64 if is_mfcc else lib.srfe_fbank_filters_f64
    assert fn(C.byref(pc), dense.ctypes.data_as(C.POINTER(C.c_double))) == 0
    rng = np.random.default_rng(7)
    power = (rng.random(n_bins) * 1e6).astype(np.float32)
    out = np.zeros(n_filt)
    reach = C.c_int(0)
    emu = emu_lib.emu_ell_project_mfcc if is_mfcc else emu_lib.emu_ell_project_fbank
    emu.restype = C.c_int
    worst = emu(C.byref(pc), power.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(reach))
    assert worst >= 1, f"emulator error {worst}"
    want = dense @ power.astype(np.float64)
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=1e-3)          # weights are stored as float32
    assert reach.value <= (272 if n_fft == 512 else 400)                 # FftGeom<N>::SCRATCH_P2 slots
    if not name.startswith("generic"):
        assert worst <= 2, f"{name}: {worst} filters of a group share a bank pair"
