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
