"""CPU: lane-by-lane emulation of the half-warp FFT (the same __host__ __device__ templates
the kernels use) against numpy -- catches permutation / twiddle mistakes without a GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

VP = C.c_void_p


@pytest.mark.parametrize("n", [4, 5, 16, 20])
def test_codelets(emu_lib, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, 2)).astype(np.float32)
    o = np.zeros_like(a)
    assert emu_lib.emu_dft(n, a.ctypes.data_as(VP), o.ctypes.data_as(VP)) == 0
    ref = np.fft.fft(a[:, 0].astype(np.float64) + 1j * a[:, 1])
    assert np.abs((o[:, 0] + 1j * o[:, 1]) - ref).max() < 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("nfft", [512, 640])
def test_real_fft_power(emu_lib, nfft):
    rng = np.random.default_rng(nfft)
    for trial in range(4):
        xw = (rng.standard_normal(nfft) * 3000).astype(np.float32)
        if trial == 1:
            xw[400:] = 0                       # 400-in-512 zero padding
        if trial == 2:
            xw[:] = 0; xw[7] = 1000            # impulse: flat spectrum
        if trial == 3:
            xw[:] = 1234.0                     # DC
        p = np.zeros(nfft // 2 + 1, np.float32)
        assert emu_lib.emu_power(nfft, xw.ctypes.data_as(VP), p.ctypes.data_as(VP)) == 0
        ref = np.abs(np.fft.rfft(xw.astype(np.float64))) ** 2
        assert np.abs(p - ref).max() <= 2e-6 * ref.max()
