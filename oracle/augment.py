"""CPU restatement of the reference's training-time PCM augmentation (TEST INFRASTRUCTURE ONLY), with the
random draws taken from a COUNTER-BASED stream so that a device implementation can reproduce them bit for bit.

What is restated (paths relative to /root/reference):

* ``Dataset.generate_silence_sample``  dataset.py:148-161  zeros, or a 1-s slice of a background-noise file x U(0,1)
* ``Dataset.add_noise_uniform``        dataset.py:185-191  np.int16(sample + U(0, upper) * noise_slice)
* ``Dataset.add_noise_snr``            dataset.py:163-183  noise slice scaled to an SNR drawn from [-5, 0, 5, 10, None] dB
* ``Dataset.time_stretching``          dataset.py:193-202  shift by randint(-range, range), fill with randint(-32, 32)
* the band selection of ``__getitem__`` dataset.py:107-116 (one uniform draw per clip; 0.4..0.6 -> shift, 0.6..0.8 -> noise;
  the 0..0.2 pitch-shift and 0.2..0.4 speed-tune bands are resampling / phase-vocoder ops that stay on the host: such clips
  pass through unchanged and are flagged)

What differs from the reference, by construction: ``random.randint`` / ``np.random.*`` (global Mersenne-Twister state, call
order dependent) are replaced by Philox4x32-10 keyed by (seed) and counted by (global clip index, draw block) -- the
"counter-based RNG contract" SURVEY.md 8(f3) asks for; the SNR powers are exact integer sums of squares instead of numpy's
float64 pairwise sums (<= 1 ulp apart in the scale factor).  The arithmetic on the samples follows the reference's dtypes:
float64 multiply then add, np.int16() truncation toward zero with two's-complement wrap, float32 at the end
(dataset.py:117).  The reference ships no test for any of this: parity is pinned to this restatement ("unpinned").

Draw layout, clip with global index g (counter = (g & 0xffffffff, g >> 32, block, 0), key = (seed & 0xffffffff, seed >> 32)):
  block 0: r0 -> op band draw u = (r0 >> 8) 2^-24          r1 -> shift = -range + (r1 (2 range + 1) >> 32)  |  noise file = r1 n_files >> 32
           r2 -> slice start = r2 (len - N + 1) >> 32       r3 -> scale u' = (r3 >> 8) 2^-24
  block 1: s0 -> SNR level = s0 5 >> 32
  block 2 + j // 4, word j % 4 -> fill sample j of a shifted clip: (w & 63) - 32
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

OP_NONE, OP_SHIFT, OP_NOISE_UNIFORM, OP_NOISE_SNR, OP_SILENCE_ZERO, OP_SILENCE_NOISE, OP_HOST_PITCH, OP_HOST_SPEED = range(8)
KIND_CLIP, KIND_SILENCE_ZERO, KIND_SILENCE_NOISE = 0, 1, 2
SNR_LEVELS_DB = (-5.0, 0.0, 5.0, 10.0)          # + None (dataset.py:172)

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11).  counter [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(_W0)) & mask
        k1 = (k1 + np.uint64(_W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def _draw(seed: int, g: int, block) -> np.ndarray:
    block = np.atleast_1d(np.asarray(block, dtype=np.uint64))
    ctr = np.zeros((block.size, 4), dtype=np.uint32)
    ctr[:, 0] = g & 0xFFFFFFFF
    ctr[:, 1] = (g >> 32) & 0xFFFFFFFF
    ctr[:, 2] = block
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, np.broadcast_to(key, (block.size, 2)))


def _u01(r: int) -> float:
    return float(np.float32((int(r) >> 8) * 2.0 ** -24))          # exact in float32


def _scaled(r: int, span: int) -> int:
    return (int(r) * int(span)) >> 32


def _to_int16(v: np.ndarray) -> np.ndarray:
    """np.int16(float64 array) as x86 numpy does it: truncate toward zero, keep the low 16 bits."""
    return np.trunc(v).astype(np.int64).astype(np.int16)


@dataclass(frozen=True)
class AugmentParams:
    seed: int = 0
    shift_band: tuple = (0.4, 0.6)          # dataset.py:112-113
    noise_band: tuple = (0.6, 0.8)          # dataset.py:114-115
    snr_band: tuple = (0.0, 0.0)            # add_noise_snr exists (dataset.py:163-183) but __getitem__ never calls it: off
    pitch_band: tuple = (0.0, 0.2)          # host-only ops: reported, clip left unchanged
    speed_band: tuple = (0.2, 0.4)
    shift_range: int = 4800                 # dataset.py:113
    noise_upper: float = 0.1                # dataset.py:115


def snr_divisors() -> np.ndarray:
    """10 ** (snr / 10) for the four finite levels, float64 (shared with the device through the parameter block)."""
    return np.array([10.0 ** (s / 10.0) for s in SNR_LEVELS_DB], dtype=np.float64)


def augment_ref(pcm: np.ndarray, noise_files: list, p: AugmentParams = AugmentParams(), kind=None,
                first_index: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """pcm int16 [B, N] -> (float32 [B, N], op codes int8 [B])."""
    pcm = np.asarray(pcm)
    assert pcm.dtype == np.int16 and pcm.ndim == 2
    B, N = pcm.shape
    kind = np.zeros(B, np.int8) if kind is None else np.asarray(kind, np.int8)
    out = np.empty((B, N), np.float32)
    ops = np.zeros(B, np.int8)
    div = snr_divisors()
    band = lambda u, b: np.float32(b[0]) <= np.float32(u) < np.float32(b[1])
    for i in range(B):
        g = first_index + i
        r = _draw(p.seed, g, 0)[0]
        s = pcm[i]

        def noise_slice():
            f = _scaled(r[1], len(noise_files))
            nz = np.asarray(noise_files[f], dtype=np.int16)
            start = _scaled(r[2], len(nz) - N + 1)
            return nz[start:start + N]

        if kind[i] == KIND_SILENCE_ZERO:                               # dataset.py:152-154
            out[i], ops[i] = 0.0, OP_SILENCE_ZERO
        elif kind[i] == KIND_SILENCE_NOISE:                            # dataset.py:155-160
            out[i] = (noise_slice().astype(np.float64) * np.float64(_u01(r[3]))).astype(np.float32)
            ops[i] = OP_SILENCE_NOISE
        else:
            u = _u01(r[0])
            if band(u, p.shift_band):                                  # dataset.py:193-202
                shift = -p.shift_range + _scaled(r[1], 2 * p.shift_range + 1)
                n_fill = abs(shift)
                j = np.arange(n_fill)
                words = _draw(p.seed, g, 2 + j // 4)[np.arange(n_fill), j % 4] if n_fill else np.zeros(0, np.uint32)
                fill = (words & 63).astype(np.int64) - 32
                y = np.concatenate((s[shift:], fill)) if shift >= 0 else np.concatenate((fill, s[:shift]))
                out[i], ops[i] = np.int16(y).astype(np.float32), OP_SHIFT
            elif band(u, p.noise_band):                                # dataset.py:185-191
                f = np.float64(_u01(r[3])) * np.float64(np.float32(p.noise_upper))
                y = s.astype(np.float64) + f * noise_slice().astype(np.float64)
                out[i], ops[i] = _to_int16(y).astype(np.float32), OP_NOISE_UNIFORM
            elif band(u, p.snr_band):                                  # dataset.py:163-183
                nz = noise_slice()
                level = _scaled(_draw(p.seed, g, 1)[0][0], 5)
                ps = int((s.astype(np.int64) ** 2).sum())
                pn = int((nz.astype(np.int64) ** 2).sum())
                if level == 4 or pn == 0:                              # None level; (a silent noise slice: left unchanged)
                    out[i] = s.astype(np.float32)
                else:
                    sp = np.float64(ps) / np.float64(2.0 ** 30) / np.float64(N)
                    npw = np.float64(pn) / np.float64(2.0 ** 30) / np.float64(N)
                    f = np.sqrt((sp / npw) / div[level])
                    out[i] = _to_int16(s.astype(np.float64) + f * nz.astype(np.float64)).astype(np.float32)
                ops[i] = OP_NOISE_SNR
            else:
                out[i] = s.astype(np.float32)                          # dataset.py:117
                ops[i] = OP_HOST_PITCH if band(u, p.pitch_band) else OP_HOST_SPEED if band(u, p.speed_band) else OP_NONE
    return out, ops
