"""On-device training-time augmentation and silence synthesis, the host-side mirror of what the
reference's ``Dataset`` does per clip on the CPU (/root/reference/dataset.py):

* ``generate_silence_sample``  :148-161   zeros for the first 185 requests, then a background slice x U(0, 1)
* ``add_noise_uniform``        :185-191   np.int16(sample + U(0, upper) * noise slice)
* ``add_noise_snr``            :163-183   noise slice scaled to an SNR drawn from [-5, 0, 5, 10, None] dB
* ``time_stretching``          :193-202   shift by randint(-range, range), ends filled with randint(-32, 32)
* the band selection of ``__getitem__`` :107-116

A whole batch of int16 clips is processed by ONE kernel launch (``srfe_augment_i16``) and the float32 result stays on
the device, ready for ``features.mfcc / spec / fbank`` on the same stream.  The reference's global Mersenne-Twister state is
replaced by a counter-based contract -- Philox4x32-10, key = seed, counter = (global clip index, draw block) -- so a batch
augments identically however it is sharded across GPUs, and the numpy oracle (oracle/augment.py) reproduces the kernel bit
for bit.  The pitch-shift / speed-tune bands (librosa phase vocoder, cv2 resampling) stay host-side: such clips come back
unchanged with their op code set, for a caller that wants to route them through the CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

__all__ = ["AugmentParams", "NoiseBank", "DeviceAugmenter", "augment", "OPS"]

OPS = ("none", "shift", "noise_uniform", "noise_snr", "silence_zero", "silence_noise", "host_pitch", "host_speed")
KIND_CLIP, KIND_SILENCE_ZERO, KIND_SILENCE_NOISE = 0, 1, 2


@dataclass(frozen=True)
class AugmentParams:
    """Bands of the per-clip uniform draw (dataset.py:107-116 literals, lifted) and the op parameters."""
    seed: int = 0
    shift_band: tuple = (0.4, 0.6)
    noise_band: tuple = (0.6, 0.8)
    snr_band: tuple = (0.0, 0.0)          # add_noise_snr is defined by the reference but never called: off by default
    pitch_band: tuple = (0.0, 0.2)        # host-only ops: flagged, not applied
    speed_band: tuple = (0.2, 0.4)
    shift_range: int = 4800
    noise_upper: float = 0.1

    def to_c(self) -> _lib.AugmentParamsC:
        return _lib.AugmentParamsC(self.seed & 0xFFFFFFFFFFFFFFFF, *self.shift_band, *self.noise_band, *self.snr_band,
                                   *self.pitch_band, *self.speed_band, self.shift_range, self.noise_upper)


class NoiseBank:
    """The reference's ``_background_noise_`` wavs (dataset.py:61-64), resident on the device: int16 samples of all files
    back to back plus an offsets table.  Every file must be at least one clip long (dataset.py:158)."""

    def __init__(self, files: Sequence, device="cuda"):
        waves = [torch.as_tensor(f).to(torch.int16).flatten() for f in files]
        if not waves:
            raise ValueError("NoiseBank needs at least one background-noise file")
        self.lengths = [int(w.numel()) for w in waves]
        off = [0]
        for n in self.lengths:
            off.append(off[-1] + n)
        self.samples = torch.cat(waves).to(device)
        self.offsets = torch.tensor(off, dtype=torch.int64, device=device)
        self.n_files = len(waves)


def augment(pcm: torch.Tensor, bank: Optional[NoiseBank], params: AugmentParams = AugmentParams(), *,
            kind: Optional[torch.Tensor] = None, first_index: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """``pcm`` int16 CUDA ``[B, N]`` -> (float32 ``[B, N]`` on the same device, int8 op codes ``[B]``, see ``OPS``).

    ``kind`` (int8 ``[B]``): 0 = an ordinary clip (augmented according to its draw), 1 = silence of zeros, 2 = silence from
    a background slice.  ``first_index``: global index of row 0, the RNG counter -- shards of one batch pass their offset."""
    if not pcm.is_cuda or pcm.dtype != torch.int16 or pcm.dim() != 2:
        raise TypeError("augment: expected an int16 CUDA tensor [n_clips, n_samples] (the wav's own type, dataset.py:103)")
    if pcm.stride(1) != 1 or (pcm.size(0) > 1 and pcm.stride(0) < pcm.size(1)):
        pcm = pcm.contiguous()
    B, N = pcm.shape
    if bank is not None and min(bank.lengths) < N:
        raise ValueError("augment: every background-noise file must be at least n_samples long")
    if kind is not None:
        kind = kind.to(device=pcm.device, dtype=torch.int8).contiguous()
        if kind.numel() != B:
            raise ValueError("augment: kind must have one entry per clip")
    out = torch.empty((B, N), dtype=torch.float32, device=pcm.device)
    ops = torch.empty((B,), dtype=torch.int8, device=pcm.device)
    cp = params.to_c()
    with torch.cuda.device(pcm.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().srfe_augment_i16(
            pcm.data_ptr(), B, N, pcm.stride(0) if B > 1 else N, kind.data_ptr() if kind is not None else None, first_index,
            bank.samples.data_ptr() if bank is not None else None, bank.offsets.data_ptr() if bank is not None else None,
            bank.n_files if bank is not None else 0, C.byref(cp), out.data_ptr(), ops.data_ptr(), stream))
    return out, ops


class DeviceAugmenter:
    """Batch counterparts of the reference Dataset's augmentation methods, same names (dataset.py:148-202).

    Keeps the two pieces of state the reference keeps: the zeros-silence counter (``silence_class_zeros_count``,
    dataset.py:152-154: the first 185 silence requests are digital silence) and, instead of the global RNG state, the
    running clip index that feeds the counter-based stream."""

    def __init__(self, bank: NoiseBank, seed: int = 0, params: Optional[AugmentParams] = None):
        self.bank = bank
        self.params = params or AugmentParams(seed=seed)
        self.next_index = 0
        self.silence_class_zeros_count = 0

    def _run(self, pcm, params, kind=None):
        out, ops = augment(pcm, self.bank, params, kind=kind, first_index=self.next_index)
        self.next_index += pcm.size(0)
        return out, ops

    def __call__(self, pcm: torch.Tensor, kind: Optional[torch.Tensor] = None):
        """training-mode ``__getitem__`` (dataset.py:107-117) for a whole batch"""
        return self._run(pcm, self.params, kind)

    def generate_silence_sample(self, n: int, n_samples: int = 16000) -> torch.Tensor:
        zeros = max(0, min(n, 185 - self.silence_class_zeros_count))
        self.silence_class_zeros_count += zeros
        kind = torch.full((n,), KIND_SILENCE_NOISE, dtype=torch.int8)
        kind[:zeros] = KIND_SILENCE_ZERO
        dummy = torch.zeros((n, n_samples), dtype=torch.int16, device=self.bank.samples.device)
        return self._run(dummy, self.params, kind)[0]

    def _only(self, **band):
        from dataclasses import replace
        off = dict(shift_band=(0.0, 0.0), noise_band=(0.0, 0.0), snr_band=(0.0, 0.0), pitch_band=(0.0, 0.0), speed_band=(0.0, 0.0))
        off.update(band)
        return replace(self.params, **off)

    def add_noise_uniform(self, pcm: torch.Tensor, upper_bound: float = 0.1) -> torch.Tensor:
        from dataclasses import replace
        return self._run(pcm, replace(self._only(noise_band=(0.0, 2.0)), noise_upper=upper_bound))[0]

    def add_noise_snr(self, pcm: torch.Tensor) -> torch.Tensor:
        return self._run(pcm, self._only(snr_band=(0.0, 2.0)))[0]

    def time_stretching(self, pcm: torch.Tensor, range: int = 4800) -> torch.Tensor:
        from dataclasses import replace
        return self._run(pcm, replace(self._only(shift_band=(0.0, 2.0)), shift_range=range))[0]
