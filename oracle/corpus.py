"""Seeded synthetic PCM for parity tests and the CPU arm (TEST INFRASTRUCTURE ONLY).

The clips imitate what ``Dataset.__getitem__`` hands to the models
(/root/reference/dataset.py:103-117): float32 arrays of 16000 samples on the
int16 scale (no /32768 normalisation), with the dataset's two silence flavours
mixed in (all-zero clips, dataset.py:153; non-integer ``noise * U(0,1)`` clips,
dataset.py:159-160).  Definition: SURVEY.md section 8d.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 20260000
N_SAMPLES = 16000


def synthetic_corpus(n_clips: int, config_index: int = 0, n_samples: int = N_SAMPLES,
                     start: int = 0) -> np.ndarray:
    """float32 [n_clips, n_samples].  Clip ``i`` (global index ``start + i``) only
    depends on (config_index, global index), so shards of a corpus can be
    generated independently on every rank."""
    out = np.empty((n_clips, n_samples), dtype=np.float32)
    for i in range(n_clips):
        gi = start + i
        rng = np.random.default_rng([SEED_BASE + config_index, gi])
        sigma = np.exp(rng.uniform(np.log(30.0), np.log(8000.0)))
        noise = np.clip(sigma * rng.standard_normal(n_samples), -32768.0, 32767.0)
        if gi % 97 == 96:
            out[i] = 0.0                                   # digital silence
        elif gi % 89 == 88:
            out[i] = (np.round(noise) * rng.uniform()).astype(np.float32)   # non-integer floats
        else:
            out[i] = np.round(noise).astype(np.float32)
    return out


def edge_suite(n_samples: int = N_SAMPLES) -> dict[str, np.ndarray]:
    """Small tonal / degenerate clips, checked with the level-aware tolerance."""
    n = np.arange(n_samples, dtype=np.float64)
    rng = np.random.default_rng(SEED_BASE + 99)
    impulse = np.zeros(n_samples)
    impulse[n_samples // 3] = 20000.0
    chirp = 8000.0 * np.sin(2 * np.pi * (50.0 * n / 16000.0 + 0.5 * 7000.0 * (n / 16000.0) ** 2))
    suite = {
        "zeros": np.zeros(n_samples),
        "dc": np.full(n_samples, 1234.0),
        "impulse": impulse,
        "tone1k": 1000.0 * np.sin(2 * np.pi * 1000.0 * n / 16000.0),
        "square_fullscale": np.where((n // 40) % 2 == 0, 32767.0, -32768.0),
        "chirp": np.round(chirp),
        "quiet_noise": np.round(3.0 * rng.standard_normal(n_samples)),
        "speechlike": np.round(3000.0 * rng.standard_normal(n_samples)
                               * (0.05 + np.abs(np.sin(2 * np.pi * 3.0 * n / 16000.0)))),
        "half_silent": np.concatenate([np.zeros(n_samples // 2),
                                       np.round(2000.0 * rng.standard_normal(n_samples - n_samples // 2))]),
    }
    return {k: v.astype(np.float32) for k, v in suite.items()}
