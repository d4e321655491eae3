#!/usr/bin/env python
"""Golden logits of the UNMODIFIED reference models models/model_{mfcc_bgru,spec_bgru,spec_cnn,fbanks_cnn}.py (TEST INFRASTRUCTURE ONLY; runs in the
build container where /root/reference is mounted).

    python oracle/make_golden_logits.py      # writes tests/golden/<model>_logits.npz

torch.manual_seed(SEED); Network() (full size: GRU(39, 512, 2 layers, bidirectional) + Linear(1024, 12),
models/model_mfcc_bgru.py:23-26); Network.forward (:28-37, the per-clip CPU loop over compute_mfcc, librosa through
oracle/librosa_shim.py) on 8 clips of the seeded corpus.  Stored: the logits, and per-parameter checksums so that the
GPU box -- where the reference tree does not exist -- can prove that the weights it regenerates from the same seed are
the reference module's weights (tests/twins.py)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tests import twins  # noqa: E402

SEED, N_CLIPS, CONFIG_INDEX = 20260005, 8, 5


def param_checksums(net) -> dict:
    return {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in net.state_dict().items()}


def main() -> None:
    x = oracle.synthetic_corpus(N_CLIPS, config_index=CONFIG_INDEX)
    for name in twins.TWINNED:
        mod = twins.load_reference(name)
        torch.manual_seed(SEED)
        net = mod.Network().eval()
        with torch.no_grad():
            logits = net(torch.from_numpy(x)).numpy()
        cs = param_checksums(net)
        out = os.path.join(ROOT, "tests", "golden", f"{name}_logits.npz")
        assert np.array_equal(x, np.round(x)) and np.abs(x).max() <= 32767      # int16-valued: stored compactly
        np.savez_compressed(out, seed=SEED, n_clips=N_CLIPS, config_index=CONFIG_INDEX, logits=logits,
                            keys=np.array(list(cs.keys())), checksums=np.stack(list(cs.values())),
                            **({"clips_i16": x.astype(np.int16)} if name == "model_mfcc_bgru" else {}))
        print("wrote", out, logits.shape, float(np.abs(logits).max()))


if __name__ == "__main__":
    main()
