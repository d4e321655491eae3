#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference functions
(TEST INFRASTRUCTURE ONLY; runs in the build container where /root/reference is
mounted -- the GPU box never runs this).

    python oracle/make_golden.py            # writes tests/golden/

What is pinned:
  * spec  : models/model_spec_bgru.py::compute_spec  and model_spec_cnn.py::compute_spec
  * fbank : models/model_fbanks_cnn.py::filter_banks
  * mfcc  : models/model_mfcc_bgru.py::compute_mfcc imported through
            oracle/librosa_shim.py (librosa itself is absent -> "restated
            librosa"); plus independent cross-checks of the restated chain
            against transformers.audio_utils and torchaudio for R-MFCC and C-MFCC
            (oracle/crosscheck.py): recomputed live by tests/test_mfcc_crosscheck.py,
            a copy recorded in tests/golden/mfcc_crosscheck.json.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import corpus, features, librosa_shim  # noqa: E402

REF = os.environ.get("SRFE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name: str):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, "models", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_inputs() -> tuple[list[str], np.ndarray]:
    names, clips = [], []
    c = corpus.synthetic_corpus(4, config_index=0)
    for i in range(c.shape[0]):
        names.append(f"corpus0_{i}")
        clips.append(c[i])
    # one of each silence flavour of the corpus definition
    names.append("corpus0_88"); clips.append(corpus.synthetic_corpus(1, 0, start=88)[0])
    names.append("corpus0_96"); clips.append(corpus.synthetic_corpus(1, 0, start=96)[0])
    for k, v in corpus.edge_suite().items():
        names.append(k)
        clips.append(v)
    return names, np.stack(clips).astype(np.float32)


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    try:                                    # import before the shim: transformers probes librosa.__spec__
        from transformers import audio_utils  # noqa: F401
    except Exception:                       # pragma: no cover
        pass
    librosa_shim.install()
    m_spec_bgru = _load("model_spec_bgru")
    m_spec_cnn = _load("model_spec_cnn")
    m_fbank = _load("model_fbanks_cnn")
    m_mfcc = _load("model_mfcc_bgru")

    names, x = golden_inputs()
    tx = torch.from_numpy(x)
    spec_ft = np.stack([m_spec_bgru.compute_spec(tx[i]).numpy() for i in range(len(names))])
    spec_tf = np.stack([m_spec_cnn.compute_spec(tx[i]).numpy() for i in range(len(names))])
    fbank = np.stack([m_fbank.filter_banks(tx[i]).numpy() for i in range(len(names))])
    mfcc = np.stack([m_mfcc.compute_mfcc(tx[i]).numpy() for i in range(len(names))])
    assert spec_ft.shape[1:] == (321, 49) and spec_tf.shape[1:] == (49, 321)
    assert fbank.shape[1:] == (98, 120) and mfcc.shape[1:] == (39, 51)

    np.savez_compressed(os.path.join(OUT, "reference_features.npz"),
                        names=np.array(names), x=x,
                        spec_ft=spec_ft, spec_tf=spec_tf, fbank=fbank, mfcc=mfcc)

    # ---- independent cross-checks of the restated librosa chain (recorded copy; the LIVE checks are
    #      tests/test_mfcc_crosscheck.py) ----------------------------------------------------------
    from dataclasses import replace
    from oracle import crosscheck
    report = {"note": "max abs difference of oracle.mfcc_truth (static coefficients) vs independent implementations; "
                      "informational record -- tests/test_mfcc_crosscheck.py recomputes these on every run",
              "clips": names}
    for preset in ("R-MFCC", "C-MFCC"):
        p = replace(features.PRESETS[preset], n_deltas=0)
        ours = [features.mfcc_truth(x[i].astype(np.float64), p) for i in range(len(names))]
        rep = {"mel_matrix_vs_transformers": float(np.abs(crosscheck.hf_mel_matrix(p).T - features.slaney_mel_filterbank(
            p.sr, p.n_fft, p.n_mels, p.fmin, p.f_hi)).max())}
        for key, fn in (("mfcc_vs_transformers_audio_utils", crosscheck.hf_mfcc),
                        ("mfcc_vs_torchaudio_f64_with_f32_tables", crosscheck.torchaudio_mfcc)):
            try:
                rep[key] = {n: float(np.abs(fn(x[i].astype(np.float64), p) - ours[i]).max()) for i, n in enumerate(names)}
            except Exception as e:  # pragma: no cover
                rep[key + "_error"] = repr(e)
        report[preset] = rep
    with open(os.path.join(OUT, "mfcc_crosscheck.json"), "w") as f:
        json.dump(report, f, indent=1)

    # ---- analytic known answers --------------------
    kat = {
        "zeros_spec": float(np.log(np.float32(1e-10))),
        "zeros_fbank": float(20 * np.log10(np.finfo(float).eps)),
        "zeros_mfcc_c0": float(-100.0 * np.sqrt(128.0)),
        "fbank_empty_filters": [0, 2, 4, 7, 9, 11, 14, 17, 21, 25],
        "tone1k_bin40_psd": 2.0 * (1000.0 * features.tukey_periodic(640).sum() / 2.0) ** 2
                            / (16000.0 * (features.tukey_periodic(640) ** 2).sum()),
        "gradient_in": [1, 4, 9, 16, 25], "gradient_d1": [3, 4, 6, 8, 9], "gradient_d2": [1, 1.5, 2, 1.5, 1],
    }
    with open(os.path.join(OUT, "known_answers.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", OUT, {k: v.shape for k, v in dict(spec_ft=spec_ft, spec_tf=spec_tf, fbank=fbank, mfcc=mfcc).items()})
    print(json.dumps({k: v for k, v in report.items() if k != "clips"}, indent=1))


if __name__ == "__main__":
    main()
