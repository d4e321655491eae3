"""GPU: one parity report over every preset and the edge suite, with the figures the level-aware tolerances leave
out of the assertion made visible and bounded ("report the max / fraction outside"):

* log-mel / log-spec: max abs error INSIDE the tolerance domain (asserted <= 1e-3), the fraction of elements outside
  it and the max abs error there (reported; bounded loosely so that a regression shows);
* raw power spectra: relative error with the floored denominator (asserted <= 1e-4) on the corpus AND on the edge
  suite (tonal / full-scale / impulsive clips);
* MFCC: max abs error on every element (asserted <= 1e-3).

The report is written to gpurun_out/parity_report.json when that directory exists (the GPU box), so that the numbers
can be committed under profiles/.
"""
from __future__ import annotations

import json
import os
from dataclasses import replace

import numpy as np
import pytest
import torch

import oracle
import speechrecognitionproject_b200 as S
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPORT: dict = {}


def _gpu(fn, x, p):
    y = fn(torch.from_numpy(x).cuda(), p)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def _sets():
    e = oracle.edge_suite()
    return {"corpus": (None, oracle.synthetic_corpus(32, config_index=2)),
            "edge": (list(e.keys()), np.stack(list(e.values())))}


@pytest.mark.parametrize("name", ["R-SPEC", "C-SPEC"])
def test_report_spec(srfe_lib, name):
    p, op = S.PRESETS[name], oracle.PRESETS[name]
    for sname, (names, x) in _sets().items():
        got = _gpu(S.spec, x, p)
        truth = H.oracle_batch(oracle.spec_truth, x, op)
        st = H.check_logspec(got, truth, f"{name} {sname}")
        raw = _gpu(S.spec, x, replace(p, log=False))
        traw = H.oracle_batch(oracle.spec_truth, x, replace(op, log=False))
        st["psd_rel_err_floored"] = H.check_psd(raw, traw, f"{name} {sname} raw PSD")
        # unfloored relative error, bins within 50 dB of the clip maximum only (reported)
        t2, g2 = traw.reshape(len(x), -1), raw.reshape(len(x), -1).astype(np.float64)
        m = t2 >= 1e-5 * t2.max(axis=1, keepdims=True)
        m &= t2 > 0
        st["psd_rel_err_raw_within_50dB"] = float((np.abs(g2 - t2)[m] / t2[m]).max()) if m.any() else 0.0
        # outside the 50 dB domain the log output is dominated by the eps floor and single-precision leakage -- for the
        # reference's own scipy float32 path too (reported next to ours for scale); sanity bound only
        ref = H.oracle_batch(oracle.spec_ref, x, op).astype(np.float64).reshape(len(x), -1)
        tl = truth.reshape(len(x), -1)
        outside = tl < tl.max(axis=1, keepdims=True) - 11.5
        st["reference_scipy_f32_max_err_outside"] = float(np.abs(ref - tl)[outside].max()) if outside.any() else 0.0
        assert st["max_err_outside"] < 10.0, st
        _REPORT[f"{name}/{sname}"] = st


@pytest.mark.parametrize("name", ["R-FBANK", "C-FBANK"])
def test_report_fbank(srfe_lib, name):
    p, op = S.PRESETS[name], oracle.PRESETS[name]
    for sname, (names, x) in _sets().items():
        got = _gpu(S.fbank, x, p)
        truth = H.oracle_batch(oracle.fbank_truth, x, op)
        st = H.check_logmel(got, truth, f"{name} {sname}")
        # what the reference's own float32 pre-emphasis + float64 chain does out there, for scale
        ref = H.oracle_batch(oracle.fbank_ref, x, op).astype(np.float64)
        t2 = truth.reshape(len(x), -1)
        outside = t2 < t2.max(axis=1, keepdims=True) - 100.0
        st["reference_dtype_path_max_err_outside"] = float(np.abs(ref.reshape(len(x), -1) - t2)[outside].max()) if outside.any() else 0.0
        _REPORT[f"{name}/{sname}"] = st


@pytest.mark.parametrize("name", ["R-MFCC", "C-MFCC", "C-MFCC-D2"])
def test_report_mfcc(srfe_lib, name):
    p, op = S.PRESETS[name], oracle.PRESETS[name]
    for sname, (names, x) in _sets().items():
        got = _gpu(S.mfcc, x, p)
        truth = H.oracle_batch(oracle.mfcc_truth, x, op)
        err = H.check_mfcc(got, truth, f"{name} {sname}")
        _REPORT[f"{name}/{sname}"] = {"max_abs_err_all_elements": err}


def test_write_report():
    assert len(_REPORT) == 14, sorted(_REPORT)
    print("\nPARITY REPORT " + json.dumps(_REPORT, sort_keys=True))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
