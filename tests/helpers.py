"""Comparison helpers shared by the parity tests (see tests/tolerances.py)."""
from __future__ import annotations

import numpy as np

from tests import tolerances as tol


def to_oracle_params(p):
    """product dataclass -> oracle dataclass (same field names by construction)."""
    import oracle
    cls = {"SpecParams": oracle.SpecParams, "FbankParams": oracle.FbankParams, "MfccParams": oracle.MfccParams}[type(p).__name__]
    return cls(**{k: getattr(p, k) for k in cls.__dataclass_fields__})


def oracle_batch(fn, x: np.ndarray, p) -> np.ndarray:
    return np.stack([fn(x[i], p) for i in range(x.shape[0])])


def check_mfcc(got: np.ndarray, truth: np.ndarray, what: str = "") -> float:
    err = float(np.abs(got.astype(np.float64) - truth).max()) if got.size else 0.0
    assert err <= tol.MFCC_ABS, f"{what}: MFCC max abs err {err:.3e} > {tol.MFCC_ABS}"
    return err


def check_logmel(got: np.ndarray, truth: np.ndarray, what: str = "", domain: float = tol.LOGMEL_DOMAIN,
                 atol: float = tol.LOGMEL_ABS) -> dict:
    """got/truth: [B, T, nfilt] (or any [B, ...]); level-aware per clip."""
    g = got.astype(np.float64).reshape(got.shape[0], -1)
    r = truth.reshape(truth.shape[0], -1)
    clip_max = r.max(axis=1, keepdims=True)
    inside = r >= clip_max - domain
    err = np.abs(g - r)
    e_in = float(err[inside].max()) if inside.any() else 0.0
    stats = {"max_err_in_domain": e_in, "frac_outside": float(1.0 - inside.mean()),
             "max_err_outside": float(err[~inside].max()) if (~inside).any() else 0.0}
    assert e_in <= atol, f"{what}: log-mel max abs err in domain {e_in:.3e} > {atol} ({stats})"
    return stats


def check_psd(got: np.ndarray, truth: np.ndarray, what: str = "") -> float:
    g = got.astype(np.float64).reshape(got.shape[0], -1)
    r = truth.reshape(truth.shape[0], -1)
    clip_max = r.max(axis=1, keepdims=True)
    den = np.maximum(np.abs(r), tol.PSD_FLOOR * clip_max)
    den = np.where(den == 0, 1.0, den)
    rel = float((np.abs(g - r) / den).max())
    assert rel <= tol.PSD_REL, f"{what}: PSD max rel err {rel:.3e} > {tol.PSD_REL}"
    return rel


def check_logspec(got: np.ndarray, truth: np.ndarray, what: str = "") -> dict:
    return check_logmel(got, truth, what, domain=tol.LOGSPEC_DOMAIN_NEPER, atol=tol.LOGSPEC_ABS)
