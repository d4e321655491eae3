"""pytest configuration: the ``gpu`` marker, in-tree builds, shared helpers."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def srfe_lib():
    """libsrfe.so, built in-tree if missing/stale (nvcc cross-compiles without a GPU)."""
    from speechrecognitionproject_b200 import _lib, build
    try:
        build.build()
    except Exception:
        if not os.path.exists(_lib.LIB_PATH):
            raise
    return _lib.lib()


@pytest.fixture(scope="session")
def emu_lib():
    """CPU lane-by-lane emulation of the half-warp FFT (tests/emu)."""
    import ctypes
    out_dir = os.path.join(ROOT, "build", "emu")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libfft_emu.so")
    srcs = [os.path.join(ROOT, "tests", "emu", "fft_emu.cpp"),
            os.path.join(ROOT, "speechrecognitionproject_b200", "csrc", "srfe_tables.cpp")]
    deps = srcs + [os.path.join(ROOT, "speechrecognitionproject_b200", "csrc", f) for f in ("srfe_fft.cuh", "srfe_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{cuda_inc}"] + srcs + ["-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_features.npz"))
    return {k: g[k] for k in g.files}
