"""In-tree build of ``libsrfe.so`` for sm_100a (explicit nvcc, no JIT cache).

    python -m speechrecognitionproject_b200.build [--force] [--verbose]

The shared object lands next to this file so that it travels with the source tree
(it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsrfe.so")
SOURCES = ["srfe_abi.cu", "srfe_tables.cpp"]
HEADERS = ["srfe_kernels.cuh", "srfe_mfcc_tc.cuh", "srfe_fbank_tc.cuh", "srfe_augment.cuh", "srfe_fft.cuh", "srfe_tables.h", os.path.join("..", "..", "include", "srfe.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC",
              # the CUDA runtime as a shared library (the one torch has already loaded, else the toolkit's): the artefact
              # then carries no copy of the runtime's entry-point table
              "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", OUT]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsrfe.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
