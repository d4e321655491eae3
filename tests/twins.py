"""Model modules for the drop-in tests and the cfg5 benchmark.

``load("model_mfcc_bgru")`` imports the UNMODIFIED reference module when the reference tree is reachable
(``$SRFE_REFERENCE`` or /root/reference -- the build container), through oracle/librosa_shim.py because librosa is not
installed.  /root/reference does not exist on the GPU box, and reference sources are never copied into this repo, so
there the *shape twin* below stands in: the same constructor calls in the same order as
models/model_mfcc_bgru.py:23-26, hence the same parameters bit for bit for a given ``torch.manual_seed`` (checked
against the real module in tests/test_patch.py::test_twin_equals_reference_module whenever the tree is mounted, and
against tests/golden/model_mfcc_bgru_logits.npz always), and the reference's own per-clip forward loop
(:28-37) for ``patch_model`` to replace.
"""
from __future__ import annotations

import importlib.util
import os
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("SRFE_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "model_mfcc_bgru.py"))


def load_reference(name: str):
    from oracle import librosa_shim
    librosa_shim.install()
    spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF_ROOT, "models", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def twin_model_mfcc_bgru(compute_mfcc=None):
    """models/model_mfcc_bgru.py shape twin (see module docstring)."""
    m = types.ModuleType("twin_model_mfcc_bgru")

    def _compute_mfcc(sample):
        import oracle
        return torch.from_numpy(oracle.mfcc_ref(sample.numpy()))

    class Network(nn.Module):
        def __init__(self, num_features=512, num_layers=2):
            super().__init__()
            self.gru = nn.GRU(39, hidden_size=num_features, num_layers=num_layers, bidirectional=True, batch_first=True)
            self.fc = nn.Linear(num_features * 2, 12)

        def forward(self, x):
            with torch.no_grad():
                inx = torch.ones(x.size(0), 39, 51)
                for i in range(x.size(0)):
                    inx[i, :, :] = m.compute_mfcc(x[i])
            inx = inx.to(next(self.parameters()).device)
            inx = torch.transpose(inx, 1, 2)
            inx, _ = self.gru(inx)
            return self.fc(inx[:, -1, :])

    m.compute_mfcc = compute_mfcc or _compute_mfcc
    m.Network = Network
    return m


def load(name: str = "model_mfcc_bgru"):
    """-> (module, "reference" | "twin")"""
    if reference_available():
        return load_reference(name), "reference"
    if name != "model_mfcc_bgru":
        raise FileNotFoundError(f"{name}: reference tree not mounted and no twin for this module")
    return twin_model_mfcc_bgru(), "twin"
