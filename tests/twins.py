"""Model modules for the drop-in tests and the cfg5 benchmark.

``load("model_mfcc_bgru")`` imports the UNMODIFIED reference module when the reference tree is reachable
(``$SRFE_REFERENCE`` or /root/reference -- the build container), through oracle/librosa_shim.py because librosa is not
installed.  /root/reference does not exist on the GPU box, and reference sources are never copied into this repo, so
there *shape twins* stand in (model_mfcc_bgru, model_spec_bgru, model_spec_cnn, model_fbanks_cnn): the same constructor
calls in the same order as the reference's Network.__init__ (e.g. models/model_mfcc_bgru.py:23-26), hence the same parameters bit for bit for a given ``torch.manual_seed`` (checked
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


def _twin_bgru(name: str, n_in: int, feature_name: str, oracle_fn):
    """model_spec_bgru.py:18-35 / model_mfcc_bgru.py:21-37 shape twin"""
    m = types.ModuleType(f"twin_{name}")

    class Network(nn.Module):
        def __init__(self, num_features=512, num_layers=2):
            super().__init__()
            self.gru = nn.GRU(n_in, hidden_size=num_features, num_layers=num_layers, bidirectional=True, batch_first=True)
            self.fc = nn.Linear(num_features * 2, 12)

        def forward(self, x):
            with torch.no_grad():
                inx = torch.stack([getattr(m, feature_name)(x[i]) for i in range(x.size(0))])
            inx = torch.transpose(inx.to(next(self.parameters()).device), 1, 2)
            inx, _ = self.gru(inx)
            return self.fc(inx[:, -1, :])

    setattr(m, feature_name, oracle_fn)
    m.Network = Network
    return m


def _twin_cnn(name: str, feature_name: str, oracle_fn, spec: bool):
    """model_spec_cnn.py:20-57 / model_fbanks_cnn.py:68-102 shape twins (same constructor calls in the same order)"""
    m = types.ModuleType(f"twin_{name}")

    class Network(nn.Module):
        def __init__(self):
            super().__init__()
            if spec:
                self.bn1 = nn.BatchNorm2d(1)
                self.conv1 = nn.Conv2d(1, 64, (3, 7), padding=(1, 3))
                self.maxpool1 = nn.MaxPool2d((1, 5))
                self.conv2 = nn.Conv2d(64, 128, (1, 7), padding=(0, 3))
                self.maxpool2 = nn.MaxPool2d((1, 5))
                self.conv3 = nn.Conv2d(128, 256, (1, 12))
                self.conv4 = nn.Conv2d(256, 512, (5, 1), padding=(2, 0))
                self.maxpool3 = nn.MaxPool1d(49)
                self.dropout = nn.Dropout(0.5)
            else:
                self.conv1 = nn.Conv2d(1, 64, kernel_size=(7, 3), padding=(3, 1))
                self.maxpool1 = nn.MaxPool2d((1, 3))
                self.conv2 = nn.Conv2d(64, 128, (1, 7), padding=(0, 3))
                self.maxpool2 = nn.MaxPool2d((1, 4))
                self.conv3 = nn.Conv2d(128, 256, (1, 10))
                self.conv4 = nn.Conv2d(256, 512, (7, 1), padding=(3, 0))
                self.maxpool3 = nn.MaxPool1d(98)
                self.dropout = nn.Dropout()
            self.fc1 = nn.Linear(512, 256)
            self.fc2 = nn.Linear(256, 12)

        def forward(self, x):
            from speechrecognitionproject_b200 import patch
            with torch.no_grad():
                inx = torch.stack([getattr(m, feature_name)(x[i]) for i in range(x.size(0))])
            return patch._cnn_tail(self, inx.to(next(self.parameters()).device))

    setattr(m, feature_name, oracle_fn)
    m.Network = Network
    return m


def twin(name: str):
    import numpy as np
    import oracle
    if name == "model_mfcc_bgru":
        return twin_model_mfcc_bgru()
    if name == "model_spec_bgru":
        return _twin_bgru(name, 321, "compute_spec", lambda s: torch.from_numpy(oracle.spec_ref(s.numpy())))
    if name == "model_spec_cnn":
        return _twin_cnn(name, "compute_spec", lambda s: torch.from_numpy(np.ascontiguousarray(oracle.spec_ref(s.numpy()).T)), True)
    if name == "model_fbanks_cnn":
        return _twin_cnn(name, "filter_banks", lambda s: torch.from_numpy(oracle.fbank_ref(s.numpy())), False)
    raise FileNotFoundError(f"{name}: no twin for this module")


TWINNED = ("model_mfcc_bgru", "model_spec_bgru", "model_spec_cnn", "model_fbanks_cnn")


def load(name: str = "model_mfcc_bgru"):
    """-> (module, "reference" | "twin")"""
    if reference_available():
        return load_reference(name), "reference"
    return twin(name), "twin"
