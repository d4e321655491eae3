"""``patch_model``: re-plumb the feature handoff of an (untouched) reference model.

In the reference every feature-consuming ``Network.forward`` does

    with torch.no_grad():
        inx = torch.ones(B, C, T)                  # CPU
        for i in range(B): inx[i] = compute_xxx(x[i])   # one clip at a time, CPU
    inx = inx.to(DEVICE)                           # H2D of FEATURES
    ... layers ...

(models/model_mfcc_bgru.py:28-37, model_mfrn_bgru.py:127-140, model_spec_bgru.py:26-35,
model_spec_cnn.py:36-57, model_fbanks_cnn.py:83-102).  ``patch_model`` swaps that
``forward`` for one that moves the *PCM* batch to the model's device (or uses it in
place if it already lives there), produces the whole feature batch with one fused
kernel launch in the consumer's layout, and then runs the model's own layers in the
reference's order.  Layers, attribute names and ``state_dict`` keys are untouched, so
existing checkpoints load unchanged.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import features as F

_KINDS = ("mfcc_bgru", "mfrn_bgru", "spec_bgru", "spec_cnn", "fbanks_cnn")


def detect_kind(module) -> str:
    """Which of the five feature-consuming reference modules is this?"""
    net = getattr(module, "Network", None)
    if net is None:
        raise TypeError("expected a reference model module with a `Network` class")
    if hasattr(module, "compute_mfcc"):
        return "mfrn_bgru" if hasattr(module, "ResNet") else "mfcc_bgru"
    if hasattr(module, "compute_spec"):
        return "spec_cnn" if hasattr(module, "F") or "conv1" in net.__init__.__code__.co_names else "spec_bgru"
    if hasattr(module, "filter_banks"):
        return "fbanks_cnn"
    raise TypeError("module has none of compute_mfcc / compute_spec / filter_banks")


_FEATURE_SET = {"mfcc_bgru": "mfcc_tf", "mfrn_bgru": "mfcc_tf", "spec_bgru": "spec_tf", "spec_cnn": "spec_tf",
                "fbanks_cnn": "fbank_tf"}


class SharedFrontEnd:
    """One PCM upload and one launch per feature set for an ENSEMBLE of patched models.

    The reference's ensemble drivers feed the same ``batch['audio']`` to every member in turn
    (analyst_training.py:91-94, predictions.py:59-60): each member recomputes its features on the CPU and
    uploads them; ``spec_cnn`` and ``spec_bgru`` even compute the same spectrogram twice.  Members patched
    with the same ``SharedFrontEnd`` share the device copy of the PCM batch and any feature set two of
    them consume (both spectrogram models take the ``[B, 49, 321]`` tensor).

    The cache holds exactly one batch.  A hit requires the *same tensor object* at the same
    ``_version`` (in-place edits bump it); the object is kept referenced while cached, so its storage
    cannot be recycled under the cache.  Cached features are shared read-only between the members
    (none of the reference layers writes in place)."""

    def __init__(self):
        self._src = None
        self._version = -1
        self._dev = None
        self._pcm = None
        self._feat = {}
        self._members = set()       # feature sets the ensemble's members consume (filled by patch_model)
        self.fuse_max_clips = 256   # spec + fbank from one launch up to this batch size (see features())
        self.uploads = 0            # statistics (tests, logs)
        self.launches = 0

    def register(self, name: str) -> None:
        self._members.add(name)

    def pcm(self, x: torch.Tensor, dev: torch.device) -> torch.Tensor:
        if x is not self._src or x._version != self._version or dev != self._dev:
            self._src, self._version, self._dev = x, x._version, dev
            self._feat.clear()
            if x.device != dev:
                self._pcm = F.to_device(x, dev) if dev.type == "cuda" else x.to(dev)
                self.uploads += 1
            else:
                self._pcm = x
        return self._pcm

    def features(self, name: str, fn: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, dev: torch.device):
        xd = self.pcm(x, dev)
        if name not in self._feat:
            if (name in ("spec_tf", "fbank_tf") and {"spec_tf", "fbank_tf"} <= self._members and xd.is_cuda
                    and fn is _DEFAULT_FN.get(name) and xd.size(0) <= self.fuse_max_clips):
                # a spectrogram member and the fbank member share the batch: both feature sets from ONE launch, the PCM
                # read from HBM once (features.spec_fbank; bit-identical to the separate calls).  Measured on B200: 33 us
                # against 49 us for two launches up to 64 clips (the reference's ensemble drivers run batch_size = 1,
                # analyst_training.py:84); from ~512 clips on two launches are as fast or faster (1.52 vs 1.80 ms at 16,384:
                # both kernels are bound on chip, not by the PCM read), so large batches keep them.
                self._feat["spec_tf"], self._feat["fbank_tf"] = F.spec_fbank(xd, F.R_SPEC, F.R_FBANK, layout="tf")
            else:
                self._feat[name] = fn(xd)
            self.launches += 1
        return self._feat[name]

    def clear(self) -> None:
        members, fuse = self._members, self.fuse_max_clips
        self.__init__()
        self._members, self.fuse_max_clips = members, fuse


def _mfcc_tf(x):
    return F.mfcc(x, F.R_MFCC, layout="tf")            # [B,51,39] == transpose(inx,1,2)


def _spec_tf(x):
    return F.spec(x, F.R_SPEC, layout="tf")            # [B,49,321] == transpose(inx,1,2) (bgru) == compute_spec(...).T (cnn)


def _fbank_tf(x):
    return F.fbank(x, F.R_FBANK)                       # [B,98,120]


_DEFAULT_FN = {"mfcc_tf": _mfcc_tf, "spec_tf": _spec_tf, "fbank_tf": _fbank_tf}


def default_feature_fn(kind: str) -> Callable[[torch.Tensor], torch.Tensor]:
    """Batched device features in the layout the model's first layer consumes."""
    if kind not in _FEATURE_SET:
        raise ValueError(kind)
    return _DEFAULT_FN[_FEATURE_SET[kind]]


def _cnn_tail(self, x):
    # model_spec_cnn.py:43-57 / model_fbanks_cnn.py:88-102 (identical layer order)
    x = x.unsqueeze(1)
    x = self.conv1(x)
    x = self.maxpool1(x)
    x = self.conv2(x)
    x = self.maxpool2(x)
    x = self.conv3(x)
    x = self.conv4(x)
    x = x.squeeze(3)
    x = self.maxpool3(x)
    x = x.squeeze(2)
    x = self.dropout(x)
    x = self.fc1(x)
    return self.fc2(x)


def _bgru_tail(self, x):
    # model_mfcc_bgru.py:35-37 / model_spec_bgru.py:33-35 (x already [B, T, C])
    x, _ = self.gru(x)
    return self.fc(x[:, -1, :])


def make_forward(kind: str, feature_fn: Optional[Callable] = None, frontend: Optional[SharedFrontEnd] = None):
    if kind not in _KINDS:
        raise ValueError(f"unknown model kind {kind!r}; expected one of {_KINDS}")
    feat = feature_fn or default_feature_fn(kind)
    if frontend is not None:
        frontend.register(_FEATURE_SET[kind])

    def forward(self, x):
        dev = next(self.parameters()).device
        with torch.no_grad():                          # features are never differentiated (:29)
            if frontend is not None:                   # ensemble: upload / launch once per batch and feature set
                f = frontend.features(_FEATURE_SET[kind], feat, x, dev)
                x = frontend.pcm(x, dev)
            else:
                if x.device != dev:                    # the only H2D: raw PCM (pageable batches through libsrfe's staging ring)
                    x = F.to_device(x, dev) if dev.type == "cuda" else x.to(dev)
                f = feat(x)
        if f.device != dev:
            f = f.to(dev)
        if kind in ("mfcc_bgru", "spec_bgru"):
            return _bgru_tail(self, f)
        if kind in ("spec_cnn", "fbanks_cnn"):
            return _cnn_tail(self, f)
        # mfrn_bgru: model_mfrn_bgru.py:135-140
        r = self.resnet(x)
        return self.gru(torch.cat((r, f), 2))

    forward.__srfe_patched__ = kind
    return forward


def patch_model(module, kind: Optional[str] = None, feature_fn: Optional[Callable] = None,
                frontend: Optional[SharedFrontEnd] = None) -> str:
    """Patch ``module.Network.forward`` in place; returns the detected model kind.

    ``feature_fn(x[B,N]) -> features`` overrides the feature producer (tests use it to
    check the re-plumbed ``forward`` against the reference's own on the CPU).
    ``frontend``: a ``SharedFrontEnd`` common to the members of an ensemble."""
    kind = kind or detect_kind(module)
    net = module.Network
    if not hasattr(net, "__srfe_original_forward__"):
        net.__srfe_original_forward__ = net.forward
    net.forward = make_forward(kind, feature_fn, frontend)
    return kind


def unpatch_model(module) -> None:
    net = module.Network
    if hasattr(net, "__srfe_original_forward__"):
        net.forward = net.__srfe_original_forward__
        del net.__srfe_original_forward__
