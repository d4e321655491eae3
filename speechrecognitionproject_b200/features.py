"""Host-side mirror of the reference's feature interface, backed by libsrfe.so.

Reference interface (citations into /root/reference):

* ``compute_spec(sample)``  models/model_spec_bgru.py:11-17 -> FloatTensor[321,49]
                            models/model_spec_cnn.py:12-18  -> FloatTensor[49,321]
* ``filter_banks(sample)``  models/model_fbanks_cnn.py:15-66 -> FloatTensor[98,120]
* ``compute_mfcc(sample)``  models/model_mfcc_bgru.py:11-19  -> FloatTensor[39,51]
  (``sample`` = 1-D float32 tensor of 16000 int16-scale samples, dataset.py:117)

plus the batched forms that replace the per-clip loops of ``Network.forward``
(model_mfcc_bgru.py:29-34 etc.): ``spec(x[B,N])``, ``fbank(x)``, ``mfcc(x)``.

Compute always happens in the CUDA kernels: CUDA tensors go through the
``srfe::*`` custom ops (stream-ordered, no sync); CPU tensors go through the
``srfe_*_host_f32`` entry points (H2D -> kernel -> D2H) and come back on the CPU,
like the reference's return values.  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass, replace
from typing import Optional

import torch

from . import _lib
from ._lib import LAYOUT_FT, LAYOUT_TF, FbankParamsC, MfccParamsC, SpecParamsC

__all__ = [
    "SpecParams", "FbankParams", "MfccParams", "PRESETS",
    "R_SPEC", "C_SPEC", "R_FBANK", "C_FBANK", "R_MFCC", "C_MFCC", "C_MFCC_D2",
    "spec", "fbank", "mfcc", "spec_fbank", "compute_spec", "filter_banks", "compute_mfcc",
    "out_shape", "bytes_per_clip", "launch_count", "set_tuning", "release_host_workspace", "to_device",
]


def _layout_code(layout: str) -> int:
    if layout not in ("ft", "tf"):
        raise ValueError("layout must be 'ft' ([B, feature, time]) or 'tf' ([B, time, feature])")
    return LAYOUT_FT if layout == "ft" else LAYOUT_TF


@dataclass(frozen=True)
class SpecParams:
    """scipy.signal.spectrogram call of model_spec_bgru.py:13-14, literals lifted."""
    fs: int = 16000
    nperseg: int = 640
    noverlap: int = 320
    log: bool = True
    eps: float = 1e-10
    layout: str = "ft"

    def op_args(self) -> tuple:
        """positional arguments of ``torch.ops.srfe.spec`` after the PCM tensor (plain Python values: traceable)"""
        return (self.fs, self.nperseg, self.noverlap, bool(self.log), float(self.eps), _layout_code(self.layout))

    def to_c(self) -> SpecParamsC:
        return SpecParamsC(self.fs, self.nperseg, self.noverlap, int(self.log), self.eps, _layout_code(self.layout))


@dataclass(frozen=True)
class FbankParams:
    """model_fbanks_cnn.py:18-45 literals, lifted."""
    fs: int = 16000
    frame_len: int = 400
    frame_step: int = 160
    nfft: int = 512
    preemph: float = 0.97
    nfilt: int = 120
    vtlp_alpha: float = 0.0     # 0 = off; legacy/model_8/dataset_top.py:251-252 warps the filter centres by alpha ~ U(0.9, 1.1)

    def op_args(self) -> tuple:
        return (self.fs, self.frame_len, self.frame_step, self.nfft, float(self.preemph), self.nfilt, float(self.vtlp_alpha))

    def to_c(self) -> FbankParamsC:
        return FbankParamsC(self.fs, self.frame_len, self.frame_step, self.nfft, self.preemph, self.nfilt, self.vtlp_alpha)


@dataclass(frozen=True)
class MfccParams:
    """librosa.feature.mfcc call of model_mfcc_bgru.py:13 (+ librosa-0.6 defaults) and
    the two np.gradient passes of :14-16."""
    sr: int = 16000
    n_fft: int = 640
    win_length: Optional[int] = None
    hop: int = 320
    n_mels: int = 128
    fmin: float = 0.0
    fmax: Optional[float] = None
    n_mfcc: int = 13
    n_deltas: int = 2
    top_db: Optional[float] = 80.0
    amin: float = 1e-10
    layout: str = "ft"

    def op_args(self) -> tuple:
        return (self.sr, self.n_fft, self.win_length or 0, self.hop, self.n_mels, float(self.fmin),
                float(self.fmax) if self.fmax else 0.0, self.n_mfcc, self.n_deltas,
                -1.0 if self.top_db is None else float(self.top_db), float(self.amin), _layout_code(self.layout))

    def to_c(self) -> MfccParamsC:
        return MfccParamsC(self.sr, self.n_fft, self.win_length or 0, self.hop, self.n_mels, self.fmin,
                           self.fmax if self.fmax else 0.0, self.n_mfcc, self.n_deltas,
                           -1.0 if self.top_db is None else self.top_db, self.amin, _layout_code(self.layout))


R_SPEC = SpecParams()
C_SPEC = SpecParams(nperseg=512, noverlap=256)
R_FBANK = FbankParams()
C_FBANK = FbankParams(nfilt=40)
R_MFCC = MfccParams()
C_MFCC = MfccParams(n_fft=512, win_length=400, hop=160, n_mfcc=40, n_deltas=0)
C_MFCC_D2 = replace(C_MFCC, n_deltas=2)
PRESETS = {"R-SPEC": R_SPEC, "C-SPEC": C_SPEC, "R-FBANK": R_FBANK, "C-FBANK": C_FBANK,
           "R-MFCC": R_MFCC, "C-MFCC": C_MFCC, "C-MFCC-D2": C_MFCC_D2}

_FAMILY = {SpecParams: "spec", FbankParams: "fbank", MfccParams: "mfcc"}


def out_shape(params, n_samples: int) -> tuple[int, int]:
    """Per-clip feature shape for ``n_samples`` input samples (host only, no GPU needed)."""
    fam = _FAMILY[type(params)]
    shp = (C.c_int64 * 2)()
    cp = params.to_c()
    _lib.check(getattr(_lib.lib(), f"srfe_{fam}_out_shape")(C.byref(cp), n_samples, C.byref(shp)))
    return int(shp[0]), int(shp[1])


def bytes_per_clip(params, n_samples: int = 16000) -> int:
    """Algorithmic bytes per clip: fp32 PCM in + fp32 features out."""
    fam = _FAMILY[type(params)]
    cp = params.to_c()
    return _lib.check(getattr(_lib.lib(), f"srfe_{fam}_bytes_per_clip")(C.byref(cp), n_samples))


def launch_count() -> int:
    return int(_lib.lib().srfe_launch_count())


_TUNING_KNOBS = ("warps", "ctas", "cpc", "dct_cb", "dct_pq", "mfcc_tc", "stage", "fbank_tc")


def set_tuning(**knobs: int) -> None:
    """Launch-shape overrides for tests and tuning sweeps (``srfe_set_tuning``; 0 = automatic).  ``set_tuning()``
    with no arguments resets every knob.  Results never depend on them."""
    if not knobs:
        knobs = {k: 0 for k in _TUNING_KNOBS}
    for k, v in knobs.items():
        _lib.check(_lib.lib().srfe_set_tuning(k.encode(), int(v)))


def to_device(x: torch.Tensor, device=None) -> torch.Tensor:
    """``x.to(device)`` for a CPU PCM batch ``[n_clips, n_samples]`` (float32 or int16), through ``srfe_upload``: pageable
    tensors -- what the reference's DataLoader yields (training.py:77, no ``pin_memory``) -- go through libsrfe's pinned
    staging ring with a few copy threads (2x torch's pageable copy on the B200 box), pinned ones straight over PCIe.
    Stream-ordered on the current stream of ``device``; CUDA tensors are returned as they are."""
    if x.is_cuda:
        return x
    if not torch.cuda.is_available():
        raise RuntimeError("srfe: no CUDA device visible")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype not in (torch.float32, torch.int16):
        return x.to(dev, non_blocking=True)
    out = torch.empty(x.shape, dtype=x.dtype, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        es = x.element_size()
        _lib.check(_lib.lib().srfe_upload(x.data_ptr(), x.size(0), x.size(1) * es, (x.stride(0) if x.size(0) > 1 else x.size(1)) * es,
                                          out.data_ptr(), dev.index, stream))
    return out


def release_host_workspace() -> None:
    """Free the per-device staging workspaces of the host entry points (``srfe_release_host_workspace``)."""
    _lib.check(_lib.lib().srfe_release_host_workspace())


# ------------------------------------------------------------------------------------
# the three custom ops (CUDA only; no autograd formula: features are always computed
# under torch.no_grad(), model_mfcc_bgru.py:29)
# ------------------------------------------------------------------------------------
def _check_pcm(x: torch.Tensor) -> None:
    if x.dtype not in (torch.float32, torch.int16):
        raise TypeError(f"PCM must be float32 (dataset.py:117) or int16 (the wav's own type, dataset.py:103), got {x.dtype}")
    if x.dim() != 2:
        raise ValueError("expected a [n_clips, n_samples] tensor")


def _prep(x: torch.Tensor) -> torch.Tensor:
    """Make a [n_clips, n_samples] view the kernels can read: unit sample stride, rows that do not overlap and start on
    a sample pair.  Only real tensors come here (the custom ops call it on their concrete inputs), never tracing fakes."""
    _check_pcm(x)
    pair = 2 * x.element_size()                    # rows aligned to a sample pair
    if (x.stride(1) != 1 or (x.size(0) > 1 and (x.stride(0) % 2 or x.stride(0) < x.size(1)))   # odd / overlapping / expanded rows
            or x.data_ptr() % pair):
        x = x.contiguous()
        if x.data_ptr() % pair:                    # odd-offset view of a larger buffer
            x = x.clone()
    return x


def _suffix(x: torch.Tensor) -> str:
    return "i16" if x.dtype == torch.int16 else "f32"


def _run_device(fam: str, cp, x: torch.Tensor, out: torch.Tensor) -> None:
    fn = getattr(_lib.lib(), f"srfe_{fam}_{_suffix(x)}")
    stride = x.stride(0) if x.size(0) > 1 else x.size(1)
    if x.device.index == torch.cuda.current_device():          # the common case: skip the device-guard round trip
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(fn(x.data_ptr(), x.size(0), x.size(1), stride, C.byref(cp), out.data_ptr(), stream))
        return
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(fn(x.data_ptr(), x.size(0), x.size(1), stride, C.byref(cp), out.data_ptr(), stream))


# Eager fast path.  The reference's ensemble drivers call the models with batch_size = 1
# (analyst_training.py:84, predictions.py:58): per-call host overhead is what such callers see.  Outside of
# torch.compile tracing the public functions therefore call the kernels directly (same code as the registered
# ops below, minus the dispatcher); parameter structs and output shapes are cached per (frozen) parameter set.
@functools.lru_cache(maxsize=256)
def _c_params(params):
    return params.to_c()


@functools.lru_cache(maxsize=4096)
def _cached_shape(params, n_samples: int) -> tuple[int, int]:
    return out_shape(params, n_samples)


def _eager_device(fam: str, params, xb: torch.Tensor) -> torch.Tensor:
    out = torch.empty((xb.size(0),) + _cached_shape(params, xb.size(1)), dtype=torch.float32, device=xb.device)
    _run_device(fam, _c_params(params), xb, out)
    return out


@torch.library.custom_op("srfe::spec", mutates_args=(), device_types="cuda")
def _spec_op(pcm: torch.Tensor, fs: int, nperseg: int, noverlap: int, take_log: bool, eps: float,
             layout: int) -> torch.Tensor:
    cp = SpecParamsC(fs, nperseg, noverlap, int(take_log), eps, layout)
    shp = (C.c_int64 * 2)()
    _lib.check(_lib.lib().srfe_spec_out_shape(C.byref(cp), pcm.size(1), C.byref(shp)))
    out = torch.empty((pcm.size(0), shp[0], shp[1]), dtype=torch.float32, device=pcm.device)
    _run_device("spec", cp, _prep(pcm), out)
    return out


@_spec_op.register_fake
def _(pcm, fs, nperseg, noverlap, take_log, eps, layout):
    t = (pcm.size(1) - noverlap) // (nperseg - noverlap)
    f = nperseg // 2 + 1
    return pcm.new_empty((pcm.size(0), f, t) if layout == LAYOUT_FT else (pcm.size(0), t, f), dtype=torch.float32)


@torch.library.custom_op("srfe::fbank", mutates_args=(), device_types="cuda")
def _fbank_op(pcm: torch.Tensor, fs: int, frame_len: int, frame_step: int, nfft: int, preemph: float,
              nfilt: int, vtlp_alpha: float) -> torch.Tensor:
    cp = FbankParamsC(fs, frame_len, frame_step, nfft, preemph, nfilt, vtlp_alpha)
    shp = (C.c_int64 * 2)()
    _lib.check(_lib.lib().srfe_fbank_out_shape(C.byref(cp), pcm.size(1), C.byref(shp)))
    out = torch.empty((pcm.size(0), shp[0], shp[1]), dtype=torch.float32, device=pcm.device)
    _run_device("fbank", cp, _prep(pcm), out)
    return out


@_fbank_op.register_fake
def _(pcm, fs, frame_len, frame_step, nfft, preemph, nfilt, vtlp_alpha):
    t = -(-abs(pcm.size(1) - frame_len) // frame_step)
    return pcm.new_empty((pcm.size(0), t, nfilt), dtype=torch.float32)


@torch.library.custom_op("srfe::mfcc", mutates_args=(), device_types="cuda")
def _mfcc_op(pcm: torch.Tensor, sr: int, n_fft: int, win_length: int, hop: int, n_mels: int, fmin: float,
             fmax: float, n_mfcc: int, n_deltas: int, top_db: float, amin: float, layout: int) -> torch.Tensor:
    cp = MfccParamsC(sr, n_fft, win_length, hop, n_mels, fmin, fmax, n_mfcc, n_deltas, top_db, amin, layout)
    shp = (C.c_int64 * 2)()
    _lib.check(_lib.lib().srfe_mfcc_out_shape(C.byref(cp), pcm.size(1), C.byref(shp)))
    out = torch.empty((pcm.size(0), shp[0], shp[1]), dtype=torch.float32, device=pcm.device)
    _run_device("mfcc", cp, _prep(pcm), out)
    return out


@_mfcc_op.register_fake
def _(pcm, sr, n_fft, win_length, hop, n_mels, fmin, fmax, n_mfcc, n_deltas, top_db, amin, layout):
    t = 1 + pcm.size(1) // hop
    r = (1 + n_deltas) * n_mfcc
    return pcm.new_empty((pcm.size(0), r, t) if layout == LAYOUT_FT else (pcm.size(0), t, r), dtype=torch.float32)


# ------------------------------------------------------------------------------------
# batched public API
# ------------------------------------------------------------------------------------
def _run_host(fam: str, cp, x: torch.Tensor, shape: tuple[int, int], device: Optional[int]) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("srfe: no CUDA device visible and there is no CPU fallback for the feature front end")
    dev = torch.cuda.current_device() if device is None else device
    out = torch.empty((x.size(0),) + shape, dtype=torch.float32, pin_memory=x.is_pinned())
    stride = x.stride(0) if x.size(0) > 1 else x.size(1)
    fn = getattr(_lib.lib(), f"srfe_{fam}_host_{_suffix(x)}")
    _lib.check(fn(x.data_ptr(), x.size(0), x.size(1), stride, C.byref(cp), out.data_ptr(), dev))
    return out


def _dispatch(fam: str, params, x: torch.Tensor, device: Optional[int]) -> torch.Tensor:
    single = x.dim() == 1
    xb = x.unsqueeze(0) if single else x
    if xb.is_cuda and not torch.compiler.is_compiling():
        y = _eager_device(fam, params, _prep(xb))
    elif xb.is_cuda:
        # under torch.compile the tensors are fakes: no data_ptr / stride fix-ups here (the op does them on the real
        # tensors at run time), only dtype / rank checks that fakes can answer
        _check_pcm(xb)
        y = getattr(torch.ops.srfe, fam)(xb, *params.op_args())
    else:
        y = _run_host(fam, params.to_c(), _prep(xb), out_shape(params, xb.size(1)), device)
    return y[0] if single else y


def spec(x: torch.Tensor, params: SpecParams = R_SPEC, *, layout: Optional[str] = None,
         device: Optional[int] = None) -> torch.Tensor:
    """Batched ``compute_spec``: ``x[B, N]`` -> ``[B, nperseg/2+1, T]`` ('ft') or ``[B, T, nperseg/2+1]`` ('tf')."""
    if layout is not None and layout != params.layout:
        params = replace(params, layout=layout)
    return _dispatch("spec", params, x, device)


def fbank(x: torch.Tensor, params: FbankParams = R_FBANK, *, device: Optional[int] = None) -> torch.Tensor:
    """Batched ``filter_banks``: ``x[B, N]`` -> ``[B, T, nfilt]``."""
    return _dispatch("fbank", params, x, device)


def mfcc(x: torch.Tensor, params: MfccParams = R_MFCC, *, layout: Optional[str] = None,
         device: Optional[int] = None) -> torch.Tensor:
    """Batched ``compute_mfcc``: ``x[B, N]`` -> ``[B, (1+n_deltas) n_mfcc, T]`` ('ft') or transposed ('tf')."""
    if layout is not None and layout != params.layout:
        params = replace(params, layout=layout)
    return _dispatch("mfcc", params, x, device)


def spec_fbank(x: torch.Tensor, spec_params: SpecParams = R_SPEC, fbank_params: FbankParams = R_FBANK, *,
               layout: Optional[str] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Spectrogram AND log-fbank features of the same CUDA batch in ONE launch (``srfe_spec_fbank_*``): what
    the reference's ensemble computes twice from the same ``batch['audio']`` (analyst_training.py:91-94).  Bit-identical
    to ``spec(x, ...)`` and ``fbank(x, ...)``; the PCM is read from HBM once for both."""
    if layout is not None and layout != spec_params.layout:
        spec_params = replace(spec_params, layout=layout)
    if not x.is_cuda:
        raise TypeError("spec_fbank: expected a CUDA tensor (host batches: upload once, then call this)")
    single = x.dim() == 1
    xb = _prep(x.unsqueeze(0) if single else x)
    n = xb.size(0)
    ys = torch.empty((n,) + _cached_shape(spec_params, xb.size(1)), dtype=torch.float32, device=xb.device)
    yf = torch.empty((n,) + _cached_shape(fbank_params, xb.size(1)), dtype=torch.float32, device=xb.device)
    fn = getattr(_lib.lib(), f"srfe_spec_fbank_{_suffix(xb)}")
    stride = xb.stride(0) if n > 1 else xb.size(1)
    with torch.cuda.device(xb.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(fn(xb.data_ptr(), n, xb.size(1), stride, C.byref(_c_params(spec_params)), C.byref(_c_params(fbank_params)),
                      ys.data_ptr(), yf.data_ptr(), stream))
    return (ys[0], yf[0]) if single else (ys, yf)


# ------------------------------------------------------------------------------------
# reference-named single-clip functions (same signatures / shapes as the reference)
# ------------------------------------------------------------------------------------
def compute_spec(sample: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """model_spec_bgru.compute_spec (``transpose=False`` -> [321,49]) or
    model_spec_cnn.compute_spec (``transpose=True`` -> [49,321])."""
    return spec(sample, R_SPEC, layout="tf" if transpose else "ft")


def filter_banks(sample: torch.Tensor) -> torch.Tensor:
    """model_fbanks_cnn.filter_banks -> [98,120] (time x features)."""
    return fbank(sample, R_FBANK)


def compute_mfcc(sample: torch.Tensor) -> torch.Tensor:
    """model_mfcc_bgru.compute_mfcc -> [39,51] (13 static, 13 delta, 13 delta-delta rows)."""
    return mfcc(sample, R_MFCC)
