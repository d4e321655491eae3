"""ctypes binding of ``libsrfe.so`` (the C ABI of include/srfe.h).

The library is built in-tree by ``speechrecognitionproject_b200.build`` /
``__graft_entry__.build()``.  There is no fallback: if the shared object is
missing the import of any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrfe.so")

SRFE_OK = 0
LAYOUT_FT, LAYOUT_TF = 0, 1
STATUS_NAMES = {0: "SRFE_OK", -1: "SRFE_ERR_BAD_ARG", -2: "SRFE_ERR_UNSUPPORTED", -3: "SRFE_ERR_CUDA",
                -4: "SRFE_ERR_NO_DEVICE", -5: "SRFE_ERR_TOO_LARGE"}


class SpecParamsC(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("nperseg", C.c_int32), ("noverlap", C.c_int32),
                ("take_log", C.c_int32), ("log_eps", C.c_float), ("layout", C.c_int32)]


class FbankParamsC(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("frame_len", C.c_int32), ("frame_step", C.c_int32),
                ("n_fft", C.c_int32), ("preemph", C.c_float), ("nfilt", C.c_int32), ("vtlp_alpha", C.c_float)]


class MfccParamsC(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop", C.c_int32),
                ("n_mels", C.c_int32), ("fmin", C.c_float), ("fmax", C.c_float), ("n_mfcc", C.c_int32),
                ("n_deltas", C.c_int32), ("top_db", C.c_float), ("amin", C.c_float), ("layout", C.c_int32)]


class AugmentParamsC(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("shift_lo", C.c_float), ("shift_hi", C.c_float), ("noise_lo", C.c_float),
                ("noise_hi", C.c_float), ("snr_lo", C.c_float), ("snr_hi", C.c_float), ("pitch_lo", C.c_float),
                ("pitch_hi", C.c_float), ("speed_lo", C.c_float), ("speed_hi", C.c_float), ("shift_range", C.c_int32),
                ("noise_upper", C.c_float)]


# every symbol include/srfe.h declares: name -> (restype, argtypes)
_i64, _i32, _vp, _dp = C.c_int64, C.c_int, C.c_void_p, C.POINTER(C.c_double)
_SHAPE = C.POINTER(C.c_int64 * 2)
SYMBOLS = {
    "srfe_version": (_i32, []),
    "srfe_last_error_string": (C.c_char_p, []),
    "srfe_device_count": (_i32, []),
    "srfe_launch_count": (_i64, []),
    "srfe_release_host_workspace": (_i32, []),
    "srfe_upload": (_i32, [_vp, _i64, _i64, _i64, _vp, _i32, _vp]),
    "srfe_spec_fbank_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(SpecParamsC), C.POINTER(FbankParamsC), _vp, _vp, _vp]),
    "srfe_spec_fbank_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(SpecParamsC), C.POINTER(FbankParamsC), _vp, _vp, _vp]),
    "srfe_augment_i16": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _i32, C.POINTER(AugmentParamsC), _vp, _vp, _vp]),
    "srfe_set_tuning": (_i32, [C.c_char_p, _i32]),
    "srfe_spec_out_shape": (_i64, [C.POINTER(SpecParamsC), _i64, _SHAPE]),
    "srfe_fbank_out_shape": (_i64, [C.POINTER(FbankParamsC), _i64, _SHAPE]),
    "srfe_mfcc_out_shape": (_i64, [C.POINTER(MfccParamsC), _i64, _SHAPE]),
    "srfe_spec_bytes_per_clip": (_i64, [C.POINTER(SpecParamsC), _i64]),
    "srfe_fbank_bytes_per_clip": (_i64, [C.POINTER(FbankParamsC), _i64]),
    "srfe_mfcc_bytes_per_clip": (_i64, [C.POINTER(MfccParamsC), _i64]),
    "srfe_spec_workspace_bytes": (_i64, [C.POINTER(SpecParamsC), _i64, _i64]),
    "srfe_fbank_workspace_bytes": (_i64, [C.POINTER(FbankParamsC), _i64, _i64]),
    "srfe_mfcc_workspace_bytes": (_i64, [C.POINTER(MfccParamsC), _i64, _i64]),
    "srfe_spec_window_f64": (_i32, [C.POINTER(SpecParamsC), _dp]),
    "srfe_fbank_window_f64": (_i32, [C.POINTER(FbankParamsC), _dp]),
    "srfe_mfcc_window_f64": (_i32, [C.POINTER(MfccParamsC), _dp]),
    "srfe_fbank_filters_f64": (_i32, [C.POINTER(FbankParamsC), _dp]),
    "srfe_mfcc_filters_f64": (_i32, [C.POINTER(MfccParamsC), _dp]),
    "srfe_mfcc_dct_f64": (_i32, [C.POINTER(MfccParamsC), _dp]),
    "srfe_spec_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(SpecParamsC), _vp, _vp]),
    "srfe_fbank_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(FbankParamsC), _vp, _vp]),
    "srfe_mfcc_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(MfccParamsC), _vp, _vp]),
    "srfe_spec_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(SpecParamsC), _vp, _vp]),
    "srfe_fbank_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(FbankParamsC), _vp, _vp]),
    "srfe_mfcc_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(MfccParamsC), _vp, _vp]),
    "srfe_spec_host_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(SpecParamsC), _vp, _i32]),
    "srfe_fbank_host_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(FbankParamsC), _vp, _i32]),
    "srfe_mfcc_host_i16": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(MfccParamsC), _vp, _i32]),
    "srfe_spec_host_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(SpecParamsC), _vp, _i32]),
    "srfe_fbank_host_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(FbankParamsC), _vp, _i32]),
    "srfe_mfcc_host_f32": (_i32, [_vp, _i64, _i64, _i64, C.POINTER(MfccParamsC), _vp, _i32]),
}

_lib = None


class SrfeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load libsrfe.so once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m speechrecognitionproject_b200.build` "
                "(there is no CPU fallback for the feature front end)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise SrfeError(int(rc), lib().srfe_last_error_string().decode("utf-8", "replace"))
    return int(rc)
