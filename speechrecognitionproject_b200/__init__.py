"""B200-native acoustic feature front end (spectrogram / log-fbank / MFCC) --
a drop-in for the feature hot path of remit0/SpeechRecognitionProject.

Public surface: :mod:`speechrecognitionproject_b200.features` (reference-named
functions and batched ops over the C ABI of ``include/srfe.h``) and
:func:`speechrecognitionproject_b200.patch.patch_model` (with
:class:`~speechrecognitionproject_b200.patch.SharedFrontEnd` for ensembles).
"""
from .features import (  # noqa: F401
    SpecParams, FbankParams, MfccParams, PRESETS,
    R_SPEC, C_SPEC, R_FBANK, C_FBANK, R_MFCC, C_MFCC, C_MFCC_D2,
    spec, fbank, mfcc, spec_fbank, compute_spec, filter_banks, compute_mfcc,
    out_shape, bytes_per_clip, launch_count, set_tuning, release_host_workspace, to_device,
)
from .patch import patch_model, SharedFrontEnd  # noqa: F401

__version__ = "0.1"
