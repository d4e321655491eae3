"""CPU oracle for the acoustic feature front end.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``speechrecognitionproject_b200``) may import,
call, link or execute anything in this directory.  The only legal users are
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker / the CPU arm.

What it restates (reference = /root/reference, remit0/SpeechRecognitionProject):

* ``models/model_spec_bgru.py:11-17`` / ``models/model_spec_cnn.py:12-18``
  (``compute_spec``; arithmetic inside scipy.signal.spectrogram)
* ``models/model_fbanks_cnn.py:15-66`` (``filter_banks``; pure numpy)
* ``models/model_mfcc_bgru.py:11-19`` / ``models/model_mfrn_bgru.py:11-19``
  (``compute_mfcc``; arithmetic inside the un-vendored, un-pinned third-party
  package librosa (inferred 0.6.x) -> restated from its published algorithm)

Pinning status: the reference ships no tests, fixtures or golden vectors, so
"parity" is pinned on outputs of the reference functions themselves, run in the
build container by ``oracle/make_golden.py`` (spec and fbank import and run
unmodified from /root/reference; MFCC cannot, because librosa is absent, and is
cross-checked LIVE -- tests/test_mfcc_crosscheck.py, both parameter sets -- against
two independent implementations instead: ``transformers.audio_utils`` and
``torchaudio``, oracle/crosscheck.py).  MFCC parity is therefore "pinned to the
restatement + cross-checks", NOT to a librosa run: **MFCC parity unpinned by the
reference itself.**  ``oracle/augment.py`` (augmentation, dataset.py:107-202) has no
reference test either and replaces the reference's global RNG by a counter-based
stream by design: **parity unpinned**, arithmetic restated line by line.
"""
from .features import (  # noqa: F401
    SpecParams, FbankParams, MfccParams,
    R_SPEC, C_SPEC, R_FBANK, C_FBANK, R_MFCC, C_MFCC, C_MFCC_D2, PRESETS,
    spec_ref, spec_truth, fbank_ref, fbank_truth, mfcc_ref, mfcc_truth,
    spec_num_frames, fbank_num_frames, mfcc_num_frames,
    htk_floor_filterbank, slaney_mel_filterbank, dct2_ortho_matrix,
    tukey_periodic, hann_periodic, hamming_symmetric,
)
from .corpus import synthetic_corpus, edge_suite  # noqa: F401
