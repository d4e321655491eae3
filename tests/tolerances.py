"""Parity tolerances (BASELINE.json north_star + SURVEY.md section 8c), in one place.

* MFCC (static, delta, delta-delta): max abs error <= 1e-3 on every element.
* log-mel / log-fbank: max abs error <= 1e-3 on every element whose *reference* value
  lies within ``LOGMEL_DOMAIN`` units (the path's own 20 log10(power) scale) of the clip
  maximum.  Elements further down are spectral nulls of the pre-emphasised signal that an
  fp32 FFT cannot resolve (SURVEY 8c measured the same for any all-fp32 pipeline); they are
  counted and bounded loosely, never silently dropped.
* power spectra: relative error <= 1e-4 with denominator max(|ref|, 1e-5 * clip max)
  against the float64 truth; the log output is checked at max abs <= 1e-3 on the same
  domain as log-mel.
"""
MFCC_ABS = 1e-3
LOGMEL_ABS = 1e-3
LOGMEL_DOMAIN = 100.0          # units below the clip maximum (20 log10 scale)
LOGSPEC_ABS = 1e-3
LOGSPEC_DOMAIN_NEPER = 11.5    # ln scale: 50 dB below the clip maximum = 11.5 neper
PSD_REL = 1e-4
PSD_FLOOR = 1e-5               # denominator floor, relative to the clip maximum
