"""Smallest run that touches every kernel family / code path (written for compute-sanitizer; the tool is closed on the
GPU pool, so it serves as an all-paths run: every output finite, no launch error)."""
import sys, torch
sys.path.insert(0, ".")
from dataclasses import replace
import speechrecognitionproject_b200 as S
torch.manual_seed(0)
x = (torch.randn(5, 16000, device="cuda") * 2000).round()
xi = x.to(torch.int16)
outs = []
for p in (S.R_SPEC, replace(S.R_SPEC, layout="tf"), S.C_SPEC, replace(S.C_SPEC, layout="tf")):
    outs.append(S.spec(x, p)); outs.append(S.spec(xi, p))
for p in (S.R_FBANK, S.C_FBANK, S.FbankParams(nfft=640, frame_len=640, frame_step=320, nfilt=80)):
    outs.append(S.fbank(x, p)); outs.append(S.fbank(xi, p))
for p in (S.R_MFCC, S.C_MFCC, S.C_MFCC_D2, replace(S.R_MFCC, layout="tf"), S.MfccParams(n_fft=640, hop=320, n_mels=100, n_mfcc=13),
          S.MfccParams(n_fft=512, win_length=400, hop=160, n_mels=40, n_mfcc=13, n_deltas=2)):
    outs.append(S.mfcc(x, p)); outs.append(S.mfcc(xi[:3], p))
outs.append(S.mfcc(x[:, :700].contiguous(), S.R_MFCC))
# round-2 paths: tcgen05 log-filterbank kernel, spectrogram + filterbank from one launch, TMA-staged frames, VTLP bank,
# the classic MFCC kernel forced on the tcgen05-sized shapes, augmentation
S.set_tuning(fbank_tc=2)
for p in (S.R_FBANK, S.C_FBANK):
    outs.append(S.fbank(x, p)); outs.append(S.fbank(xi, p))
S.set_tuning()
outs.extend(S.spec_fbank(x, replace(S.R_SPEC, layout="tf"), S.R_FBANK)); outs.extend(S.spec_fbank(xi, S.C_SPEC, S.C_FBANK))
S.set_tuning(stage=2, warps=16, ctas=1)
outs.append(S.spec(x, replace(S.C_SPEC, layout="tf"))); outs.append(S.spec(xi, replace(S.R_SPEC, layout="tf")))
S.set_tuning()
S.set_tuning(mfcc_tc=1)
outs.append(S.mfcc(x, S.C_MFCC)); outs.append(S.mfcc(x, S.C_MFCC_D2))
S.set_tuning()
outs.append(S.fbank(x, replace(S.R_FBANK, vtlp_alpha=1.07)))
from speechrecognitionproject_b200 import augment as GA
bank = GA.NoiseBank([(torch.randn(20000 + 3000 * i) * 500).round().to(torch.int16) for i in range(3)])
kind = torch.tensor([0, 1, 2, 0, 0], dtype=torch.int8)
outs.append(GA.augment(xi, bank, GA.AugmentParams(seed=3), kind=kind, first_index=2**33 + 5)[0])
torch.cuda.synchronize()
print("ok", len(outs), all(torch.isfinite(o).all().item() for o in outs))
