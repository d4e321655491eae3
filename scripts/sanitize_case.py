"""Smallest run that touches every kernel family / code path (for compute-sanitizer)."""
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
torch.cuda.synchronize()
print("ok", len(outs), all(torch.isfinite(o).all().item() for o in outs))
