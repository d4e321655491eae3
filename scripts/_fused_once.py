import torch, sys
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
x = (torch.randn(16384, 16000, device="cuda") * 3000).round()
for cpc in (1, 2, 4, 8):
    S.set_tuning(cpc=cpc)
    for _ in range(2):
        S.spec_fbank(x, S.R_SPEC, S.R_FBANK, layout="tf")
    torch.cuda.synchronize()
S.set_tuning()
S.spec(x, S.R_SPEC, layout="tf"); S.fbank(x, S.R_FBANK)
torch.cuda.synchronize()
