"""three fused spec + fbank launches on 16,384 clips (target of the ncu captures in scripts/ncu_profiles_r2.sh)"""
import torch, sys
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
x = (torch.randn(16384, 16000, device="cuda") * 3000).round()
for _ in range(3):
    S.spec_fbank(x, S.R_SPEC, S.R_FBANK, layout="tf")
torch.cuda.synchronize()
