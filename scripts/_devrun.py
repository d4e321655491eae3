import sys, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
xb = (torch.randn(16384, 16000, device="cuda") * 3000).round()
for name in ("C-MFCC", "R-MFCC", "C-MFCC-D2"):
    print(name, flush=True)
    S.mfcc(xb, S.PRESETS[name]); torch.cuda.synchronize()
