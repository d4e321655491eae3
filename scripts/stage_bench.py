"""spectrogram: TMA-staged frames (stage=2) vs direct global loads (stage=1); dev tool"""
import torch, sys, json
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
from dataclasses import replace
x = (torch.randn(16384, 16000, device="cuda") * 3000).round()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, p, xx in (("C-SPEC tf", replace(S.C_SPEC, layout="tf"), x), ("C-SPEC ft", S.C_SPEC, x), ("C-SPEC tf i16", replace(S.C_SPEC, layout="tf"), x.to(torch.int16)),
                    ("R-SPEC tf i16", replace(S.R_SPEC, layout="tf"), x.to(torch.int16)), ("R-SPEC ft i16", S.R_SPEC, x.to(torch.int16))):
    r = {}
    outs = {}
    for st in (1, 2):
        S.set_tuning(stage=st)
        try:
            for n in (1, 7, 333, 16384):
                outs[(st, n)] = S.spec(xx[:n], p)
            r["direct" if st == 1 else "staged"] = round(16384 / t(lambda: S.spec(xx, p)) / 1e3, 2)
        except RuntimeError as e:
            r["staged"] = str(e)[:60]
    eq = all(torch.equal(outs[(1, n)], outs[(2, n)]) for n in (1, 7, 333, 16384) if (2, n) in outs)
    print(name, r, "bit-identical" if eq else "MISMATCH", flush=True)
S.set_tuning()
