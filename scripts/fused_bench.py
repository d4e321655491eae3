"""spec + fbank: one fused launch vs two launches, per clips-per-group and batch size (dev tool)."""
import torch, sys, json
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
xall = (torch.randn(16384, 16000, device="cuda") * 3000).round()
for B in (1, 8, 64, 512, 2048, 16384):
    x = xall[:B]
    S.set_tuning()
    sep = t(lambda: (S.spec(x, S.R_SPEC, layout="tf"), S.fbank(x, S.R_FBANK)), 30)
    row = {"B": B, "us_two_launches": round(sep * 1e3, 1)}
    for cpc in (0, 16, 32, 64):
        S.set_tuning(cpc=cpc)
        row[f"us_fused_cpc{cpc or 'auto'}"] = round(t(lambda: S.spec_fbank(x, S.R_SPEC, S.R_FBANK, layout="tf"), 30) * 1e3, 1)
    print(json.dumps(row), flush=True)
S.set_tuning()
