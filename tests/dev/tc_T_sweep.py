"""tcgen05 MFCC kernel vs classic as a function of frames per clip and batch size (dev tool)."""
import sys, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
from dataclasses import replace
def rate(p, x, tc):
    S.set_tuning(mfcc_tc=tc)
    try:
        for _ in range(3): S.mfcc(x, p)
    except RuntimeError as e:
        return None
    best = 1e9
    reps = 10 if x.size(0) > 2000 else 40
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): S.mfcc(x, p)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return round(x.size(0) / best / 1e3, 3)
for name in ("R-MFCC", "C-MFCC"):
    p = S.PRESETS[name]
    for T in (4, 8, 16, 24, 32, 40, 51, 64, 80, 101):
        ns = p.hop * (T - 1)
        for nb in (16384, 1024, 64):
            x = (torch.randn(nb, ns, device="cuda") * 3000).round()
            assert S.out_shape(p, ns)[1] == T, (S.out_shape(p, ns), T)
            print(name, "T", T, "B", nb, "classic / tc Mclips/s", rate(p, x, 1), rate(p, x, 2), flush=True)
S.set_tuning()
