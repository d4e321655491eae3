"""tcgen05 FBANK kernel vs the classic kernel: throughput at several batch sizes (dev tool)."""
import sys, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
for nb in (64, 200, 296, 444, 740, 888, 1024, 1184, 2048, 4096):
    xb = (torch.randn(nb, 16000, device="cuda") * 3000).round()
    for name in ("R-FBANK", "C-FBANK"):
        p = S.PRESETS[name]
        out = []
        for tc in (1, 2):
            S.set_tuning(fbank_tc=tc)
            for _ in range(3): S.fbank(xb, p)
            reps = 10 if nb > 2000 else 50
            best = 1e9
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps): S.fbank(xb, p)
                b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / reps)
            out.append(round(nb / best / 1e3, 3))
        print(name, nb, "classic / tc Mclips/s", out, flush=True)
S.set_tuning()
