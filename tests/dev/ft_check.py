"""tcgen05 FBANK kernel vs the classic CUDA-core kernel and the oracle (dev tool; run under `timeout`)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")      # run from the repo root: python tests/dev/<script>.py
import oracle
import speechrecognitionproject_b200 as S
from tests import helpers as H

def run(p, x, tc):
    S.set_tuning(fbank_tc=tc)
    y = S.fbank(x, p)
    torch.cuda.synchronize()
    return y

quick = "--quick" in sys.argv
xs = torch.from_numpy(oracle.synthetic_corpus(333, config_index=4)).cuda()
for name in ("R-FBANK", "C-FBANK"):
    p = S.PRESETS[name]
    for n in (1, 2, 7, 148, 149, 333):
        a = run(p, xs[:n], 2); b = run(p, xs[:n], 1)
        a2 = run(p, xs[:n], 2)
        d = (a - b).abs()
        big = b > b.amax(dim=(1, 2), keepdim=True) - 100.0
        print(name, n, "tc vs classic: max abs diff", float(d.max()), "inside the 100-unit domain", float(d[big].max()),
              "deterministic", bool(torch.equal(a, a2)), flush=True)
x8 = oracle.synthetic_corpus(8, config_index=0)
e = oracle.edge_suite(); xe = np.stack(list(e.values()))
for name in ("R-FBANK", "C-FBANK"):
    for tag, xx in (("corpus", x8), ("edge", xe)):
        for tc in (1, 2):
            got = run(S.PRESETS[name], torch.from_numpy(xx).cuda(), tc).cpu().numpy()
            truth = np.stack([oracle.fbank_truth(c, oracle.PRESETS[name]) for c in xx])
            print(name, tag, "tc" if tc == 2 else "classic", H.check_logmel(got, truth, name), flush=True)
if not quick:
    for nb in (1024, 16384):
        xb = (torch.randn(nb, 16000, device="cuda") * 3000).round()
        for name in ("R-FBANK", "C-FBANK"):
            for tc in (1, 2):
                p = S.PRESETS[name]
                for _ in range(3): run(p, xb, tc)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10 if nb > 2000 else 50
                a.record()
                for _ in range(reps): S.fbank(xb, p)
                b.record(); torch.cuda.synchronize()
                print(name, nb, "tc" if tc == 2 else "classic", "Mclips/s", round(nb * reps / a.elapsed_time(b) / 1e3, 3), flush=True)
S.set_tuning()
