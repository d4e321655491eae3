"""tcgen05 MFCC kernel vs the classic CUDA-core kernel and the oracle (dev tool; run under `timeout`)."""
import sys, json, time, numpy as np, torch
sys.path.insert(0, ".")      # run from the repo root: python tests/dev/<script>.py
import oracle
import speechrecognitionproject_b200 as S
from dataclasses import replace

def run(p, x, tc):
    S.set_tuning(mfcc_tc=tc)
    y = S.mfcc(x, p)
    torch.cuda.synchronize()
    return y

res = {}
xs = torch.from_numpy(oracle.synthetic_corpus(333, config_index=4)).cuda()
for name in ("C-MFCC", "R-MFCC", "C-MFCC-D2"):
    for lay in ("ft", "tf"):
        p = replace(S.PRESETS[name], layout=lay)
        for n in (1, 7, 148, 333):
            a = run(p, xs[:n], 2); b = run(p, xs[:n], 1)
            d = float((a - b).abs().max())
            a2 = run(p, xs[:n], 2)
            res[f"{name}/{lay}/{n}"] = (d, bool(torch.equal(a, a2)))
            print(name, lay, n, "tc vs classic max abs diff", d, "deterministic", bool(torch.equal(a, a2)), flush=True)
x8 = oracle.synthetic_corpus(8, config_index=0)
e = oracle.edge_suite(); xe = np.stack(list(e.values()))
for name in ("C-MFCC", "R-MFCC"):
    for tag, xx in (("corpus", x8), ("edge", xe)):
        got = run(S.PRESETS[name], torch.from_numpy(xx).cuda(), 2).cpu().numpy()
        truth = np.stack([oracle.mfcc_truth(c, oracle.PRESETS[name]) for c in xx])
        print(name, tag, "tc vs truth max abs err", float(np.abs(got - truth).max()), flush=True)
xb = (torch.randn(16384, 16000, device="cuda") * 3000).round()
for name in ("C-MFCC", "R-MFCC", "C-MFCC-D2"):
    for tc in (1, 2):
        p = S.PRESETS[name]
        for _ in range(3): run(p, xb, tc)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): S.mfcc(xb, p)
        b.record(); torch.cuda.synchronize()
        print(name, "tc" if tc == 2 else "classic", "Mclips/s", round(16384 * 10 / a.elapsed_time(b) / 1e3, 3), flush=True)
S.set_tuning()
