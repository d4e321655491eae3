"""Sweep launch configurations (srfe_set_tuning: warps / ctas / cpc) per preset. Dev tool."""
import os, sys, json, itertools, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(S.PRESETS)
x = (torch.randn(B, 16000, device="cuda") * 3000).round()
def run(fn, p, n=6):
    try:
        for _ in range(2): fn(x, p)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn(x, p)
        b.record(); torch.cuda.synchronize()
        return B * n / (a.elapsed_time(b) * 1e-3) / 1e6
    except Exception as e:
        return None
for name in names:
    from dataclasses import replace
    base, _, lay = name.partition(":")
    p = S.PRESETS[base]
    if lay: p = replace(p, layout=lay)
    fn = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}[type(p).__name__]
    S.set_tuning()
    res = [("default", run(fn, p))]
    cpcs = [0] if type(p).__name__ == "MfccParams" else [1, 2, 4, 8]
    grid = [(c, w, 0) for c in (1,) for w in range(8, 17)] + [(2, w, 0) for w in range(4, 9)] + [(3, w, 0) for w in range(3, 6)] + [(4, w, 0) for w in range(2, 5)] if os.environ.get("TUNE_FINE") else None
    for ctas, warps, pf in grid or [(2, 4, 0), (2, 5, 0), (2, 6, 0), (2, 7, 0), (2, 8, 0), (2, 4, 1), (2, 5, 1), (2, 6, 1), (1, 8, 0), (1, 10, 0), (1, 12, 0), (1, 14, 0), (1, 16, 0), (1, 8, 1), (1, 10, 1), (1, 12, 1), (3, 4, 0), (3, 5, 0)]:
        best = None
        for cpc in cpcs:
            S.set_tuning(warps=warps, ctas=ctas, cpc=cpc)
            r = run(fn, p)
            if r and (best is None or r > best[0]): best = (r, cpc)
        if best: res.append((f"ctas{ctas} w{warps} pf{pf} cpc{best[1]}", best[0]))
    res = [(k, v) for k, v in res if v]; res.sort(key=lambda t: -t[1])
    print(name, " | ".join(f"{k}: {v:.2f}" for k, v in res[:7] if v), "|| default:", [v for k, v in res if k == "default"])
