"""Quick per-preset kernel timing (device-resident input, CUDA events). Dev tool, not the bench."""
import os, sys, json, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(S.PRESETS)
x = (torch.randn(B, 16000, device="cuda") * 3000).round()
if len(sys.argv) > 3:                      # e.g. "warps=8,ctas=2,cpc=1" -> srfe_set_tuning
    S.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in sys.argv[3].split(","))})
PEAK = 6460.5
for name in names:
    from dataclasses import replace
    base, _, lay = name.partition(":")
    p = S.PRESETS[base]
    if lay: p = replace(p, layout=lay)
    fn = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}[type(p).__name__]
    for _ in range(3):
        y = fn(x, p)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record(); y = fn(x, p); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    ms = ts[len(ts) // 2]
    cps = B / ms * 1e3
    gbs = cps * S.bytes_per_clip(p) / 1e9
    print(json.dumps({"preset": name, "B": B, "ms": round(ms, 3), "best_ms": round(ts[0], 3), "Mclips_s": round(cps / 1e6, 3),
                      "GBs": round(gbs, 1), "hbm_frac": round(gbs / PEAK, 4)}))
