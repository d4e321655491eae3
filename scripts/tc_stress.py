"""Stress of the tcgen05 MFCC kernel's lock-free clip pipeline (ring, counters, mbarriers): thousands of launches with random
batch sizes / presets / streams, every result compared bit for bit with a reference run of the same input (dev tool; run
under `timeout`)."""
import sys, time, random, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
from dataclasses import replace
random.seed(0)
xall = (torch.randn(4096, 16000, device="cuda") * 3000).round()
xall[5] = 0; xall[77, 8000:] = 0
presets = [S.C_MFCC, S.C_MFCC_D2, replace(S.C_MFCC, layout="tf"), replace(S.C_MFCC, n_mels=64, n_mfcc=20, n_deltas=1),
           replace(S.R_MFCC, n_deltas=2)]
S.set_tuning(mfcc_tc=2)
ref = {i: S.mfcc(xall, p) for i, p in enumerate(presets)}
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(3)]
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < float(sys.argv[1]) if len(sys.argv) > 1 else 40:
    i = random.randrange(len(presets)); b = random.choice([1, 2, 3, 7, 147, 148, 149, 295, 297, 1000, 4096]); o = random.randrange(0, 4096 - b + 1)
    st = random.choice(streams)
    with torch.cuda.stream(st):
        y = S.mfcc(xall[o:o + b], presets[i])
    st.synchronize()
    if not torch.equal(y, ref[i][o:o + b]):
        bad += 1; print("MISMATCH", i, b, o, float((y - ref[i][o:o + b]).abs().max()), flush=True)
    n += 1
S.set_tuning()
print(f"{n} launches, {bad} mismatches")
