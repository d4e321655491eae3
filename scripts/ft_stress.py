"""Stress of the tcgen05 FBANK kernel's lock-free tile pipeline (three tile buffers, two accumulators, MMAs issued by the
completing frame warp): thousands of launches with random batch sizes / presets / clip lengths / streams, several in flight
at once, every result compared bit for bit with a reference run of the same input (dev tool; run under `timeout`)."""
import sys, time, random, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
from dataclasses import replace
random.seed(1)
xall = (torch.randn(4096, 16000, device="cuda") * 3000).round()
xall[5] = 0; xall[77, 8000:] = 0
presets = [S.R_FBANK, S.C_FBANK, replace(S.R_FBANK, nfilt=26), replace(S.R_FBANK, frame_len=512, frame_step=128, nfilt=80)]
lengths = [16000, 15840, 8000]
S.set_tuning(fbank_tc=2)
ref = {(i, n): S.fbank(xall[:, :n].contiguous(), p) for i, p in enumerate(presets) for n in lengths}
xs = {n: xall[:, :n].contiguous() for n in lengths}
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(3)]
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < (float(sys.argv[1]) if len(sys.argv) > 1 else 40):
    jobs = []
    for st in streams:                       # three launches in flight on three streams
        i = random.randrange(len(presets)); ln = random.choice(lengths)
        b = random.choice([1, 2, 3, 7, 147, 148, 149, 295, 297, 1000, 4096]); o = random.randrange(0, 4096 - b + 1)
        with torch.cuda.stream(st):
            jobs.append((i, ln, b, o, S.fbank(xs[ln][o:o + b], presets[i])))
    torch.cuda.synchronize()
    for i, ln, b, o, y in jobs:
        if not torch.equal(y, ref[(i, ln)][o:o + b]):
            bad += 1; print("MISMATCH", i, ln, b, o, float((y - ref[(i, ln)][o:o + b]).abs().max()), flush=True)
        n += 1
S.set_tuning()
print(f"{n} launches, {bad} mismatches")
