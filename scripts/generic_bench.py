import sys, json, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
B = 16384
x = (torch.randn(B, 16000, device="cuda") * 3000).round()
cases = {"mfcc 80 mels 13 coef 512/400/160": (S.mfcc, S.MfccParams(n_fft=512, win_length=400, hop=160, n_mels=80, n_mfcc=13)),
         "mfcc 40 mels 13 coef +d2": (S.mfcc, S.MfccParams(n_fft=512, win_length=400, hop=160, n_mels=40, n_mfcc=13, n_deltas=2)),
         "mfcc 128 mels 20 coef 640/320": (S.mfcc, S.MfccParams(n_mfcc=20)),
         "fbank 80 filt": (S.fbank, S.FbankParams(nfilt=80)), "fbank 26 filt": (S.fbank, S.FbankParams(nfilt=26)),
         "spec 512/384 overlap tf": (S.spec, S.SpecParams(nperseg=512, noverlap=384, layout="tf"))}
for name, (fn, p) in cases.items():
    for _ in range(3): y = fn(x, p)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): y = fn(x, p)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({"case": name, "ms": round(ms, 3), "Mclips_s": round(B / ms / 1e3, 2), "GBs": round(B / ms / 1e6 * S.bytes_per_clip(p), 0), "shape": list(y.shape[1:])}))
