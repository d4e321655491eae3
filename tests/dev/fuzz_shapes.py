"""Random parameter sets / clip lengths through the CUDA path vs the float64 oracle (dev tool; the regular parity
tests cover the presets and a few generic shapes, this walks many more launch shapes and DCT tile choices)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")      # run from the repo root: python tests/dev/<script>.py
import oracle
import speechrecognitionproject_b200 as S
from tests import helpers as H

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
worst = {"mfcc": 0.0, "fbank": 0.0, "spec": 0.0}
for case in range(n_cases):
    n_fft = int(rng.choice([512, 640]))
    n_samples = int(rng.choice([4000, 8000, 16000, 16000, 24000, 12346]))
    B = int(rng.choice([1, 3, 5]))
    x = oracle.synthetic_corpus(B, config_index=20 + case % 7, n_samples=n_samples)
    xd = torch.from_numpy(x).cuda()
    fam = rng.choice(["mfcc", "mfcc", "fbank", "spec"])
    try:
        if fam == "mfcc":
            n_mels = int(rng.choice([20, 26, 40, 64, 80, 128, 200]))
            p = S.MfccParams(n_fft=n_fft, win_length=int(rng.choice([n_fft, 400, 320, 256])), hop=int(rng.choice([80, 128, 160, 200, 320])),
                             n_mels=n_mels, n_mfcc=int(rng.integers(1, min(64, n_mels) + 1)), n_deltas=int(rng.integers(0, 3)),
                             layout=str(rng.choice(["ft", "tf"])))
            if p.win_length > n_fft: continue
            got = S.mfcc(xd, p).cpu().numpy()
            truth = H.oracle_batch(oracle.mfcc_truth, x, H.to_oracle_params(p))
            if p.layout == 'tf': truth = truth.transpose(0, 2, 1)          # the oracle is always [coeff, time]
            worst["mfcc"] = max(worst["mfcc"], H.check_mfcc(got, truth, str(p)))
        elif fam == "fbank":
            p = S.FbankParams(nfft=n_fft, frame_len=int(rng.choice([400, 320, 512])), frame_step=int(rng.choice([80, 160, 200])),
                              nfilt=int(rng.choice([13, 26, 40, 64, 120])))
            if p.frame_len > n_fft: continue
            got = S.fbank(xd, p).cpu().numpy()
            truth = H.oracle_batch(oracle.fbank_truth, x, H.to_oracle_params(p))
            st = H.check_logmel(got, truth, str(p))
            worst["fbank"] = max(worst["fbank"], st["max_err_in_domain"])
        else:
            p = S.SpecParams(nperseg=n_fft, noverlap=int(rng.choice([n_fft // 2, n_fft // 4, n_fft - 160, 0])), log=False,
                             layout=str(rng.choice(["ft", "tf"])))
            got = S.spec(xd, p).cpu().numpy()
            truth = H.oracle_batch(oracle.spec_truth, x, H.to_oracle_params(p))     # honours p.layout
            worst["spec"] = max(worst["spec"], H.check_psd(got, truth, str(p)))
    except RuntimeError as e:
        if "SRFE_ERR_TOO_LARGE" in str(e) or "SRFE_ERR_UNSUPPORTED" in str(e):
            print("rejected:", p, str(e)[:80]); continue
        raise
print(json.dumps({"cases": n_cases, "worst": worst}))
