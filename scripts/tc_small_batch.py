"""C-MFCC per-call time at small batches: tcgen05 kernel vs classic kernel (dev tool)"""
import sys, time, json, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
for B in (1, 8, 64, 148, 296, 512, 1024, 2048):
    x = (torch.randn(B, 16000, device="cuda") * 3000).round()
    row = {"B": B}
    for tc in (1, 2):
        S.set_tuning(mfcc_tc=tc)
        for _ in range(20): S.mfcc(x, S.C_MFCC)
        torch.cuda.synchronize()
        evs = []
        for _ in range(50):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); S.mfcc(x, S.C_MFCC); b.record(); evs.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        row["classic_us" if tc == 1 else "tc_us"] = round(ts[len(ts) // 2] * 1e3, 1)
    print(json.dumps(row), flush=True)
S.set_tuning()
