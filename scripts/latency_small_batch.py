"""Per-call latency at the reference's own batch sizes (ensemble drivers use batch_size=1). Dev tool."""
import sys, time, json, torch
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
for name in ("R-MFCC", "R-SPEC", "R-FBANK", "C-MFCC"):
    p = S.PRESETS[name]
    fn = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}[type(p).__name__]
    for B in (1, 8, 64, 512):
        x = (torch.randn(B, 16000, device="cuda") * 3000).round()
        for _ in range(20): fn(x, p)
        torch.cuda.synchronize()
        n = 300
        t0 = time.perf_counter()
        for _ in range(n): y = fn(x, p)
        t_issue = (time.perf_counter() - t0) / n          # host time to issue (async)
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / n
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y = fn(x, p); b.record(); torch.cuda.synchronize()
        print(json.dumps({"preset": name, "B": B, "host_issue_us": round(t_issue * 1e6, 1), "per_call_us": round(t_all * 1e6, 1),
                          "kernel_us": round(a.elapsed_time(b) * 1e3, 1)}))

# the same call captured in a CUDA graph (three front ends of an ensemble step, batch 1)
x = (torch.randn(1, 16000, device="cuda") * 3000).round()
fns = lambda: (S.mfcc(x, S.R_MFCC, layout="tf"), S.fbank(x, S.R_FBANK), S.spec(x, S.R_SPEC, layout="tf"))
fns(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    outs = fns()
for _ in range(20): g.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): g.replay()
torch.cuda.synchronize()
t_graph = (time.perf_counter() - t0) / 300
t0 = time.perf_counter()
for _ in range(300): fns()
torch.cuda.synchronize()
t_eager = (time.perf_counter() - t0) / 300
print(json.dumps({"ensemble_step_B1_three_front_ends": {"eager_us": round(t_eager * 1e6, 1), "cuda_graph_us": round(t_graph * 1e6, 1)}}))

# reference-named single-clip functions on a CPU tensor (H2D -> kernel -> D2H inside the call), as a per-clip loop would call them
xc = (torch.randn(16000) * 3000).round()
for name, fn in (("compute_mfcc", S.compute_mfcc), ("compute_spec", S.compute_spec), ("filter_banks", S.filter_banks)):
    for _ in range(20): fn(xc)
    t0 = time.perf_counter()
    for _ in range(300): y = fn(xc)
    print(json.dumps({"single_clip_cpu_tensor": name, "per_call_us": round((time.perf_counter() - t0) / 300 * 1e6, 1), "shape": list(y.shape)}))
