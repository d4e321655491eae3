"""BASELINE.json configs[4] / SURVEY 8d cfg5: model_mfcc_bgru-shaped inference fed by on-device MFCC.

The network below has the reference's architecture (models/model_mfcc_bgru.py:23-26: GRU(39, 512, 2 layers, bidirectional,
batch_first) + Linear(1024, 12)) with random weights (no checkpoints offline); its forward is the reference's loop
(:28-37) re-plumbed by patch_model: PCM -> fused MFCC kernel in the GRU's [B, 51, 39] layout -> cuDNN GRU -> fc.
Reports front-end and model time separately (the BGRU, ~0.65 GFLOP/clip, dominates)."""
import json, sys, types, torch, torch.nn as nn
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
from speechrecognitionproject_b200 import patch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mod = types.ModuleType("model_mfcc_bgru_shape")
mod.compute_mfcc = S.compute_mfcc
class Network(nn.Module):
    def __init__(self, num_features=512, num_layers=2):
        super().__init__()
        self.gru = nn.GRU(39, hidden_size=num_features, num_layers=num_layers, bidirectional=True, batch_first=True)
        self.fc = nn.Linear(num_features * 2, 12)
mod.Network = Network
patch.patch_model(mod, kind="mfcc_bgru")
net = mod.Network().cuda().eval()
x = (torch.randn(B, 16000, device="cuda") * 3000).round()
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    t_all = timed(lambda: net(x))
    t_fe = timed(lambda: S.mfcc(x, S.R_MFCC, layout="tf"))
    f = S.mfcc(x, S.R_MFCC, layout="tf")
    t_model = timed(lambda: patch._bgru_tail(net, f))
print(json.dumps({"batch": B, "ms_total": t_all, "ms_front_end": t_fe, "ms_model": t_model,
                  "clips_per_s_total": B / t_all * 1e3, "front_end_share": t_fe / t_all}))
