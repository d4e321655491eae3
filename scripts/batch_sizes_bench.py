"""every preset at BASELINE's batch sizes and at 16,384 clips, L2 flushed before each timed launch (dev tool)"""
import torch, sys, json
sys.path.insert(0, ".")
import speechrecognitionproject_b200 as S
import bench
x = (torch.randn(16384, 16000, device="cuda") * 3000).round()
rows = bench.preset_table(S, x, 6460.5)
for k, v in rows.items():
    print(k, round(v["clips_per_s"] / 1e6, 2), round(v["hbm_frac"], 3), round(v["ms_median"], 4))
