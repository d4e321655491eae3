#!/bin/bash
# cfg5 (model_mfcc_bgru end to end) at 1/2/4/8 GPUs of one box
rm -f gpurun_out/r2_cfg5.jsonl
port=29600
for n in 1 2 4 8; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      scripts/cfg5_model_mfcc_bgru.py --out gpurun_out/r2_cfg5.jsonl > gpurun_out/r2_cfg5_$n.log 2>&1 || tail -5 gpurun_out/r2_cfg5_$n.log
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_cfg5.jsonl"):
    d=json.loads(l); print(d["n_gpus"], round(d["clips_per_s_device_resident"]/1e6,3), round(d["clips_per_s_host_pageable_pcm"]/1e6,3), round(d["ms_front_end"],3), round(d["ms_model"],3), round(d["front_end_share"],4))
PY
