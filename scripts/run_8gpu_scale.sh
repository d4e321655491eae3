#!/bin/bash
# headline bench at 1/2/4/8 GPUs of one box (what the driver does at round end)
rm -f gpurun_out/r2_scale.jsonl
port=29700
for n in 1 2 4 8; do
  port=$((port+1))
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-presets --no-cpu-baseline >> gpurun_out/r2_scale.jsonl 2> gpurun_out/r2_scale_$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 >> gpurun_out/r2_scale.jsonl 2> gpurun_out/r2_scale_$n.err
  fi
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_scale.jsonl"):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["n_gpus"], round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,3), "i16", round(d["e2e_int16_ingest"]["value"]/1e6,3), "pageable", round(d["e2e_pageable_host"]["value"]/1e6,3), "h2d GB/s", round(d["e2e_h2d_ceiling"]["GBps"],1), "frac", round(d["e2e_h2d_ceiling"]["e2e_frac"],3))
PY
