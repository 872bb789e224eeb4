#!/bin/bash
# 8 GPUs with the final build: the driver's scaling command at N = 8 (without the CPU baseline leg)
set -x
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/au_b200_n8.json 2> gpurun_out/au_b200_n8.err ) 2> gpurun_out/au_b200_n8.time
grep real gpurun_out/au_b200_n8.time
tail -3 gpurun_out/au_b200_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/au_b200_n8.json").read().strip().splitlines()[-1])
    print("n8", round(d["value"]/1e9,3), round(d["ms_per_step"],2), d["e2e"]["value"]/1e9, (d.get("config5_tgv") or {}).get("value"), d.get("kernel_time_share"), d["clocks"])
except Exception as e:
    print("failed", e)
PY
