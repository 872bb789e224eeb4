#!/bin/bash
# 8 GPUs: the driver's scaling command at N = 8 (defaults), once
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/q_topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/q_b200_n8.json 2> gpurun_out/q_b200_n8.err ) 2> gpurun_out/q_b200_n8.time
cat gpurun_out/q_b200_n8.time | grep real
tail -5 gpurun_out/q_b200_n8.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/q_b200_n8.json"))
    print("n8", round(d["value"]/1e9,3), round(d["ms_per_step"],2), d["e2e"]["value"]/1e9, d["e2e"].get("host_affinity_rank0"), (d.get("config5_tgv") or {}), d.get("kernel_time_share"), d["clocks"])
except Exception as e:
    print("failed", e)
PY
