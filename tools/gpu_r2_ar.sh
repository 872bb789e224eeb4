#!/bin/bash
# two GPUs with the final build: multi-process tests, N = 2 bench line (driver command)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/ar_pytest_multi.log 2>&1
tail -4 gpurun_out/ar_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/ar_b200_n2.json 2> gpurun_out/ar_b200_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ar_b200_n2.json").read().strip().splitlines()[-1])
print(round(d["value"]/1e9,3), round(d["ms_per_step"],2), d["e2e"]["value"]/1e9, d.get("halo"), d.get("kernel_time_share"), d["clocks"])
PY
