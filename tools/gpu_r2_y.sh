#!/bin/bash
# two GPUs, two processes: MPI-mortar goldens and the other multi-process tests, then the driver's N = 2 bench command
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/y_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/y_pytest_multi.log
tail -6 gpurun_out/y_pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/y_b200_n2.json 2> gpurun_out/y_b200_n2.err ) 2> gpurun_out/y_b200_n2.time
( time timeout 900 python bench.py > gpurun_out/y_b200_n1.json 2> gpurun_out/y_b200_n1.err ) 2> gpurun_out/y_b200_n1.time
python - <<'PY'
import json
for n in ("b200_n2","b200_n1"):
    try:
        d=json.load(open(f"gpurun_out/y_{n}.json"))
        print(n, round(d["value"]/1e9,3), round(d["ms_per_step"],2), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("value"), d["e2e"]["value"]/1e9, (d.get("config5_tgv") or {}).get("value"), d.get("kernel_time_share"), d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
