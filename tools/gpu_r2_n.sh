#!/bin/bash
# what the driver runs, at N = 2: reference arm, then our arm (defaults), then N = 1 default
set -x
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/n_ref_n2.json 2> gpurun_out/n_ref_n2.err ) 2> gpurun_out/n_ref_n2.time
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n_b200_n2.json 2> gpurun_out/n_b200_n2.err ) 2> gpurun_out/n_b200_n2.time
( time timeout 900 python bench.py > gpurun_out/n_b200_n1.json 2> gpurun_out/n_b200_n1.err ) 2> gpurun_out/n_b200_n1.time
cat gpurun_out/n_*.time | grep real
tail -3 gpurun_out/n_b200_n2.err
python - <<'PY'
import json
for n in ("ref_n2","b200_n2","b200_n1"):
    try:
        d=json.load(open(f"gpurun_out/n_{n}.json"))
        print(n, round(d["value"]/1e9,3), round(d["ms_per_step"],2), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("value"), d["e2e"]["value"]/1e9, d["e2e"].get("host_affinity_rank0"), (d.get("config5_tgv") or {}).get("value"), d.get("kernel_time_share"))
    except Exception as e:
        print(n, "failed", e)
PY
