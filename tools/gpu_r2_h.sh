#!/bin/bash
# 2 GPUs: full test-suite (incl. 2-process tests), weak-scaling bench, curved workloads
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log
tail -8 gpurun_out/h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err
python - <<'PY'
import json
for n in ("n1","n2"):
    try:
        d=json.load(open(f"gpurun_out/h_bench_{n}.json"))
        print(n, d["value"]/1e9, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["kernel_time_share"], d["clocks"]["sm_mhz"], d["e2e"]["value"]/1e9, d["finite"])
    except Exception as e:
        print(n, "failed", e)
PY
