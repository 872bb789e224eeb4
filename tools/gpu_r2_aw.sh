#!/bin/bash
# core flux switch in the line-sweep kernels: suite, shock-capturing workload
mkdir -p gpurun_out
timeout 110 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 60 python bench.py --workload euler_sc --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/aw_bench_euler_sc.json 2> gpurun_out/aw_bench_euler_sc.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/aw_bench_euler_sc.json").read().strip().splitlines()[-1])
print("euler_sc", round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
PY
