#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -6 gpurun_out/m_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B > gpurun_out/m_bench_single.json 2> gpurun_out/m_bench_single.err
timeout 600 $B --two-copy-face-flux > gpurun_out/m_bench_twocopy.json 2> gpurun_out/m_bench_twocopy.err
timeout 600 $B > gpurun_out/m_bench_single_b.json 2> gpurun_out/m_bench_single_b.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("m_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
