#!/bin/bash
# fast normal-direction LLF/HLL: whole suite, curved benches
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/ah_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ah_pytest.log
tail -8 gpurun_out/ah_pytest.log
B="python bench.py --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
for w in structured_curved p4est_curved structured_ec p4est_ec; do
  timeout 600 $B --workload $w > gpurun_out/ah_bench_$w.json 2> gpurun_out/ah_bench_$w.err
done
timeout 600 python bench.py --workload p4est_tgv_p5 --level 5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/ah_bench_p4est_tgv_p5.json 2> gpurun_out/ah_bench_p4est_tgv_p5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ah_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("ah_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
