#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B > gpurun_out/p_bench_single.json 2> gpurun_out/p_bench_single.err
timeout 600 $B --two-copy-face-flux > gpurun_out/p_bench_twocopy.json 2> gpurun_out/p_bench_twocopy.err
timeout 600 $B > gpurun_out/p_bench_single_b.json 2> gpurun_out/p_bench_single_b.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 1 -o gpurun_out/p_prof python bench.py --level 6 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/p_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/p_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("p_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
