#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "p5 or structured_3d_euler_ec or curved" > gpurun_out/k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest.log
tail -6 gpurun_out/k_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 $B --workload p4est_tgv_p5 --level 5 > gpurun_out/k_bench_p4est_tgv_p5_l5.json 2> gpurun_out/k_bench_p4est_tgv_p5_l5.err
timeout 600 $B --workload p4est_tgv_p5 --level 6 > gpurun_out/k_bench_p4est_tgv_p5_l6.json 2> gpurun_out/k_bench_p4est_tgv_p5_l6.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_curved_pn -s 6 -c 1 -o gpurun_out/k_prof_p5 python bench.py --workload p4est_tgv_p5 --level 5 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/k_ncu_p5.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/k_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("k_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["kernel_time_share"]["max_dt_ms"], d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
