#!/bin/bash
# indicator fused into the SC stage kernel + shima records: whole GPU suite, A/B benches
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/ad_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ad_pytest.log
tail -8 gpurun_out/ad_pytest.log
B="python bench.py --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B --workload euler_sc > gpurun_out/ad_bench_sc_fused.json 2> gpurun_out/ad_bench_sc_fused.err
timeout 600 $B --workload euler_sc --no-fused-cfl > gpurun_out/ad_bench_sc_unfused.json 2> gpurun_out/ad_bench_sc_unfused.err
timeout 600 $B --workload euler_shima > gpurun_out/ad_bench_shima_rec.json 2> gpurun_out/ad_bench_shima_rec.err
timeout 600 $B --workload euler_shima --kernel-path 2 > gpurun_out/ad_bench_shima_plain.json 2> gpurun_out/ad_bench_shima_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ad_launches_sc.csv $B --workload euler_sc --steps 2 --warmup 1 > gpurun_out/ad_launches.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ad_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("ad_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
