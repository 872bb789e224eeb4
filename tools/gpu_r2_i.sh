#!/bin/bash
# 1 GPU: curved flux-differencing workloads (tuned vs generic), reference GPU benchmark config, level-7 ncu traffic
set -x
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1"
for w in structured_ec p4est_ec; do
  timeout 600 $B --workload $w --level 6 > gpurun_out/i_bench_${w}.json 2> gpurun_out/i_bench_${w}.err
  timeout 600 $B --workload $w --level 6 --generic-kernels > gpurun_out/i_bench_${w}_generic.json 2> gpurun_out/i_bench_${w}_generic.err
done
timeout 600 $B --workload p4est_tgv_p5 --level 5 > gpurun_out/i_bench_p4est_tgv_p5_l5.json 2> gpurun_out/i_bench_p4est_tgv_p5_l5.err
timeout 600 $B --workload p4est_tgv_p5 --level 6 > gpurun_out/i_bench_p4est_tgv_p5_l6.json 2> gpurun_out/i_bench_p4est_tgv_p5_l6.err
timeout 600 $B --workload mhd_ec --level 6 > gpurun_out/i_bench_mhd_ec.json 2> gpurun_out/i_bench_mhd_ec.err
timeout 600 $B --workload euler_sc --level 6 > gpurun_out/i_bench_euler_sc.json 2> gpurun_out/i_bench_euler_sc.err
timeout 900 ncu --set full --clock-control none -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 1 -o gpurun_out/i_prof_l7 python bench.py --level 7 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/i_ncu_l7.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_curved -s 6 -c 1 -o gpurun_out/i_prof_curved python bench.py --workload p4est_ec --level 6 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/i_ncu_curved.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/i_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("i_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
