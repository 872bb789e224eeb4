#!/bin/bash
# MHD node records in the rotated frame: parity, bench, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "mhd" > gpurun_out/v_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/v_pytest.log
tail -8 gpurun_out/v_pytest.log
B="python bench.py --workload mhd_ec --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B > gpurun_out/v_bench_mhd_rec.json 2> gpurun_out/v_bench_mhd_rec.err
timeout 600 $B > gpurun_out/v_bench_mhd_rec_b.json 2> gpurun_out/v_bench_mhd_rec_b.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_fd3d_p3 -s 6 -c 1 -o gpurun_out/v_prof_mhd $B --steps 2 --warmup 1 > gpurun_out/v_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_interface_flux -s 6 -c 1 -o gpurun_out/v_prof_mhd_if $B --steps 2 --warmup 1 > gpurun_out/v_ncu_if.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("v_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
