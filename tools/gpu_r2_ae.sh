#!/bin/bash
# SC line-sweep kernel with the subcell FV part as its own pass: parity, bench, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "shockcapturing or shock_capturing or blast or tuned_kernels or halo_exchange" > gpurun_out/ae_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ae_pytest.log
tail -6 gpurun_out/ae_pytest.log
B="python bench.py --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B --workload euler_sc > gpurun_out/ae_bench_sc.json 2> gpurun_out/ae_bench_sc.err
timeout 600 $B --workload euler_sc > gpurun_out/ae_bench_sc_b.json 2> gpurun_out/ae_bench_sc_b.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ae_launches_sc.csv $B --workload euler_sc --steps 2 --warmup 1 > gpurun_out/ae_launches.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ae_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("ae_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
