#!/bin/bash
# flux_shima_etal / flux_kennedy_gruber on node records: parity, A/B bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "shima or kennedy or shockcapturing or shock_capturing or blast or tuned_kernels or chandrashekar" > gpurun_out/ac_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ac_pytest.log
tail -8 gpurun_out/ac_pytest.log
B="python bench.py --workload euler_shima --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B > gpurun_out/ac_bench_shima_rec.json 2> gpurun_out/ac_bench_shima_rec.err
timeout 600 $B --kernel-path 2 > gpurun_out/ac_bench_shima_plain.json 2> gpurun_out/ac_bench_shima_plain.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ac_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("ac_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/ac_bench_shima_rec.err
